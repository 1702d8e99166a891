"""CPU: the C-ABI shared library builds, loads without a GPU and exports every symbol include/ssdn_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "ssdn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssdn_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(engine):
    lib = ctypes.CDLL(engine.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_binding_matches_header(engine):
    assert sorted(engine.EXPORTS) == _header_symbols()


def test_version_and_error_string(engine):
    lib = engine.lib()
    assert lib.ssdn_b200_version() >= 100
    assert isinstance(lib.ssdn_b200_last_error(), bytes)


def test_argument_validation_needs_no_gpu(engine):
    import pytest
    h = ctypes.c_void_p()
    with pytest.raises(ValueError):
        engine.check(engine.lib().ssdn_net_create(2, 3, 9, 48, 48, 1, ctypes.byref(h)))       # not a multiple of 32
    with pytest.raises(ValueError):
        engine.check(engine.lib().ssdn_net_create(2, 3, 9, 64, 32, 1, ctypes.byref(h)))       # blind-spot needs squares
    engine.check(engine.lib().ssdn_net_create(2, 3, 9, 64, 64, 1, ctypes.byref(h)))
    assert engine.lib().ssdn_net_param_count(h) == 1269129
    assert engine.lib().ssdn_net_workspace_bytes(h) > 0
    engine.lib().ssdn_net_destroy(h)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "selfsupervised-denoising_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "ssdn_oracle" not in open(os.path.join(dirpath, f)).read(), f


def test_python_binding_arity_matches_header(engine):
    """Every ctypes signature in ssdn/_engine.py has exactly as many arguments as the prototype in include/ssdn_b200.h
    (a drifted prototype would corrupt the call silently: ctypes cannot check it)."""
    text = open(os.path.join(ROOT, "include", "ssdn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\b(ssdn_\w+)\s*\(([^)]*)\)\s*;", text))
    assert set(protos) == set(engine.EXPORTS)
    from ssdn import _engine
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(_engine._SIGNATURES[name][1]), (name, n, len(_engine._SIGNATURES[name][1]))
