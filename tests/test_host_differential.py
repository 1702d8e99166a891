"""CPU: the host-side interface of the drop-in package against the UNMODIFIED reference, call for call.

tests/host_battery.py is ONE script written against the reference's public API (configuration inference and run names,
LR ramp, metric bookkeeping, image helpers, NoisyDataset incl. padding / metadata / RNG consumption, samplers, Noise2Void
masking and loss, noise styles, Shift2d / Crop2d, NoiseNetwork parameter schema and seeded initialisation, Denoiser
state-dict schema, axis-order strings, tensor <-> PIL conversion, the image-folder data set on files it writes itself and
the training / test loaders the trainer builds from a configuration; 299 entries, error types included).  It runs unchanged on either package; the reference's answers
are committed as tests/golden/host_battery_reference.json.gz so the comparison also runs where /root/reference does not
exist."""
import gzip
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "host_battery_reference.json.gz")


def _run(mode: str) -> dict:
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    out = subprocess.run([sys.executable, os.path.join(HERE, "host_battery.py"), mode], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("JSON ")][-1][5:])


def _golden() -> dict:
    with gzip.open(GOLDEN, "rb") as f:
        return json.loads(f.read().decode())


def _differences(a, b, path="", out=None):
    """Structural comparison.  Tensors are described by shape, dtype, a hash of their bytes and their sum: equal hashes
    (what this container produces) or, on a host whose CPU kernels round differently, sums equal to 1e-5."""
    out = [] if out is None else out
    if isinstance(a, dict) and isinstance(b, dict):
        if "sha256" in a and "sha256" in b and "shape" in a and "shape" in b:
            if a["shape"] != b["shape"] or a["dtype"] != b["dtype"]:
                out.append((path, a["shape"], a["dtype"], b["shape"], b["dtype"]))
            elif a["sha256"] != b["sha256"] and not abs(a["sum"] - b["sum"]) <= 1e-5 * max(1.0, abs(a["sum"])):
                out.append((path, a["sum"], b["sum"]))
            return out
        if list(a) != list(b):
            out.append((path, "keys", [k for k in a if k not in b][:4], [k for k in b if k not in a][:4]))
        for k in a:
            if k in b:
                _differences(a[k], b[k], path + "/" + k, out)
    elif isinstance(a, list) and isinstance(b, list):
        if len(a) != len(b):
            out.append((path, "length", len(a), len(b)))
        for i, (x, y) in enumerate(zip(a, b)):
            _differences(x, y, "{}/{}".format(path, i), out)
    elif isinstance(a, float) or isinstance(b, float):
        if not (isinstance(a, (int, float)) and isinstance(b, (int, float))) or not abs(a - b) <= 1e-9 * max(1.0, abs(a), abs(b)):
            out.append((path, a, b))
    elif a != b:
        out.append((path, a, b))
    return out


def test_mirror_answers_every_battery_call_like_the_reference():
    gold = _golden()
    mine = _run("mirror")
    assert len(gold) >= 299 and list(mine) == list(gold)
    diff = _differences(gold, mine)
    assert not diff, diff[:10]
    # the battery is not vacuous: only the calls that must fail do fail, and with the reference's exception types
    failing = {k: v for k, v in gold.items() if isinstance(v, str) and v.startswith("raises ")}
    assert failing == {
        "cfg.infer_datasets./data/unknown_things./data/kodak": "raises ValueError", "cfg.test_length.bsds300": "raises KeyError",
        "cfg.test_length.ilsvrc": "raises KeyError", "cfg.test_length.nothing": "raises KeyError", "rotate.45": "raises NotImplementedError",
        "rotate.-90": "raises NotImplementedError", "rotate.360": "raises NotImplementedError", "dataset.badstyle": "raises NotImplementedError",
        "n2v.manipulate.even": "raises ValueError", "noise.speckle3": "raises NotImplementedError", "data_format.bad": "raises AssertionError",
        "tensor2image.two_channels": "raises NotImplementedError", "folder.empty": "raises RuntimeError"}


@pytest.mark.skipif(not os.path.isdir("/root/reference/ssdn"), reason="the reference only exists in the build container")
def test_committed_battery_answers_are_the_references():
    diff = _differences(_run("reference"), _golden())
    assert not diff, diff[:10]
