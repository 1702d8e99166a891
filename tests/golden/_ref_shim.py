"""Import the UNMODIFIED reference package from /root/reference on CPU (build container only).

The reference needs five third-party modules that are absent here and that never touch the
hot-path arithmetic; they are stubbed.  Used by make_golden.py; nothing under tests/ that runs
on the GPU box imports this file."""
import logging
import sys
import types

REF_ROOT = "/root/reference"


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)

    stub("colorlog", ColoredFormatter=lambda fmt, datefmt=None: logging.Formatter("%(message)s"))
    stub("colored_traceback", Colorizer=object)
    stub("h5py")
    stub("imagesize")
    stub("overrides", EnforceOverrides=type("EnforceOverrides", (), {}), overrides=lambda f: f)
    # make sure the reference (not this repo's drop-in of the same name) is what gets imported
    for k in [k for k in sys.modules if k == "ssdn" or k.startswith("ssdn.")]:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT + "/ssdn")
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
    import ssdn  # noqa
    return ssdn
