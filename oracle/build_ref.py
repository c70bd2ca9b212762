"""Compile the reference's own funk-SVD/RSVD extension into ``oracle/_ref/``.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

The reference's only native code on the hot path is
``/root/reference/util/matrix_factorization.pyx`` (built by its ``setup.py:48-52``
as ``util.matrix_factorization``).  This recipe cythonizes that file *where it
lies* (the reference tree is read-only and is never copied into the repo):
the generated C and the built module go to ``oracle/_ref/`` only, which is
git-ignored but travels to the GPU box with the snapshot.  Nothing is built
when ``/root/reference`` is absent (the GPU box) -- the prebuilt files are used.

    python oracle/build_ref.py           # build if missing / stale
    python -c "from oracle.build_ref import load_ref; m = load_ref(); m.SVD"
"""
from __future__ import annotations

import glob
import importlib.util
import os
import subprocess
import sys
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DAISY_REFERENCE", "/root/reference")
REF_PYX = os.path.join(REF_ROOT, "util", "matrix_factorization.pyx")
OUT_DIR = os.path.join(_HERE, "_ref")
MOD_NAME = "matrix_factorization"


def _built_so():
    hits = glob.glob(os.path.join(OUT_DIR, MOD_NAME + "*.so"))
    return hits[0] if hits else None


def build(force=False, quiet=True):
    """Returns the path of the built module, or None when it cannot be built here."""
    so = _built_so()
    if not os.path.exists(REF_PYX):
        return so                                    # GPU box: use whatever travelled
    if so and not force and os.path.getmtime(so) >= os.path.getmtime(REF_PYX):
        return so
    import numpy as np
    os.makedirs(OUT_DIR, exist_ok=True)
    c_file = os.path.join(OUT_DIR, MOD_NAME + ".c")
    # 1. Cython: .pyx (read in place) -> C in oracle/_ref/   (language_level=3 as Cython 3 defaults)
    subprocess.check_call([sys.executable, "-m", "cython", "-3", REF_PYX, "-o", c_file],
                          stdout=subprocess.DEVNULL if quiet else None)
    # 2. gcc with CPython's own extension flags + numpy include dir (setup.py:50-52 passes np.get_include())
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT_DIR, MOD_NAME + ext)
    cflags = (sysconfig.get_config_var("CFLAGS") or "-O2").split()
    cmd = ["gcc", "-shared", "-fPIC", *cflags, "-w",
           "-I", sysconfig.get_paths()["include"], "-I", np.get_include(),
           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", c_file, "-o", so]
    subprocess.check_call(cmd)
    return so


def load_ref():
    """Import the compiled reference module (classes ``SVD``, ``RSVD``, ``SVDpp``); None if unavailable."""
    so = build()
    if not so:
        return None
    full = "oracle._ref." + MOD_NAME
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(MOD_NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[full] = mod
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, quiet=False))
