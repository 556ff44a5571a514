"""Load the UNMODIFIED reference sources from ``/root/reference`` (build container only).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference`` does not exist on the
GPU box: nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call into this module at
run time; it exists to (a) generate ``tests/golden/*.npz`` and (b) re-validate the restatement in
the CPU test-suite whenever the reference tree is mounted.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

from . import tf_shim

REFERENCE_ROOT = os.environ.get("TKBO_REFERENCE_ROOT", "/root/reference")
_PKG = "_tkbo_reference_pkg"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "fourier_features.py"))


def _load(modname, relpath, package):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = package
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """-> namespace(fourier_features, pes, rank_pes, dts) of reference modules.

    ``rank_pes.py`` and ``indiff_pes.py`` need NumPy/SciPy only.  The other three import tensorflow / tfp / gpflow at
    module scope, which resolve to ``tf_shim`` (the package ``__init__`` chain is bypassed because it
    pulls in the training code and EI, which are outside the hot path)."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    if _PKG + ".acquisitions.dts" in sys.modules:
        m = sys.modules
        return types.SimpleNamespace(fourier_features=m[_PKG + ".fourier_features"],
                                     pes=m[_PKG + ".acquisitions.pes"],
                                     rank_pes=m[_PKG + ".acquisitions.rank_pes"],
                                     indiff_pes=m[_PKG + ".acquisitions.indiff_pes"],
                                     dts=m[_PKG + ".acquisitions.dts"])
    tf_shim.install()
    pkg = types.ModuleType(_PKG)
    pkg.__path__ = [REFERENCE_ROOT]
    sys.modules[_PKG] = pkg
    acq = types.ModuleType(_PKG + ".acquisitions")
    acq.__path__ = [os.path.join(REFERENCE_ROOT, "acquisitions")]
    sys.modules[_PKG + ".acquisitions"] = acq
    ff = _load(_PKG + ".fourier_features", "fourier_features.py", _PKG)
    pkg.fourier_features = ff
    rank_pes = _load(_PKG + ".acquisitions.rank_pes", "acquisitions/rank_pes.py", _PKG + ".acquisitions")
    indiff_pes = _load(_PKG + ".acquisitions.indiff_pes", "acquisitions/indiff_pes.py", _PKG + ".acquisitions")
    pes = _load(_PKG + ".acquisitions.pes", "acquisitions/pes.py", _PKG + ".acquisitions")
    dts = _load(_PKG + ".acquisitions.dts", "acquisitions/dts.py", _PKG + ".acquisitions")   # `from .. import fourier_features`
    return types.SimpleNamespace(fourier_features=ff, pes=pes, rank_pes=rank_pes, indiff_pes=indiff_pes, dts=dts)
