"""Importable alias of the ``md-structure-factor_b200`` package (its name is not an identifier)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("md-structure-factor_b200")
native, dens, load_traj, distributed = _pkg.native, _pkg.dens, _pkg.load_traj, _pkg.distributed
plot2d_gpu = _pkg.plot2d_gpu
