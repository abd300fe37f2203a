"""md-structure-factor_b200 -- B200-native engine behind the reference's ``dens.compute_sf``.

The directory mirrors the reference's flat module layout (``dens.py``, ``load_traj.py``,
``main_gromacs.py``), so it can be put on ``PYTHONPATH`` as a drop-in, or imported as a package
through ``importlib.import_module("md-structure-factor_b200")`` (the name is not a Python
identifier) or the ``mdsf_b200`` alias module at the repository root.  Either way the same module
objects are used: this ``__init__`` puts the directory on ``sys.path`` and re-exports the flat
modules.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import mdsf_native as native  # noqa: E402
import dens  # noqa: E402
import npz_writer  # noqa: E402
import load_traj  # noqa: E402
import sf_distributed as distributed  # noqa: E402
import plot2d_gpu  # noqa: E402

__all__ = ["native", "dens", "load_traj", "distributed", "plot2d_gpu"]
