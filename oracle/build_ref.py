#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference hot path, runnable on the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product package).

The reference's hot path is one Python file (dens.py) whose only missing import in this image is
``past.utils.old_div`` (dens.py:4).  This script copies ``dens.py`` byte for byte from
``/root/reference`` into the git-ignored ``oracle/_ref/`` (it travels to the GPU box with the
snapshot, like a built .so) and writes the 6-line ``past/utils.py`` shim next to it.  Nothing of the
reference enters the repository history.  ``oracle/ref_runner.py`` imports the copy.

Run by ``__graft_entry__.build()`` whenever /root/reference is present (the build container).
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MDSF_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

SHIM = '''"""Stand-in for future's past.utils (not installed here): old_div with the same semantics."""
import numbers


def old_div(a, b):
    if isinstance(a, numbers.Integral) and isinstance(b, numbers.Integral):
        return a // b
    return a / b
'''


def main():
    src = os.path.join(REF, "dens.py")
    if not os.path.exists(src):
        print("build_ref: %s not present; keeping whatever oracle/_ref/ already holds" % src)
        return 0
    os.makedirs(os.path.join(OUT, "past"), exist_ok=True)
    shutil.copyfile(src, os.path.join(OUT, "dens.py"))
    shutil.copyfile(os.path.join(REF, "radii.txt"), os.path.join(OUT, "radii.txt"))
    with open(os.path.join(OUT, "past", "__init__.py"), "w") as fh:
        fh.write("")
    with open(os.path.join(OUT, "past", "utils.py"), "w") as fh:
        fh.write(SHIM)
    with open(os.path.join(OUT, "SHA256"), "w") as fh:
        fh.write(hashlib.sha256(open(src, "rb").read()).hexdigest() + "  dens.py\n")
    print("build_ref: oracle/_ref/dens.py <- %s" % src)
    return 0


if __name__ == "__main__":
    sys.exit(main())
