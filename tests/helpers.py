"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if not p.endswith("helpers.npz"))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    c = {k: z[k] for k in z.files}
    c["rad"] = {str(l): (float(a), float(b)) for l, a, b in zip(c["rad_labels"], c["rad_nel"], c["rad_sigma"])}
    c["typ"] = np.array([str(t) for t in c["typ"]])
    c["sres"] = float(c["sres"])
    return c


def sf_errors(sf, ref):
    """(worst per-bin relative error, worst error normalised by max(ref))."""
    diff = np.abs(sf - ref)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(ref != 0, diff / np.abs(ref), np.where(diff == 0, 0.0, np.inf))
    return float(rel.max()), float(diff.max() / np.abs(ref).max())
