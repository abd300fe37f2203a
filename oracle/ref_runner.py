"""Harness around oracle/_ref/dens.py -- the UNMODIFIED reference file (see build_ref.py).

TEST / BASELINE INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's CPU legs, never by the
product package.  The harness changes nothing numeric: it silences the two tqdm bars
(reference dens.py:277,283), swallows the final ``np.savez_compressed`` (dens.py:346; file output is
outside the frames/s metric and takes minutes at 512^3) and takes time stamps around
``np.fft.rfftn`` (dens.py:313) so that the per-frame loop can be timed on its own.
"""
import importlib.util
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mod = None


def available():
    return os.path.exists(os.path.join(REF_DIR, "dens.py")) and os.path.exists(os.path.join(REF_DIR, "past", "utils.py"))


def load():
    """Import oracle/_ref/dens.py as module ``mdsf_reference_dens`` (the product also has a dens.py)."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise ImportError("oracle/_ref/ is missing: run `python oracle/build_ref.py` where /root/reference exists")
    sys.path.insert(0, REF_DIR)          # for the `past` shim only
    try:
        spec = importlib.util.spec_from_file_location("mdsf_reference_dens", os.path.join(REF_DIR, "dens.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REF_DIR)

    class _Quiet:                        # stands in for the tqdm module inside the reference only
        @staticmethod
        def tqdm(it=None, *a, **k):
            _Quiet.calls.append(time.perf_counter())
            return it
        calls = []

    mod.tqdm = _Quiet
    mod.trange = lambda *a, **k: range(*a)
    _mod = mod
    return mod


def run(coords, dims, typ, rad, ucell, sres, keep_output=True):
    """compute_sf of the unmodified reference.  Returns dict(out=<the six npz arrays or None>, loop_s=<seconds from
    the start of the frame loop (dens.py:277) to the return of the last rfftn>, frames=T, d1=[...] if keep_output)."""
    ref = load()
    ref.tqdm.calls.clear()
    real_rfftn, real_save = np.fft.rfftn, np.savez_compressed
    stamps, d1, saved = [], [], {}

    def spy(a, *args, **kw):
        if keep_output:
            d1.append(np.array(a, copy=True))
        out = real_rfftn(a, *args, **kw)
        stamps.append(time.perf_counter())
        return out

    def swallow(name, **arrays):
        if keep_output:
            saved.update(arrays)

    np.fft.rfftn, np.savez_compressed = spy, swallow
    try:
        ref.compute_sf(coords, dims, typ, "unused", rad, ucell, sres)
    finally:
        np.fft.rfftn, np.savez_compressed = real_rfftn, real_save
    t0 = ref.tqdm.calls[0]               # the frame loop's own tqdm call is the first one
    return dict(out=saved if keep_output else None, loop_s=stamps[-1] - t0, frames=len(stamps), d1=d1)
