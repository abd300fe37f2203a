#!/usr/bin/env python
"""bench.py -- frames/s of the structure-factor hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference]

A step is one pass of the whole per-frame path (rescale/wrap/cell index, deterministic Gaussian
splat, 3-D FFT, |F|^2 accumulation) over one batch of ``--frames-per-step`` synthetic frames.
* ``value``  : frames/s with the frame pool already resident in HBM, timed with CUDA events.
* ``e2e``    : frames/s through the C ABI (Engine.push_frames, the call dens.compute_sf makes) from
               pinned HOST memory, H2D copies and the final S(q) device->host read inside the timed
               region.
* ``roofline``: SURVEY 8(d) algorithmic bytes x frames/s of the whole step vs the measured HBM peak, with the
               dominant kernel's own figure (its CUDA-event launch duration) beside it.
* ``cpu_baseline``: the numpy restatement of the reference's loop (oracle/, kind "port") on a
               bounded sample of the same workload, on one host core (the reference is a serial loop).
``--impl reference`` times that same port on ALL host cores (one frame per worker process and step).
Multi-GPU (torchrun, one rank per GPU): frames shard across ranks with no data-path collective
(weak scaling: every rank runs the same per-GPU work) and one NCCL reduce of the partial S(q)
closes the timed region; the time is the max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--frames-per-step", type=int, default=64)
    ap.add_argument("--pool", type=int, default=128, help="distinct synthetic frames cycled through")
    ap.add_argument("--cpu-frames", type=int, default=3, help="frames of the CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm: worker processes (0 = all host cores)")
    ap.add_argument("--ref-budget", type=float, default=200.0, help="reference arm: stop timing after this many seconds")
    ap.add_argument("--fft", default="auto")
    ap.add_argument("--tile", default="0x0")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_pool(wl, nframes):
    from importlib import import_module
    workloads = import_module("workloads")
    return workloads.jitter_frames(wl["base"], wl["box"], nframes, wl["jitter"], wl["seed0"])


# ----------------------------------------------------------------------------- CPU arm
def cpu_port_frames_per_s(wl, nframes):
    """The reference's per-frame loop as restated in oracle/dens_oracle.py, 1 thread (the reference
    is a serial Python loop over atoms)."""
    from oracle import dens_oracle as orc
    coords = make_pool(wl, nframes)
    dims = np.repeat(wl["box"][None, :], nframes, axis=0)
    stamps = []
    t0 = time.perf_counter()
    orc.structure_factor(coords, dims, wl["typ"], wl["rad"], wl["ucell"], wl["sres"], frame_callback=lambda t: stamps.append(time.perf_counter()))
    per_frame = (stamps[-1] - t0) / nframes       # frame loop only; plot grids / npz excluded like the GPU number
    return 1.0 / per_frame, per_frame


_REF_WL = None


def _ref_worker(args):
    """One frame of the workload through the oracle port in a worker process; returns the frame-loop seconds."""
    name, seed = args
    global _REF_WL
    from importlib import import_module
    from oracle import dens_oracle as orc
    workloads = import_module("workloads")
    if _REF_WL is None or _REF_WL[0] != name:
        _REF_WL = (name, workloads.get(name))
    wl = _REF_WL[1]
    coords = workloads.jitter_frames(wl["base"], wl["box"], 1, wl["jitter"], wl["seed0"] + seed)
    dims = wl["box"][None, :]
    stamps = []
    t0 = time.perf_counter()
    orc.structure_factor(coords, dims, wl["typ"], wl["rad"], wl["ucell"], wl["sres"], frame_callback=lambda t: stamps.append(time.perf_counter()))
    return stamps[-1] - t0


def run_reference_arm(args, wl, rank):
    """The reference's CPU path (its numpy restatement in oracle/: the reference is Python and /root/reference does not
    exist on the GPU box) on ALL host cores: the reference itself is one serial loop, frames are independent, so one
    step = one frame per worker process, timed by wall clock."""
    if rank != 0:
        return
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    procs = args.ref_procs or (os.cpu_count() or 1)
    t_begin = time.perf_counter()
    times, single = [], []
    with ProcessPoolExecutor(max_workers=procs, mp_context=mp.get_context("fork")) as pool:
        for i in range(args.steps + args.warmup):
            t0 = time.perf_counter()
            secs = list(pool.map(_ref_worker, [(args.workload, i * procs + k) for k in range(procs)]))
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
                single.extend(secs)
            if len(times) >= 3 and time.perf_counter() - t_begin > args.ref_budget:
                break
    ms = 1e3 * float(np.mean(times))
    value = procs * 1e3 / ms
    base = {"kind": "port", "cores": procs, "value": value, "unit": "frames/s",
            "sample": "%d timed steps (of %d asked; %.0f s budget) of %d frames of %s, one per worker process, through oracle/dens_oracle.py "
                      "(numpy restatement of reference dens.py:277-321); %.2f s per frame inside a worker"
                      % (len(times), args.steps, args.ref_budget, procs, args.workload, float(np.mean(single)))}
    print(json.dumps({
        "impl": "reference", "metric": "trajectory frames/sec into 3D S(q)", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, wl["desc"]), "frames_per_step": procs},
        "cpu_baseline": base, "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": os.cpu_count()}))


# ----------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import mdsf_b200
    workloads = __import__("workloads")
    wl = workloads.get(args.workload)

    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dens, native = mdsf_b200.dens, mdsf_b200.native
    native.load()
    dens.PRINT_DETAILS = False

    F, K, W = args.frames_per_step, args.steps, args.warmup
    natoms = wl["base"].shape[0]
    box = wl["box"]
    tile = tuple(int(v) for v in args.tile.split("x"))
    eng, n, dr, nb = dens.make_engine(box, wl["typ"], wl["rad"], wl["ucell"], wl["sres"], np.float32, np.float32,
                                      device=local, batch_frames=F, fft_mode=args.fft, tile=tile)
    assert tuple(int(v) for v in n) == tuple(wl["grid"]), (n, wl["grid"])
    # frame pool: pinned host memory (e2e leg) and a device-resident copy (value leg); rank r starts
    # at a different offset so ranks do not process identical frames
    pool_n = max(args.pool, F)
    pool = native.pinned_empty((pool_n, natoms, 3), np.float32)
    workloads.jitter_frames(wl["base"], box, pool_n, wl["jitter"], wl["seed0"] + 7919 * rank, out=pool)
    dpool = torch.from_numpy(pool).to("cuda:%d" % local)
    scale = np.ones((F, 3))
    frame_bytes = natoms * 12
    sf_shape = (int(n[0]), int(n[1]), int(n[2]) // 2 + 1)
    red = torch.empty(sf_shape, dtype=torch.float64, device="cuda:%d" % local)
    sf_host = native.pinned_empty(sf_shape, np.float64)          # destination of the final S(q) read-out

    def step_device(i):
        s = (i * F) % (pool_n - F + 1)
        eng.push_frames_ptr(dpool.data_ptr() + s * frame_bytes, F, scale)

    def step_host(i):
        s = (i * F) % (pool_n - F + 1)
        eng.push_frames(pool[s:s + F], scale)

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def finish():
        """close a job: cross-GPU reduce of the partial S(q) (N>1) -- part of the timed region"""
        if world > 1:
            eng.sync()
            eng.export_sf_device(red.data_ptr())
            dist.reduce(red, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()

    # ---- value: device-resident inputs, CUDA events
    for i in range(W):
        step_device(i)
    finish()
    barrier()
    eng.reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.kernel_launches
    eng.enable_timing(True)
    barrier()
    t_host0 = time.perf_counter()
    eng.timer_start()
    for i in range(K):
        step_device(W + i)
    ms_dev = eng.timer_stop()
    finish()
    barrier()
    t_host = time.perf_counter() - t_host0
    stage, nbatch = eng.stage_ms()
    eng.enable_timing(False)
    launches = eng.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    # per-job time = device time of the K steps (+ the reduce for N>1, host-timed around synchronised work)
    job_ms = ms_dev if world == 1 else 1e3 * t_host
    t = torch.tensor([job_ms], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    job_ms = float(t.item())
    value = world * K * F / (job_ms * 1e-3)

    # ---- e2e: pinned host frames through the public call, H2D + final S(q) D2H inside the region
    eng.reset()
    for i in range(W):
        step_host(i)
    barrier()
    eng.reset()
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        step_host(W + i)
    finish()
    sf = eng.read_sf(out=sf_host) if rank == 0 or world == 1 else None
    if sf is None:
        eng.sync()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * K * F / float(t.item())
    sf_bytes = int(np.prod(sf_shape)) * 8

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_frame = workloads.algorithmic_bytes_per_frame(wl["grid"], natoms)
        # prep_bin runs on its own stream underneath the previous batch's FFT passes: its event span includes waiting
        kern = {k: v for k, v in stage.items() if k not in ("copy", "total", "prep_bin")}
        dom = max(kern, key=kern.get)
        dom_ms = kern[dom] / max(nbatch, 1)                      # average duration of one launch of that stage
        dom_gbs = alg_frame * F / (dom_ms * 1e-3) / 1e9
        # dram bytes per step (all kernels / the dominant one) from the committed ncu --set full capture
        traffic = dom_traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic_%s.json" % args.workload)
        if os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh)
            key = {"splat_zfft": "splat_zfft", "fft_y": "fft_y", "fft_x_accum": "fft_x_accum"}.get(dom, dom)
            traffic = 0.0
            for kname, kv in tj.get("kernels", {}).items():
                per_step = kv["dram_bytes_per_launch"] * F / kv["frames_per_launch"]
                traffic += per_step
                if key in kname:
                    dom_traffic = per_step
        step_gbs = alg_frame * value / world / 1e9
        out = {
            "metric": "trajectory frames/sec into 3D S(q)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": job_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, wl["desc"]), "grid": list(wl["grid"]), "atoms": natoms,
                       "frames_per_step": F, "pool_frames": pool_n, "fft": eng.fft_path, "splat": eng.splat_path, "Nborder": nb,
                       "geometry": eng.geometry,
                       "l2": "per-step working set (%.0f MB of pair volumes + accumulator) exceeds the 126 MB L2; no explicit flush"
                             % ((F // 2) * np.prod(wl["grid"]) * 16 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": F * frame_bytes,
                    "d2h_bytes_per_step": sf_bytes // K, "note": "pinned host frames -> Engine.push_frames; S(q) read once per job"},
            "gpu_launches": launches,
            # achieved = B_alg x frames/s over the WHOLE step (all four kernels of the path); the dominant
            # kernel's own figure (B_alg x F / its launch duration) is given beside it
            "roofline": {"bound": "hbm", "kernel": "whole step: prep+bin, splat_zfft, fft_y, fft_x_accum", "achieved": step_gbs,
                         "peak": peak, "unit": "GB/s", "frac": step_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_frame": alg_frame, "frames_per_launch": F,
                         "dominant_kernel": {"name": dom, "ms_per_launch": dom_ms, "achieved": dom_gbs, "frac": dom_gbs / peak,
                                             "traffic": dom_traffic}},
            "stage_ms_per_step": {k: v / max(nbatch, 1) for k, v in stage.items()},
            "clocks": clocks, "host_cores": os.cpu_count(),
        }
        if not args.no_cpu:
            fps, sec = cpu_port_frames_per_s(wl, args.cpu_frames)
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port",
                                   "sample": "%d frames of %s, numpy restatement of reference dens.py:277-321 (oracle/dens_oracle.py), "
                                             "%.2f s/frame" % (args.cpu_frames, args.workload, sec)}
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
