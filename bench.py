#!/usr/bin/env python
"""bench.py -- frames/s of the structure-factor hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

A step is one pass of the whole per-frame path (rescale/wrap/cell index, deterministic Gaussian
splat, 3-D FFT, |F|^2 accumulation) over one batch of ``--frames-per-step`` synthetic frames.
The default workload is BASELINE.json configs[2] (c3: 334 080 atoms, 512^3 grid), the configuration
north_star's roofline target is quoted on; c2 (256^3) and a condensed-phase-density variant of c3
("c3d") are measured beside it under ``other_workloads`` (``--no-extra`` skips them).
* ``value``  : frames/s with the frame pool already resident in HBM, timed with CUDA events.
* ``e2e``    : frames/s through the C ABI (Engine.push_frames, the call dens.compute_sf makes) from
               pinned HOST memory, H2D copies and the final S(q) device->host read inside the timed
               region.
* ``roofline``: SURVEY 8(d) algorithmic bytes x frames/s of the whole step vs the measured HBM peak; per kernel,
               the bytes ncu saw it move (profiles/r02_traffic_<workload>.json) over its CUDA-event duration.
* ``cpu_baseline``: the UNMODIFIED reference (oracle/_ref/dens.py, kind "reference") on one frame of the same
               workload, one host core (the reference is a serial loop); the numpy port (kind "port") when the
               reference copy is not on the box.
``--impl reference`` times that same reference on the host cores (one frame per worker process and step).
Multi-GPU (torchrun, one rank per GPU): frames shard across ranks with no data-path collective
(weak scaling: every rank runs the same per-GPU work) and one NCCL reduce of the partial S(q)
closes the timed region; the time is the max over ranks.  After the timed region every rank pushes
the SAME frames and rank 0 checks the NCCL-reduced S(q) against world x its own partial.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_FRAMES = {"c1": 64, "c2": 64, "c3": 16, "c3d": 2, "c4": 4, "c5": 2, "tiny": 8}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--frames-per-step", type=int, default=0, help="0 = the workload's default")
    ap.add_argument("--pool", type=int, default=0, help="distinct synthetic frames cycled through (0 = 2 steps' worth)")
    ap.add_argument("--cpu-frames", type=int, default=1, help="frames of the CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the c2 / c3d side measurements")
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm: worker processes (0 = as many as cores and memory allow)")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="reference arm: stop timing after this many seconds")
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


def make_pool(wl, nframes, seed_shift=0):
    from importlib import import_module
    workloads = import_module("workloads")
    return workloads.jitter_frames(wl["base"], wl["box"], nframes, wl["jitter"], wl["seed0"] + seed_shift)


# ----------------------------------------------------------------------------- CPU arm
def cpu_frames_per_s(wl, nframes, seed_shift=0):
    """Frame loop of the reference on one core: the unmodified reference file when oracle/_ref/ travelled to this box
    (kind "reference"; timed from the start of its frame loop to the end of the last rfftn, its plot lattices and npz
    write excluded like on the GPU side), else the numpy restatement in oracle/dens_oracle.py (kind "port")."""
    coords = make_pool(wl, nframes, seed_shift)
    dims = np.repeat(wl["box"][None, :], nframes, axis=0)
    from oracle import ref_runner
    if ref_runner.available():
        with contextlib.redirect_stdout(io.StringIO()):
            out = ref_runner.run(coords, dims, list(wl["typ"]), wl["rad"], wl["ucell"], wl["sres"], keep_output=False)
        per_frame = out["loop_s"] / out["frames"]
        return 1.0 / per_frame, per_frame, "reference"
    from oracle import dens_oracle as orc
    stamps = []
    t0 = time.perf_counter()
    orc.structure_factor(coords, dims, wl["typ"], wl["rad"], wl["ucell"], wl["sres"], frame_callback=lambda t: stamps.append(time.perf_counter()))
    per_frame = (stamps[-1] - t0) / nframes
    return 1.0 / per_frame, per_frame, "port"


_REF_WL = None


def _ref_worker(args):
    """One frame of the workload through the reference in a worker process; returns (frame-loop seconds, kind)."""
    name, seed = args
    global _REF_WL
    from importlib import import_module
    workloads = import_module("workloads")
    if _REF_WL is None or _REF_WL[0] != name:
        _REF_WL = (name, workloads.get(name))
    fps, sec, kind = cpu_frames_per_s(_REF_WL[1], 1, seed)
    return sec, kind


def reference_memory_gb(wl):
    """Peak host memory of one reference worker: dens.py:237-256 allocates the padded density and a 4-channel coordinate
    lattice of the padded grid; the frame loop adds the folded density, the spectrum, |F|^2 and S(q); after the loop
    dens.py:323-344 builds sfplt, kgrid and kgridplt (another 11 grid-sized arrays) before anything is freed."""
    from oracle import dens_oracle as orc
    dr = np.asarray(wl["box"], dtype=np.float64) / np.asarray(wl["grid"], dtype=np.float64)
    nb = orc.border_cells(orc.half_widths(wl["rad"], dr, set(wl["typ"])))
    n = np.array(wl["grid"], dtype=np.float64) + 2 * nb
    return float(np.prod(n) * 8 * 5 + np.prod(wl["grid"]) * 8 * 15) / 1e9


def run_reference_arm(args, wl, rank):
    """The reference's own CPU implementation of the path on the host cores: the reference is one serial loop and frames
    are independent, so one step = one frame per worker process (as many workers as cores AND host memory allow), timed
    by wall clock."""
    if rank != 0:
        return
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    try:
        with open("/proc/meminfo") as fh:
            avail = [int(l.split()[1]) for l in fh if l.startswith("MemAvailable")][0] / 1e6
    except (OSError, IndexError):
        avail = 64.0
    procs = args.ref_procs or max(1, min(cores, int(0.6 * avail / reference_memory_gb(wl))))
    t_begin = time.perf_counter()
    times, single, kind = [], [], "port"
    with ProcessPoolExecutor(max_workers=procs, mp_context=mp.get_context("fork")) as pool:
        for i in range(args.steps + args.warmup):
            t0 = time.perf_counter()
            res = list(pool.map(_ref_worker, [(args.workload, 100000 + i * procs + k) for k in range(procs)]))
            dt = time.perf_counter() - t0
            kind = res[0][1]
            if i >= args.warmup or time.perf_counter() - t_begin > args.ref_budget:
                times.append(dt)
                single.extend(r[0] for r in res)
            if times and time.perf_counter() - t_begin > args.ref_budget:
                break
    ms = 1e3 * float(np.mean(times))
    value = procs * 1e3 / ms
    what = ("the unmodified reference dens.py (oracle/_ref, frame loop dens.py:277-321)" if kind == "reference"
            else "oracle/dens_oracle.py (numpy restatement of reference dens.py:277-321)")
    base = {"kind": kind, "cores": procs, "value": value, "unit": "frames/s",
            "sample": "%d timed steps (of %d asked; %.0f s budget) of %d frames of %s, one per worker process (%d host cores, "
                      "%.0f GB per worker), through %s; %.2f s per frame inside a worker"
                      % (len(times), args.steps, args.ref_budget, procs, args.workload, cores, reference_memory_gb(wl), what, float(np.mean(single)))}
    print(json.dumps({
        "impl": "reference", "metric": "trajectory frames/sec into 3D S(q)", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, wl["desc"]), "frames_per_step": procs},
        "cpu_baseline": base, "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": cores}))


# ----------------------------------------------------------------------------- GPU arm
def load_traffic(workload):
    """{kernel-name fragment: dram bytes per FRAME} from the committed ncu --set full capture of this round."""
    path = os.path.join(ROOT, "profiles", "r02_traffic_%s.json" % workload)
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        tj = json.load(fh)
    return {k: v["dram_bytes_per_launch"] / v["frames_per_launch"] for k, v in tj.get("kernels", {}).items()}


def measure(mdsf_b200, workloads, name, wl, args, local, rank, world, dist, torch, F, K, W, full):
    """value / e2e / stage times of one workload.  ``full``: also the clocks sampler, e2e leg and NCCL check."""
    dens, native = mdsf_b200.dens, mdsf_b200.native
    natoms = wl["base"].shape[0]
    box = wl["box"]
    tile = tuple(int(v) for v in args.tile.split("x"))
    eng, n, dr, nb = dens.make_engine(box, wl["typ"], wl["rad"], wl["ucell"], wl["sres"], np.float32, np.float32,
                                      device=local, batch_frames=F, fft_mode=args.fft, tile=tile)
    assert tuple(int(v) for v in n) == tuple(wl["grid"]), (n, wl["grid"])
    F = eng.batch_frames                      # the engine may shrink the batch (pair-list memory)
    pool_n = max(args.pool or 2 * F, F)
    pool = native.pinned_empty((pool_n, natoms, 3), np.float32)
    workloads.jitter_frames(wl["base"], box, pool_n, wl["jitter"], wl["seed0"] + 7919 * rank, out=pool)
    dpool = torch.from_numpy(pool).to("cuda:%d" % local)
    scale = np.ones((F, 3))
    frame_bytes = natoms * 12
    sf_shape = (int(n[0]), int(n[1]), int(n[2]) // 2 + 1)
    red = torch.empty(sf_shape, dtype=torch.float64, device="cuda:%d" % local)
    sf_host = native.pinned_empty(sf_shape, np.float64)

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
    if rank == 0 and full:
        sampler.start()
    launches0 = eng.kernel_launches
    eng.enable_timing(True)
    barrier()
    t_host0 = time.perf_counter()
    eng.timer_start()
    for i in range(K):
        step_device(W + i)
    ms_dev = eng.timer_stop()
    t_red0 = time.perf_counter()
    finish()
    barrier()
    reduce_ms = 1e3 * (time.perf_counter() - t_red0) if world > 1 else 0.0
    t_host = time.perf_counter() - t_host0
    stage, nbatch = eng.stage_ms()
    eng.enable_timing(False)
    launches = eng.kernel_launches - launches0
    clocks = sampler.stop() if (rank == 0 and full) else None
    job_ms = ms_dev if world == 1 else 1e3 * t_host
    t = torch.tensor([job_ms], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    job_ms = float(t.item())
    res = {"value": world * K * F / (job_ms * 1e-3), "job_ms": job_ms, "F": F, "K": K, "launches": launches, "clocks": clocks,
           "stage": {k: v / max(nbatch, 1) for k, v in stage.items()}, "n": n, "nb": nb, "natoms": natoms, "pool_n": pool_n,
           "fft": eng.fft_path, "splat": eng.splat_path, "geometry": eng.geometry, "reduce_ms": reduce_ms}

    if full:
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
        res["e2e_value"] = world * K * F / float(t.item())
        res["sf_bytes"] = int(np.prod(sf_shape)) * 8
        res["frame_bytes"] = frame_bytes
        if world > 1:
            # ---- NCCL check (outside every timed region): all ranks push the SAME frames; the reduced S(q) on rank 0
            # must equal world x rank 0's own partial (fp64 sum of identical terms: exact up to the reduce order)
            eng.reset()
            same = native.pinned_empty((F, natoms, 3), np.float32)
            workloads.jitter_frames(wl["base"], box, F, wl["jitter"], wl["seed0"] + 424242, out=same)
            eng.push_frames(same, scale)
            eng.sync()
            eng.export_sf_device(red.data_ptr())
            mine = red.clone()
            dist.reduce(red, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            if rank == 0:
                err = float(((red - world * mine).abs().max() / mine.abs().max()).item())
                res["nccl_check"] = {"max_err_normalised": err, "ok": bool(err <= 1e-14), "ranks": world}
                assert err <= 1e-14, "NCCL-reduced S(q) differs from world x partial: %g" % err
    eng.close()
    del dpool, red
    torch.cuda.empty_cache()
    return res


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
    mdsf_b200.native.load()
    mdsf_b200.dens.PRINT_DETAILS = False

    F = args.frames_per_step or DEFAULT_FRAMES.get(args.workload, 8)
    K, W = args.steps, args.warmup
    r = measure(mdsf_b200, workloads, args.workload, wl, args, local, rank, world, dist, torch, F, K, W, True)
    F = r["F"]

    extra = {}
    if not args.no_extra and world == 1 and args.workload == "c3":
        for name in ("c2", "c3d"):
            try:
                w2 = workloads.get(name)
                k2 = max(3, min(K, 6))
                r2 = measure(mdsf_b200, workloads, name, w2, args, local, rank, world, dist, torch, DEFAULT_FRAMES[name], k2, 3, False)
                alg2 = workloads.algorithmic_bytes_per_frame(w2["grid"], r2["natoms"])
                peak, _ = measured_peak()
                extra[name] = {"workload": "%s: %s" % (name, w2["desc"]), "value": r2["value"], "unit": "frames/s", "steps": k2,
                               "frames_per_step": r2["F"], "roofline_frac": alg2 * r2["value"] / 1e9 / peak,
                               "stage_ms_per_step": r2["stage"], "splat": r2["splat"], "geometry": r2["geometry"]}
            except Exception as exc:          # a side measurement must not lose the headline
                extra[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        peak, peak_src = measured_peak()
        natoms = r["natoms"]
        alg_frame = workloads.algorithmic_bytes_per_frame(wl["grid"], natoms)
        stage = r["stage"]
        kern = {k: v for k, v in stage.items() if k not in ("copy", "total", "prep_bin")}
        dom = max(kern, key=kern.get)
        traffic_frame = load_traffic(args.workload)
        traffic = None
        per_kernel = {}
        if traffic_frame:
            traffic = sum(traffic_frame.values()) * F
            for kname, ms in kern.items():
                b = sum(v for k, v in traffic_frame.items() if kname in k or k in kname)
                if b:
                    per_kernel[kname] = {"ms_per_launch": ms, "dram_bytes_per_launch": b * F,
                                         "achieved": b * F / (ms * 1e-3) / 1e9, "frac": b * F / (ms * 1e-3) / 1e9 / peak}
        step_gbs = alg_frame * r["value"] / world / 1e9
        out = {
            "metric": "trajectory frames/sec into 3D S(q)", "value": r["value"], "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": r["job_ms"] / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, wl["desc"]), "grid": list(wl["grid"]), "atoms": natoms,
                       "frames_per_step": F, "pool_frames": r["pool_n"], "fft": r["fft"], "splat": r["splat"], "Nborder": r["nb"],
                       "geometry": r["geometry"],
                       "l2": "per-step working set (%.0f MB of pair volumes + accumulator) exceeds the 126 MB L2; no explicit flush"
                             % ((F // 2) * np.prod(wl["grid"]) * 16 / 1e6)},
            "e2e": {"value": r["e2e_value"], "unit": "frames/s", "h2d_bytes_per_step": F * r["frame_bytes"],
                    "d2h_bytes_per_step": r["sf_bytes"] // K, "note": "pinned host frames -> Engine.push_frames; S(q) read once per job"},
            "gpu_launches": r["launches"],
            # achieved = B_alg x frames/s over the WHOLE step (every kernel of the path).  dominant_kernel is
            # the longest stage with ITS OWN bytes: what ncu saw it move (committed capture) over its CUDA-event duration.
            "roofline": {"bound": "hbm", "kernel": "whole step: prep+bin, splat_zfft, fft_y, fft_x_accum", "achieved": step_gbs,
                         "peak": peak, "unit": "GB/s", "frac": step_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_frame": alg_frame, "frames_per_launch": F,
                         "dominant_kernel": dict({"name": dom}, **per_kernel.get(dom, {"ms_per_launch": kern[dom], "dram_bytes_per_launch": None,
                                                                                    "achieved": None, "frac": None})),
                         "kernels": per_kernel},
            "stage_ms_per_step": stage,
            "clocks": r["clocks"], "host_cores": os.cpu_count(),
        }
        if world > 1:
            out["reduce_ms"] = r["reduce_ms"]
            out["nccl_check"] = r.get("nccl_check")
        if extra:
            out["other_workloads"] = extra
        if not args.no_cpu and world == 1:
            fps, sec, kind = cpu_frames_per_s(wl, args.cpu_frames)
            what = ("the unmodified reference dens.py (oracle/_ref), frame loop dens.py:277-321" if kind == "reference"
                    else "numpy restatement of reference dens.py:277-321 (oracle/dens_oracle.py)")
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": 1, "kind": kind,
                                   "sample": "%d frame(s) of %s, %s, %.2f s/frame" % (args.cpu_frames, args.workload, what, sec)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
