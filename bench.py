#!/usr/bin/env python
"""Benchmark of the hot path: 1920-wide frame pairs / second through the coarse-to-fine solver.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode fp32_redblack]

A "step" is one BATCH of B frame pairs per GPU (default B=24; each pair 1920x1080 RGB, alpha=0.012
ratio=0.75 minWidth=20 -> 15 levels, 7/1/30 iterations: BASELINE.json configs[2], the headline
single-GPU case), the B solves running concurrently on B streams of the GPU (pairs are independent;
the latency-bound coarse levels of one pair hide behind the bandwidth-bound fine levels of another).
One process per GPU; ranks shard pairs with no data-path collective ("scaling": "weak").
Rank 0 prints exactly one JSON line (see the task contract):
  value      pairs/s, inputs resident in HBM, K graph replays timed with CUDA events on the
             launching stream, max over ranks
  e2e        pairs/s through the public plan API with HOST (pinned) buffers: H2D + solve + D2H per step
  roofline   SOR kernel (k_sor_rb_tma) at pyramid level 0: algorithmic bytes (40 B per pixel-sweep in
             FP32) / CUDA-event time of its launches, against the measured HBM peak
  cpu_baseline  the unmodified reference (oracle/_ref, Serial build, 1 core) on a bounded sample
--impl reference times the reference's own OpenMP build on all host cores on the same workload
(bounded sample per step).  oracle/ is used here ONLY as that measured baseline, never by our arm.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
# before ANY CUDA context exists (torch included): one hardware work queue per in-flight pair stream
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ROOT)

H, W, CH = 1080, 1920, 3
PARAMS = dict(alpha=0.012, ratio=0.75, minWidth=20, nOuter=7, nInner=1, nSOR=30, colType=0)
SAMPLE_ROWS = 216          # CPU baselines run a 1920x216 band (1/5 of the pair) per step
WORKLOAD = "HoChiMinhTraffic_10FPS_1920 pair, 1920x1080 RGB, defaults alpha=0.012 ratio=0.75 minWidth=20 (15 levels) 7/1/30"


def load_frames():
    """Frames 1..3 of the reference's 1920-wide collection (byte copies under tests/golden/frames,
    decoded like Par/OpticalFlowCalculation.py:66-71); synthetic textured frames if absent."""
    try:
        from PIL import Image
        fr = [np.array(Image.open(os.path.join(ROOT, "tests", "golden", "frames", "hcm1920_%05d.jpg" % i))).astype(float) / 255.
              for i in (1, 2, 3)]
        return fr, "real: HoChiMinhTraffic_10FPS_1920 frames 1-3 (fixture copies), pairs (1,2),(2,3) alternating"
    except Exception:
        rng = np.random.default_rng(0)
        base = rng.random((H // 8 + 4, W // 8 + 4, CH))
        big = np.kron(base, np.ones((8, 8, 1)))
        fr = [np.ascontiguousarray(big[8 + 2 * i:8 + 2 * i + H, 8 + 3 * i:8 + 3 * i + W]) for i in range(3)]
        return fr, "synthetic: block-noise texture translated by (3,2) px per frame"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def pinned_like(lib, arr):
    n = arr.nbytes
    ptr = lib.pf_host_alloc(n)
    if not ptr:
        return np.ascontiguousarray(arr), False
    buf = (C.c_double * (n // 8)).from_address(ptr)
    out = np.frombuffer(buf, dtype=np.float64).reshape(arr.shape)
    out[...] = arr
    return out, True


def cpu_reference_sample(parallel, frames, reps=1):
    """Runs the UNMODIFIED reference on a 1920 x SAMPLE_ROWS band; returns (pairs/s equivalent, cores, text)."""
    from oracle import ref
    if not ref.available():
        return None
    a = np.ascontiguousarray(frames[0][432:432 + SAMPLE_ROWS]); b = np.ascontiguousarray(frames[1][432:432 + SAMPLE_ROWS])
    r = ref.parallel() if parallel else ref.serial()
    cores = (os.cpu_count() or 1) if parallel else 1
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        r.coarse2fine_flow_levels(a, b, 15, cores)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    frac = SAMPLE_ROWS / float(H)
    return frac / best, cores, "rows 432..%d of frames 1-2 (1920x%d band, %.0f%% of the pair's pixels, 15 levels); pairs/s = %.2f / seconds" % (432 + SAMPLE_ROWS, SAMPLE_ROWS, 100 * frac, frac), best


def dist_setup(n):
    if n <= 1:
        return None, 0, 0
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return dist, rank, local


def dist_max(dist, local, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda:%d" % local)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist):
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames, data = load_frames()
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (reference tree absent at build time)"}))
        return
    for _ in range(args.warmup):
        cpu_reference_sample(True, frames)
    t0 = time.perf_counter()
    vals = [cpu_reference_sample(True, frames) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    frac = SAMPLE_ROWS / float(H)
    value = args.steps * frac / dt
    cores = vals[0][1]
    line = {"impl": "reference", "metric": "frame_pairs_per_sec_1920w", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps / frac, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": data,
            "config": {"workload": WORKLOAD, "mode": "reference OpenMP build (Code/Parallel), nCores=%d" % cores},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "reference", "sample": vals[0][2]},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    dist, rank, local = dist_setup(args.gpus)
    import pyflow
    from papteam_opticalflow_b200 import _lib
    lib = _lib.lib()
    if lib.pf_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device and no CPU fallback")
    frames, data = load_frames()
    B = max(1, args.batch)
    plans = [pyflow.FlowPlan(H, W, CH, mode=args.mode, device=local, **PARAMS) for _ in range(B)]
    plan = plans[0]
    pin = [pinned_like(lib, f) for f in frames]
    fr = [p[0] for p in pin]
    pinned = all(p[1] for p in pin)
    pairs = [(fr[0], fr[1]), (fr[1], fr[2])]

    # ---- device-resident throughput: K steps of B concurrent graph replays, one pair resident per
    #      plan; CUDA events on the launching stream around the whole region (pf_multi_solve) ----
    for i, p in enumerate(plans):
        p.upload(*pairs[i % 2])
    pyflow.multi_solve(plans, max(1, args.warmup))
    sampler = ClockSampler(local); sampler.start()
    barrier(dist)
    ms = pyflow.multi_solve(plans, args.steps)
    barrier(dist)
    ms = dist_max(dist, local, ms)
    value = args.gpus * args.steps * B / (ms / 1000.0)

    # ---- end to end through the public batch API with HOST buffers: every pair is copied
    #      host->device, solved and copied back inside the timed region.  All K steps (K*B pairs) go
    #      through ONE pf_batch_flow call so that the copy legs of some pairs overlap the solves of
    #      others instead of every step starting with B simultaneous uploads. ----
    os.environ.setdefault("PF_BATCH_STREAMS", str(min(B, 8)))
    ring = B + 8                                        # output slots, reused cyclically (more than the workers in flight)
    host_outs = [tuple(pinned_like(lib, np.zeros(s))[0] for s in ((H, W), (H, W), (H, W, CH))) for _ in range(ring)]
    def e2e_run(nsteps):
        n = nsteps * B
        pyflow.coarse2fine_flow_batch([pairs[i % 2] for i in range(n)], PARAMS["alpha"], PARAMS["ratio"], PARAMS["minWidth"],
                                      PARAMS["nOuter"], PARAMS["nInner"], PARAMS["nSOR"], PARAMS["colType"], mode=args.mode,
                                      devices=[local], outs=[host_outs[i % ring] for i in range(n)])
    e2e_run(max(1, min(2, args.warmup)))
    barrier(dist)
    t0 = time.perf_counter()
    e2e_run(args.steps)
    e2e_s = time.perf_counter() - t0
    barrier(dist)
    e2e_s = dist_max(dist, local, e2e_s)
    e2e = args.gpus * args.steps * B / e2e_s

    # ---- the same batch call on PAGEABLE numpy arrays (what a Python caller of the reference has): the library
    #      moves them through its multi-threaded pinned-slot stager (csrc/staging.hpp).  Extra key. ----
    page_outs = [tuple(np.zeros(s) for s in ((H, W), (H, W), (H, W, CH))) for _ in range(ring)]
    for o in page_outs:
        for a_ in o:
            a_.fill(0)          # touch the pages: a caller's buffers are normally mapped already
    def page_run(nsteps):
        n = nsteps * B
        pyflow.coarse2fine_flow_batch([(frames[i % 2], frames[i % 2 + 1]) for i in range(n)], PARAMS["alpha"], PARAMS["ratio"],
                                      PARAMS["minWidth"], PARAMS["nOuter"], PARAMS["nInner"], PARAMS["nSOR"], PARAMS["colType"],
                                      mode=args.mode, devices=[local], outs=[page_outs[i % ring] for i in range(n)])
    page_run(1)
    barrier(dist)
    t0 = time.perf_counter()
    page_run(max(1, args.steps // 2))
    page_s = time.perf_counter() - t0
    barrier(dist)
    page_s = dist_max(dist, local, page_s)
    e2e_page = args.gpus * max(1, args.steps // 2) * B / page_s

    # ---- sequence mode (reported beside the headline, not instead of it): K*B+1 uint8
    #      frames in, K*B float32 flows out, every frame's pyramid built once, conversion from uint8 on the device ----
    def pinned_raw(shape, dtype):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = lib.pf_host_alloc(nbytes)
        if not ptr:
            return np.zeros(shape, dtype)
        return np.frombuffer((C.c_ubyte * nbytes).from_address(ptr), dtype=dtype).reshape(shape)
    u8 = []
    for f in frames:
        a = pinned_raw(f.shape, np.uint8)
        a[...] = np.rint(f * 255.0).astype(np.uint8)
        u8.append(a)
    seq_outs = [pinned_raw((H, W, 2), np.float32) for _ in range(ring)]
    def seq_run(nsteps):
        n = nsteps * B
        pyflow.sequence_flow([u8[i % 3] for i in range(n + 1)], PARAMS["alpha"], PARAMS["ratio"], PARAMS["minWidth"], PARAMS["nOuter"],
                             PARAMS["nInner"], PARAMS["nSOR"], PARAMS["colType"], mode=args.mode, devices=[local],
                             outs=[seq_outs[i % ring] for i in range(n)])
    seq_run(max(1, min(2, args.warmup)))
    barrier(dist)
    t0 = time.perf_counter()
    seq_run(args.steps)
    seq_s = time.perf_counter() - t0
    barrier(dist)
    seq_s = dist_max(dist, local, seq_s)
    clocks = sampler.summary()
    seq = args.gpus * args.steps * B / seq_s

    line = None
    if rank == 0:
        # ---- the drop-in call itself: pyflow.coarse2fine_flow on plain (pageable) numpy arrays, fresh outputs, one
        #      pair at a time -- what a caller of the reference's module sees (extra key, not the headline) ----
        one_shot = []
        for _ in range(4):
            t0 = time.perf_counter()
            pyflow.coarse2fine_flow(frames[0], frames[1], PARAMS["alpha"], PARAMS["ratio"], PARAMS["minWidth"], PARAMS["nOuter"],
                                    PARAMS["nInner"], PARAMS["nSOR"], PARAMS["colType"], mode=args.mode, device=local)
            one_shot.append(1000 * (time.perf_counter() - t0))
        one_shot_ms = float(np.median(one_shot[1:]))
        # ---- per-phase attribution + SOR roofline from one eager, event-instrumented solve ----
        single_ms = plan.solve(3) / 3
        plan.profile()
        tp, cnt = plan.profile()
        peak, peak_src = hbm_peak()
        word = 4 if args.mode.startswith("fp32") else 8
        sor_ms_l0, sor_launch_l0, ps_l0 = cnt[3], cnt[4], cnt[5]
        bytes_per_launch = ps_l0 * 10 * word / max(1.0, sor_launch_l0)
        achieved = (ps_l0 * 10 * word / 1e9) / (sor_ms_l0 / 1e3) if sor_ms_l0 > 0 else 0.0
        sor_all = (cnt[2] * 10 * word / 1e9) / (tp[_lib.T_PHASE5] / 1e3) if tp[_lib.T_PHASE5] > 0 else 0.0
        traffic = None
        tf = os.path.join(ROOT, "profiles", "sor_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        if args.gpus == 1 and not args.no_cpu:
            try:
                r = cpu_reference_sample(False, frames)
                if r:
                    cpu = {"value": r[0], "unit": "pairs/s", "cores": r[1], "kind": "reference", "sample": r[2], "seconds": r[3]}
            except Exception as e:   # the baseline must never break the GPU line
                cpu = {"value": None, "unit": "pairs/s", "cores": 1, "kind": "reference", "sample": "failed: %r" % (e,)}
        phases = {k: round(float(tp[i]), 3) for i, k in enumerate(
            ["total", "pyramid", "features_upsample_warp", "getDxs", "phi", "psi(fused)", "assemble", "sor", "update_warp", "bicubic_export"])}
        line = {
            "metric": "frame_pairs_per_sec_1920w", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_pair": ms / args.steps / B,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.mode.startswith("fp32") else "f64", "data": data,
            "config": {"workload": WORKLOAD, "mode": args.mode, "pairs_per_gpu_per_step": B,
                       "concurrency": "%d pairs in flight per GPU, one CUDA stream + graph each" % B,
                       "l2": "per-solve working set ~0.5 GB of planes > 126 MB L2 (no explicit flush)",
                       "host_buffers": "pinned" if pinned else "pageable"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(B * 2 * H * W * CH * 8),
                    "d2h_bytes_per_step": int(B * (2 * H * W + H * W * CH) * 8), "ms_per_step": 1000 * e2e_s / args.steps,
                    "api": "pyflow.coarse2fine_flow_batch -> pf_batch_flow (host float64 HWC in, host float64 out)"},
            "e2e_pageable": {"value": e2e_page, "unit": "pairs/s",
                             "api": "the same batch call with plain (pageable) numpy arrays in and out, staged by the library"},
            "e2e_sequence": {"value": seq, "unit": "pairs/s", "h2d_bytes_per_step": B * H * W * CH, "d2h_bytes_per_step": B * H * W * 8,
                             "api": "pyflow.sequence_flow -> pf_sequence_flow_u8 (host uint8 frames in, host float32 (u,v) out; consecutive "
                                    "pairs share a frame, its pyramid is built once)"},
            "gpu_launches": int(cnt[0]) * args.steps * B,
            "single_pair_latency_ms": single_ms,
            "one_shot_call_ms": one_shot_ms,
            "roofline": {"bound": "hbm", "kernel": "k_sor_rb_tma (level 0, 1920x1080)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "launches_per_solve_level0": int(sor_launch_l0),
                         "avg_launch_ms": sor_ms_l0 / max(1.0, sor_launch_l0),
                         "sor_all_levels_GBps": sor_all, "sor_all_levels_frac": sor_all / peak,
                         "timing": "CUDA events on the launching stream around the SOR launches of one eager solve"},
            "cpu_baseline": cpu,
            "phases_ms": phases,
            "launches_per_solve": int(cnt[0]),
        }
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=24, help="frame pairs in flight per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fp32_redblack", choices=["fp32_redblack", "fp64_wavefront", "fp64_redblack", "fp32_wavefront"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
