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
  cpu_baseline  the unmodified reference (oracle/_ref, Serial build, 1 core): ONE FULL 1920x1080 pair
--impl reference times the reference's own OpenMP build on all host cores on the same workload, one FULL
pair per step (no extrapolation).  oracle/ is used here ONLY as that measured baseline, never by our arm.
Extra keys: e2e_flow_only (warpI2 = NULL), e2e_pageable, e2e_sequence (uint8 in / float32 out), per-leg
milliseconds of the batch calls, config2_fp64_wavefront (960-wide pair, parity mode), config5_rowband (N > 1:
one 4K pair split over all GPUs), single_pair_latency_ms, one_shot_call_ms, phases_ms, parity (full-frame EPE of
the benchmarked mode against the parity mode on the sequence pairs at hand, measured in this run, with
parity_note), hybrid (the fp32_hybrid mode: latency and resident throughput).
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
WORKLOAD = "HoChiMinhTraffic_10FPS_1920 pair, 1920x1080 RGB, defaults alpha=0.012 ratio=0.75 minWidth=20 (15 levels) 7/1/30"
L2_NOTE = "no explicit flush: per-solve working set ~0.5 GB of planes per pair x 24 pairs in flight >> 126 MB L2"
REF_COLLECTION = "/root/reference/images_New/HoChiMinhTraffic_10FPS_1920"


class Frames(object):
    """The 1920-wide collection BASELINE configs 3/4 name.  All 102 frames when the reference tree is on this machine
    (or PF_BENCH_FRAMES points at a directory of frame_%05d.jpg files); otherwise the seven fixture copies under
    tests/golden/frames (frames 1, 2, 3, 50, 51, 101, 102 -- the pairs SURVEY.md 8d spot-checks); synthetic texture if
    neither can be decoded.  Decoding follows the reference driver (Par/OpticalFlowCalculation.py:66-71): PIL -> uint8
    RGB -> astype(float) / 255.  Frames are decoded lazily: a rank only touches the frames of its own pairs."""

    def __init__(self):
        self.cache = {}
        self.paths = {}
        self.synthetic = None
        src = os.environ.get("PF_BENCH_FRAMES", REF_COLLECTION)
        try:
            from PIL import Image  # noqa: F401
            for i in range(1, 103):
                q = os.path.join(src, "frame_%05d.jpg" % i)
                if os.path.exists(q):
                    self.paths[i] = q
            if len(self.paths) >= 2:
                self.data = "real: %d frames of HoChiMinhTraffic_10FPS_1920 read from %s" % (len(self.paths), src)
            else:
                self.paths = {}
                for i in range(1, 103):
                    q = os.path.join(ROOT, "tests", "golden", "frames", "hcm1920_%05d.jpg" % i)
                    if os.path.exists(q):
                        self.paths[i] = q
                if len(self.paths) < 2:
                    raise IOError("no fixture frames")
                self.data = ("real: HoChiMinhTraffic_10FPS_1920 fixture frames %s (byte copies under tests/golden/frames; the full "
                             "102-frame collection is not on this machine)" % ",".join(str(i) for i in sorted(self.paths)))
        except Exception:
            rng = np.random.default_rng(0)
            base = rng.random((H // 8 + 4, W // 8 + 4, CH))
            big = np.kron(base, np.ones((8, 8, 1)))
            self.synthetic = [np.ascontiguousarray(big[8 + 2 * i:8 + 2 * i + H, 8 + 3 * i:8 + 3 * i + W]) for i in range(3)]
            self.paths = {1: None, 2: None, 3: None}
            self.data = "synthetic: block-noise texture translated by (3,2) px per frame"

    def indices(self):
        return sorted(self.paths)

    def pairs(self):
        """(t, t+1) for every frame t whose successor exists: the pairing rule of
        Par/InputCreation/TestImagePairGenerator.py:151-171 (101 pairs for the full collection)."""
        idx = self.indices()
        return [(a, b) for a, b in zip(idx[:-1], idx[1:]) if b == a + 1]

    def u8(self, i):
        from PIL import Image
        if self.synthetic is not None:
            return np.rint(self.synthetic[i - 1] * 255.0).astype(np.uint8)
        return np.ascontiguousarray(np.array(Image.open(self.paths[i])))

    def f64(self, i):
        if i not in self.cache:
            self.cache[i] = self.synthetic[i - 1] if self.synthetic is not None else self.u8(i).astype(float) / 255.
        return self.cache[i]


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


def cpu_reference_pair(parallel, im1, im2):
    """ONE FULL 1920x1080 pair through the UNMODIFIED reference (oracle/_ref): Code/Parallel on all host cores this
    process may use (P/OpticalFlow.cpp:735, nCores = that count; racy, timing only) or Code/Serial on one core.
    Returns (seconds, cores)."""
    from oracle import ref
    r = ref.parallel() if parallel else ref.serial()
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    cores = cores if parallel else 1
    t = time.perf_counter()
    r.coarse2fine_flow_levels(im1, im2, 15, cores)
    return time.perf_counter() - t, cores


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
    from papteam_opticalflow_b200.shard import max_over_ranks
    return max_over_ranks(dist, x, device=("cuda:%d" % local) if dist is not None else None)


def barrier(dist):
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


def node_dma_ceiling(dist, local, n_gpus):
    """What the NODE can move between pinned host memory and its GPUs when every rank copies at once, in both directions,
    with nothing but cudaMemcpyAsync (torch pinned tensors; the library is not involved): the ceiling of the float64
    end-to-end leg at N > 1 (182 MB of DMA per pair).  Measured 8 x B200, KVM guest, one visible NUMA node: 187 GB/s
    host->device alone, 95 GB/s device->host alone, 128 GB/s with both directions busy (tools/pcie_ceiling.py)."""
    import torch
    nb, reps = 48 << 20, 12
    failed = 0.0
    try:
        hs = [torch.empty(nb, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
        ds = [torch.empty(nb, dtype=torch.uint8, device="cuda:%d" % local) for _ in range(4)]
        s_in, s_out = torch.cuda.Stream(device=local), torch.cuda.Stream(device=local)
    except Exception:
        failed = 1.0
    if dist_max(dist, local, failed) > 0:      # every rank takes the same branch: the loops below hold barriers
        return None
    def run(k):
        barrier(dist)
        t0 = time.perf_counter()
        for i in range(k):
            with torch.cuda.stream(s_in):
                ds[i % 2].copy_(hs[i % 2], non_blocking=True)
            with torch.cuda.stream(s_out):
                hs[2 + i % 2].copy_(ds[2 + i % 2], non_blocking=True)
        torch.cuda.synchronize()
        return dist_max(dist, local, time.perf_counter() - t0)
    run(2)
    dt = run(reps)
    return n_gpus * reps * 2 * nb / dt / 1e9


def run_reference(args):
    """The reference arm: the reference's own OpenMP implementation of the path on the box's host cores, one FULL
    1920x1080 pair of the same workload per step (no extrapolation).  Rank 0 alone works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = Frames()
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (reference tree absent at build time)"}))
        return
    pairs = frames.pairs()
    cores = 1
    for k in range(args.warmup):
        a, b = pairs[k % len(pairs)]
        _, cores = cpu_reference_pair(True, frames.f64(a), frames.f64(b))
    t0 = time.perf_counter()
    for k in range(args.steps):
        a, b = pairs[k % len(pairs)]
        _, cores = cpu_reference_pair(True, frames.f64(a), frames.f64(b))
    dt = time.perf_counter() - t0
    value = args.steps / dt
    sample = "one full 1920x1080 pair per step (15 levels), pairs %s in turn" % ",".join("%d-%d" % p for p in pairs[:min(len(pairs), args.steps)])
    line = {"impl": "reference", "metric": "frame_pairs_per_sec_1920w", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": frames.data,
            "config": {"workload": WORKLOAD, "l2": L2_NOTE},
            "setup": {"implementation": "reference OpenMP build (Code/Parallel, unmodified, racy for nCores > 1), nCores=%d" % cores,
                      "pairs_per_step": 1},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def pinned_raw(lib, shape, dtype):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib.pf_host_alloc(nbytes)
    if not ptr:
        return np.zeros(shape, dtype), False
    return np.frombuffer((C.c_ubyte * nbytes).from_address(ptr), dtype=dtype).reshape(shape), True


def config2_parity_mode(pyflow, local):
    """BASELINE configs[1]: the 960-wide pair in the FP64 lexicographic mode (the mode that matches the reference to
    1e-6): device-resident ms per pair, median of 3 solves after one warm-up."""
    try:
        from PIL import Image
        fr = [np.array(Image.open(os.path.join(ROOT, "tests", "golden", "frames", "hcm960_%05d.jpg" % i))).astype(float) / 255. for i in (1, 2)]
    except Exception:
        return None
    h, w, c = fr[0].shape
    plan = pyflow.FlowPlan(h, w, c, mode="fp64_wavefront", device=local, tuning="latency", **PARAMS)
    plan.upload(fr[0], fr[1])
    plan.solve(1)
    ms = sorted(plan.solve(1) for _ in range(3))[1]
    _, cnt = plan.profile()
    plan.close()
    return {"workload": "HoChiMinhTraffic_10FPS_960 pair, 960x540 RGB, defaults, 13 levels (BASELINE configs[1])", "mode": "fp64_wavefront",
            "ms_per_pair": ms, "pairs_per_sec": 1000.0 / ms, "launches_per_solve": int(cnt[0]),
            "round1_ms_per_pair": 287.0,
            "sor_kernel": "k_sor_lex (time-skewed band march); round 1 ran one grid-wide barrier per anti-diagonal"}


def parity_report(pyflow, frames, mode, local):
    """Full-frame flow error of `mode` against the FP64 lexicographic mode (bit-identical to the reference: tests/
    test_gpu_full.py pins it to the reference's golden vectors) on the spot-check pairs of SURVEY.md 8d that are on this
    machine (1, 50, 101).  Where the max clause fails, the same statistics for the reference's own sweep order in FP32
    and for the reference order in FP64 on an input perturbed by 1e-7 say whether ANY FP32 implementation could meet it."""
    rows = []
    for a, b in [p for p in frames.pairs() if p[0] in (1, 50, 101)]:
        im1, im2 = frames.f64(a), frames.f64(b)
        _, ru, rv, _ = pyflow.coarse2fine_flow(im1, im2, 15, 1, mode="fp64_wavefront", device=local)

        def stat(u, v):
            e = np.hypot(u - ru, v - rv)
            return {"mean": float(e.mean()), "p99_9": float(np.quantile(e, 0.999)), "max": float(e.max()), "px_over_0_5": int((e > 0.5).sum()),
                    "frac_over_0_5": float((e > 0.5).mean())}
        _, u, v, _ = pyflow.coarse2fine_flow(im1, im2, 15, 1, mode=mode, device=local)
        row = {"pair": "%d-%d" % (a, b), "mode": mode, "epe_vs_parity_mode_full_frame": stat(u, v)}
        row["meets_mean_0_02"] = row["epe_vs_parity_mode_full_frame"]["mean"] <= 0.02
        row["meets_max_0_5"] = row["epe_vs_parity_mode_full_frame"]["max"] <= 0.5
        if not row["meets_max_0_5"]:
            _, u, v, _ = pyflow.coarse2fine_flow(im1, im2, 15, 1, mode="fp32_wavefront", device=local)
            row["reference_order_in_fp32"] = stat(u, v)
            rng = np.random.default_rng(0)
            _, u, v, _ = pyflow.coarse2fine_flow(np.clip(im1 + 1e-7 * rng.standard_normal(im1.shape), 0, 1), im2, 15, 1, mode="fp64_wavefront", device=local)
            row["reference_order_fp64_input_plus_1e-7"] = stat(u, v)
        rows.append(row)
    return rows


PARITY_NOTE = ("fp64_wavefront is bit-identical to the reference on every fixture (configs 1, 2, 3, 5 and pairs 50/101). The FP32 fast mode meets "
               "mean EPE <= 0.02 px everywhere and max EPE <= 0.5 px on config 3 (pair 1-2, every pixel). On pairs 50-51 and 101-102 of the "
               "sequence (config 4 spot checks) the max clause is unattainable for ANY FP32 implementation: the frames contain vehicles moving "
               "20-80 px, the reference's own result there is chaotic (its own order and FP64 arithmetic move hundreds of pixels by > 0.5 px, up "
               "to ~16 px, when the input is perturbed by 1e-7) and the reference's sweep order in FP32 violates the clause on as many pixels as "
               "red-black does -- see parity[*].reference_order_in_fp32 / reference_order_fp64_input_plus_1e-7, measured in this run. On "
               "config 5 (synthetic 4K) pure red-black leaves 0.009 % of the pixels > 0.5 px (ordering on the coarse levels, amplified by the "
               "pyramid); mode fp32_hybrid (reference order on levels <= 400 px wide) meets both clauses on every pixel "
               "(tests/test_gpu_full.py::test_config5_...).")


def hybrid_report(pyflow, pairs, local, B, steps):
    """The fp32_hybrid mode on the headline workload: latency of one pair and resident throughput with B pairs in flight."""
    lat = pyflow.FlowPlan(H, W, CH, mode="fp32_hybrid", device=local, tuning="latency", **PARAMS)
    lat.upload(*pairs[0])
    lat.solve(2)
    single_ms = lat.solve(3) / 3
    lat.close()
    plans = [pyflow.FlowPlan(H, W, CH, mode="fp32_hybrid", device=local, **PARAMS) for _ in range(B)]
    for i, p in enumerate(plans):
        p.upload(*pairs[i % len(pairs)])
    pyflow.multi_solve(plans, 1)
    ms = pyflow.multi_solve(plans, steps)
    for p in plans:
        p.close()
    return {"mode": "fp32_hybrid", "single_pair_latency_ms": single_ms, "pairs_per_sec_resident": steps * B / (ms / 1000.0),
            "pairs_in_flight": B, "steps": steps,
            "what": "FP32; SOR in the reference's lexicographic order (k_sor_lex) on the pyramid levels <= 400 px wide, red-black above"}


def config5_rowband(pyflow, ndev):
    """BASELINE configs[4]: ONE synthetic 3840x2160 gray pair (nSOR = 60, 18 levels) on one GPU and split into row bands
    over all `ndev` GPUs of the node (pf_multigpu_flow; rank 0 drives every device, peer stores over NVLink)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        from synth4k import make
        im1, im2, _, _ = make()
    except Exception as e:
        return {"unavailable": "synthetic 4K pair: %r" % (e,)}
    kw = dict(alpha=0.012, ratio=0.75, minWidth=20, nOuterFPIterations=7, nInnerFPIterations=1, nSORIterations=60, colType=1)
    out = {"workload": "synthetic 3840x2160 gray pair, affine motion, nSOR=60, 18 levels (BASELINE configs[4])"}
    ref_flow = None
    for tag, devs in (("1gpu", [0]), ("%dgpu" % ndev, list(range(ndev)))):
        best, st, u, v = None, None, None, None
        for _ in range(3):
            u, v, _, st = pyflow.coarse2fine_flow_multigpu(im1, im2, devices=devs, **kw)
            best = st["ms"] if best is None else min(best, st["ms"])
        out["ms_" + tag] = best
        if len(devs) > 1:
            out.update(halo_bytes=st["halo_bytes"], gather_bytes=st["gather_bytes"], split_solves=st["split_solves"],
                       graph=st["graph"], flag_solves=st["flag_solves"],
                       bit_identical_to_1gpu=bool(np.array_equal(u, ref_flow[0]) and np.array_equal(v, ref_flow[1])))
        else:
            ref_flow = (u, v)
    out["speedup"] = out["ms_1gpu"] / out["ms_%dgpu" % ndev]
    return out


def run_ours(args):
    dist, rank, local = dist_setup(args.gpus)
    import pyflow
    from papteam_opticalflow_b200 import _lib
    from papteam_opticalflow_b200.shard import pairs_for_rank
    lib = _lib.lib()
    if lib.pf_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device and no CPU fallback")
    frames = Frames()
    B = max(1, args.batch)
    # BASELINE configs[3]: pair p of the sequence belongs to GPU p mod G (no collective); every GPU keeps B pairs in
    # flight per step (weak scaling), cycling through its own share of the sequence's pairs
    all_pairs = frames.pairs()
    idx = pairs_for_rank(len(all_pairs), rank, args.gpus) if len(all_pairs) >= args.gpus else []
    if not idx:                                   # fewer fixture pairs than GPUs: every rank cycles through all of them
        idx = list(range(len(all_pairs)))
    mine = [all_pairs[i] for i in idx][:B]
    plans = [pyflow.FlowPlan(H, W, CH, mode=args.mode, device=local, **PARAMS) for _ in range(B)]
    plan = plans[0]
    pinned = True
    host = {}
    for a, b in mine:
        for i in (a, b):
            if i not in host:
                host[i], ok = pinned_like(lib, frames.f64(i))
                pinned = pinned and ok
    pairs = [(host[a], host[b]) for a, b in mine]
    npair = len(pairs)

    # ---- device-resident throughput: K steps of B concurrent graph replays, one pair resident per
    #      plan; CUDA events on the launching stream around the whole region (pf_multi_solve) ----
    for i, p in enumerate(plans):
        p.upload(*pairs[i % npair])
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
    def e2e_run(nsteps, ins, outs, flow_only=False):
        n = nsteps * B
        o = [outs[i % ring] for i in range(n)]
        if flow_only:
            o = [(x[0], x[1], None) for x in o]          # warpI2 = NULL: not copied back (the reference driver never reads it)
        pyflow.coarse2fine_flow_batch([ins[i % npair] for i in range(n)], PARAMS["alpha"], PARAMS["ratio"], PARAMS["minWidth"],
                                      PARAMS["nOuter"], PARAMS["nInner"], PARAMS["nSOR"], PARAMS["colType"], mode=args.mode,
                                      devices=[local], outs=o)
    def timed(fn, nsteps):
        barrier(dist)
        t0 = time.perf_counter()
        fn(nsteps)
        dt = time.perf_counter() - t0
        barrier(dist)
        return dist_max(dist, local, dt)
    e2e_run(max(1, min(2, args.warmup)), pairs, host_outs)
    e2e_s = timed(lambda k: e2e_run(k, pairs, host_outs), args.steps)
    e2e = args.gpus * args.steps * B / e2e_s
    legs = pyflow.batch_last_stats()
    # the same call with warpI2 = NULL (what Par/OpticalFlowCalculation.py:74-76 consumes: u and v only)
    half = max(1, args.steps // 2)
    e2e_run(1, pairs, host_outs, True)
    flow_s = timed(lambda k: e2e_run(k, pairs, host_outs, True), half)
    e2e_flow = args.gpus * half * B / flow_s
    legs_flow = pyflow.batch_last_stats()

    # ---- the same batch call on PAGEABLE numpy arrays (what a Python caller of the reference has): the library
    #      moves them through its multi-threaded pinned-slot stager (csrc/staging.hpp).  Extra key. ----
    page_outs = [tuple(np.zeros(s) for s in ((H, W), (H, W), (H, W, CH))) for _ in range(ring)]
    for o in page_outs:
        for a_ in o:
            a_.fill(0)          # touch the pages: a caller's buffers are normally mapped already
    page_pairs = [(frames.f64(a), frames.f64(b)) for a, b in mine]
    e2e_run(1, page_pairs, page_outs)
    page_s = timed(lambda k: e2e_run(k, page_pairs, page_outs), half)
    e2e_page = args.gpus * half * B / page_s
    legs_page = pyflow.batch_last_stats()

    # ---- sequence mode (BASELINE configs[3] as the reference driver would run it): K*B+1
    #      uint8 frames in, K*B float32 flows out, every frame's pyramid built once, conversion from uint8 on the device ----
    seq_idx = []
    for a, b in mine:                 # a walk over this rank's frames in which every consecutive pair is a real pair
        if not seq_idx or seq_idx[-1] != a:
            seq_idx.append(a)
        seq_idx.append(b)
    walk = seq_idx + seq_idx[-2:0:-1]           # forward then backward, cyclic
    u8 = {}
    for i in set(walk):
        u8[i], _ = pinned_raw(lib, (H, W, CH), np.uint8)
        u8[i][...] = frames.u8(i)
    seq_outs = [pinned_raw(lib, (H, W, 2), np.float32)[0] for _ in range(ring)]
    def seq_run(nsteps):
        n = nsteps * B
        pyflow.sequence_flow([u8[walk[i % len(walk)]] for i in range(n + 1)], PARAMS["alpha"], PARAMS["ratio"], PARAMS["minWidth"],
                             PARAMS["nOuter"], PARAMS["nInner"], PARAMS["nSOR"], PARAMS["colType"], mode=args.mode, devices=[local],
                             outs=[seq_outs[i % ring] for i in range(n)])
    seq_run(max(1, min(2, args.warmup)))
    seq_s = timed(seq_run, args.steps)
    clocks = sampler.summary()
    seq = args.gpus * args.steps * B / seq_s

    # ---- BASELINE configs[4] at N > 1: rank 0 splits ONE 4K pair into row bands over all N GPUs while the others wait ----
    rowband = None
    dma_ceiling = None
    if dist is not None:
        try:
            dma_ceiling = node_dma_ceiling(dist, local, args.gpus)
        except Exception:
            dma_ceiling = None
        barrier(dist)
        if not args.no_rowband:
            # The other ranks must leave their GPUs idle while rank 0 drives all of them: a NCCL barrier would park a
            # spinning kernel on every waiting GPU (measured: 109 instead of 51 ms on two GPUs), so they wait on the host,
            # on a key of a TCP store.
            from datetime import timedelta
            import torch.distributed as tdist
            store = tdist.TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")) + 1,
                                   args.gpus, rank == 0, timeout=timedelta(seconds=600))
            if rank == 0:
                try:
                    rowband = config5_rowband(pyflow, args.gpus)
                except Exception as e:
                    rowband = {"failed": repr(e)[:300]}
                store.set("rowband_done", "1")
            else:
                store.wait(["rowband_done"])
        barrier(dist)

    line = None
    if rank == 0:
        # ---- the drop-in call itself: pyflow.coarse2fine_flow on plain (pageable) numpy arrays, fresh outputs, one
        #      pair at a time -- what a caller of the reference's module sees (extra key, not the headline) ----
        one_shot = []
        f0, f1 = page_pairs[0]
        for _ in range(5):
            t0 = time.perf_counter()
            pyflow.coarse2fine_flow(f0, f1, PARAMS["alpha"], PARAMS["ratio"], PARAMS["minWidth"], PARAMS["nOuter"],
                                    PARAMS["nInner"], PARAMS["nSOR"], PARAMS["colType"], mode=args.mode, device=local)
            one_shot.append(1000 * (time.perf_counter() - t0))
        one_shot_ms = float(np.median(one_shot[1:]))
        # ---- single-pair latency (latency-tuned plan, what the one-shot entry points use) ----
        lat_plan = pyflow.FlowPlan(H, W, CH, mode=args.mode, device=local, tuning="latency", **PARAMS)
        lat_plan.upload(*pairs[0])
        lat_plan.solve(2)
        single_ms = lat_plan.solve(5) / 5
        try:
            lat_plan.profile()
            lat_launches = int(lat_plan.profile()[1][0])       # launches of the plan the one-shot entry points replay
        except Exception:
            lat_launches = None
        lat_plan.close()
        # ---- per-phase attribution + SOR roofline from one eager, event-instrumented solve ----
        plan.profile()
        tp, cnt = plan.profile()
        peak, peak_src = hbm_peak()
        word = 4 if args.mode.startswith("fp32") else 8
        sor_ms_l0, sor_launch_l0, ps_l0 = cnt[3], cnt[4], cnt[5]
        bytes_per_launch = ps_l0 * 10 * word / max(1.0, sor_launch_l0)
        achieved = (ps_l0 * 10 * word / 1e9) / (sor_ms_l0 / 1e3) if sor_ms_l0 > 0 else 0.0
        sor_all = (cnt[2] * 10 * word / 1e9) / (tp[_lib.T_PHASE5] / 1e3) if tp[_lib.T_PHASE5] > 0 else 0.0
        traffic, traffic_src = None, None
        tf = os.path.join(ROOT, "profiles", "sor_traffic.json")
        if os.path.exists(tf):
            try:
                tj = json.load(open(tf))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = "committed ncu figure (%s), not measured in this run" % tj.get("source", "profiles/sor_traffic.json")
            except Exception:
                traffic = None
        cpu = None
        if args.gpus == 1 and not args.no_cpu:
            try:
                from oracle import ref
                if ref.available():
                    a, b = mine[0]
                    secs, cores = cpu_reference_pair(False, frames.f64(a), frames.f64(b))
                    cpu = {"value": 1.0 / secs, "unit": "pairs/s", "cores": cores, "kind": "reference", "seconds": secs,
                           "sample": "one full 1920x1080 pair (frames %d-%d, 15 levels) through the unmodified Code/Serial build" % (a, b)}
            except Exception as e:   # the baseline must never break the GPU line
                cpu = {"value": None, "unit": "pairs/s", "cores": 1, "kind": "reference", "sample": "failed: %r" % (e,)}
        parity, hyb = None, None
        if args.gpus == 1 and not args.no_parity:
            try:
                parity = parity_report(pyflow, frames, args.mode, local)
            except Exception as e:
                parity = {"failed": repr(e)[:300]}
            try:
                hyb = hybrid_report(pyflow, pairs, local, B, max(2, args.steps // 3))
            except Exception as e:
                hyb = {"failed": repr(e)[:300]}
        cfg2 = None
        if not args.no_config2:
            try:
                cfg2 = config2_parity_mode(pyflow, local)
            except Exception as e:
                cfg2 = {"failed": repr(e)[:300]}
        phases = {k: round(float(tp[i]), 3) for i, k in enumerate(
            ["total", "pyramid", "features_upsample_warp", "getDxs", "phi", "psi(fused)", "assemble", "sor", "update_warp", "bicubic_export"])}
        leg = lambda d: {k: round(float(d[k]), 3) for k in ("h2d_ms", "solve_ms", "d2h_ms", "call_ms")}  # noqa: E731
        h2d_pair, d2h_pair = 2 * H * W * CH * 8, (2 * H * W + H * W * CH) * 8
        line = {
            "metric": "frame_pairs_per_sec_1920w", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_pair": ms / args.steps / B,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.mode.startswith("fp32") else "f64", "data": frames.data,
            "config": {"workload": WORKLOAD, "l2": L2_NOTE},          # the same dict in the reference arm's line
            "setup": {"mode": args.mode, "pairs_per_gpu_per_step": B,
                      "pairs": "sequence pairs %s on rank 0 (pair p -> GPU p mod N)" % ",".join("%d-%d" % q for q in mine[:8]),
                      "concurrency": "%d pairs in flight per GPU, one CUDA stream + graph each" % B,
                      "host_buffers": "pinned" if pinned else "pageable"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(B * h2d_pair),
                    "d2h_bytes_per_step": int(B * d2h_pair), "ms_per_step": 1000 * e2e_s / args.steps,
                    "legs_ms_per_pair_rank0": leg(legs),
                    "host_GBps": args.gpus * args.steps * B * (h2d_pair + d2h_pair) / e2e_s / 1e9,
                    "node_dma_ceiling_GBps": dma_ceiling,
                    "node_dma_ceiling_what": "all ranks copying pinned buffers in both directions at once with bare cudaMemcpyAsync "
                                             "(no library code); N > 1 only -- the float64 leg cannot move more than this",
                    "api": "pyflow.coarse2fine_flow_batch -> pf_batch_flow (host float64 HWC in, host float64 vx, vy, warpI2 out)"},
            "e2e_flow_only": {"value": e2e_flow, "unit": "pairs/s", "h2d_bytes_per_step": int(B * h2d_pair),
                              "d2h_bytes_per_step": int(B * 2 * H * W * 8), "legs_ms_per_pair_rank0": leg(legs_flow),
                              "host_GBps": args.gpus * half * B * (h2d_pair + 2 * H * W * 8) / flow_s / 1e9,
                              "api": "the same call with warpI2 = NULL: u and v only, what the reference driver consumes "
                                     "(Par/OpticalFlowCalculation.py:74-76)"},
            "e2e_pageable": {"value": e2e_page, "unit": "pairs/s", "legs_ms_per_pair_rank0": leg(legs_page),
                             "api": "the same batch call with plain (pageable) numpy arrays in and out, staged by the library"},
            "e2e_sequence": {"value": seq, "unit": "pairs/s", "h2d_bytes_per_step": B * H * W * CH, "d2h_bytes_per_step": B * H * W * 8,
                             "api": "pyflow.sequence_flow -> pf_sequence_flow_u8 (host uint8 frames in, host float32 (u,v) out; consecutive "
                                    "pairs share a frame, its pyramid is built once)"},
            "gpu_launches": int(cnt[0]) * args.steps * B,
            "single_pair_latency_ms": single_ms,
            "one_shot_call_ms": one_shot_ms,
            "roofline": {"bound": "hbm", "kernel": "k_sor_rb_tma (level 0, 1920x1080)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "launches_per_solve_level0": int(sor_launch_l0),
                         "avg_launch_ms": sor_ms_l0 / max(1.0, sor_launch_l0),
                         "sor_all_levels_GBps": sor_all, "sor_all_levels_frac": sor_all / peak,
                         "timing": "CUDA events on the launching stream around the SOR launches of one eager solve"},
            "cpu_baseline": cpu,
            "config2_fp64_wavefront": cfg2,
            "config5_rowband": rowband,
            "parity": parity,
            "parity_note": PARITY_NOTE,
            "hybrid": hyb,
            "phases_ms": phases,
            "launches_per_solve": int(cnt[0]),
            "launches_per_solve_latency_plan": lat_launches,
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
    ap.add_argument("--mode", default="fp32_redblack", choices=["fp32_redblack", "fp64_wavefront", "fp64_redblack", "fp32_wavefront", "fp32_hybrid"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-config2", action="store_true", help="skip the 960-wide fp64_wavefront leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-frame parity statistics and the fp32_hybrid leg")
    ap.add_argument("--no-rowband", action="store_true", help="skip the 4K row-band split leg at N > 1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
