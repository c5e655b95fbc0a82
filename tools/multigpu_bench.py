"""BASELINE config 5: synthetic 3840x2160 gray pair, nSOR=60, solved by 1..N GPUs with the row-band split."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, pyflow
from papteam_opticalflow_b200 import _lib
from synth4k import make
n = _lib.lib().pf_device_count()
im1, im2, gu, gv = make()
args = (0.012, 0.75, 20, 7, 1, 60, 1)
plan = pyflow.FlowPlan(2160, 3840, 1, alpha=0.012, ratio=0.75, minWidth=20, nOuter=7, nInner=1, nSOR=60, colType=1, mode="fp32_redblack")
plan.upload(im1, im2); plan.solve(1)
ms1 = plan.solve(3) / 3
_, u0, v0, w0 = plan.execute(im1, im2)
print("1 GPU (CUDA graph): %.2f ms per 4K pair" % ms1)
thr = int(sys.argv[1]) if len(sys.argv) > 1 else -1
for G in [g for g in (1, 2, 4, 8) if g <= n]:
    best = None
    for rep in range(3):
        u, v, w2, st = pyflow.coarse2fine_flow_multigpu(im1, im2, *args, devices=list(range(G)), split_min_pixels=thr)
        best = st["ms"] if best is None else min(best, st["ms"])
    same = np.array_equal(u, u0) and np.array_equal(v, v0) and np.array_equal(w2, w0)
    print("%d GPU(s) row-band split (graph): %.2f ms  speed-up vs 1 GPU %.2fx  identical=%s  halo %.1f MB gather %.1f MB split solves %d"
          % (G, best, (base / best) if G > 1 else 1.0, same, st["halo_bytes"] / 1e6, st["gather_bytes"] / 1e6, st["split_solves"]))
    if G == 1: base = best
