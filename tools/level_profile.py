"""Per-level, per-phase CUDA-event times of one eager solve (developer tool).
usage: python tools/level_profile.py [width] [mode] [tuning]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pyflow
from conftest import load_frame

w = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32_redblack"
tuning = sys.argv[3] if len(sys.argv) > 3 else "throughput"
a, b = load_frame(w, 1), load_frame(w, 2)
plan = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode=mode, tuning=tuning)
plan.upload(a, b)
plan.solve(2)
ms = plan.solve(5) / 5
plan.profile()
t, cnt = plan.profile()
lt = plan.level_timings()
names = ["tot", "pyr", "feat", "getDxs", "phi", "psi", "asm", "SOR", "upd", "post"]
print("graph solve: %.3f ms/pair   eager profiled: %.3f ms   launches %d (SOR %d)" % (ms, t[12], cnt[0], cnt[1]))
print("lvl " + " ".join("%8s" % n for n in names[1:]))
for k in range(plan.levels):
    print("%3d " % k + " ".join("%8.3f" % lt[k][i] for i in range(1, 10)))
print("sum " + " ".join("%8.3f" % t[i] for i in range(1, 10)))
