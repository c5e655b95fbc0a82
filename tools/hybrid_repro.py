"""Is fp32_hybrid at 3840x2160 reproducible run to run, and equal between the two tunings? (developer tool)
usage: python tools/hybrid_repro.py [repeats]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import pyflow
from synth4k import make

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
im1, im2, gu, gv = make()
h, w, c = im1.shape
ref = pyflow.FlowPlan(h, w, c, nSOR=60, colType=1, mode="fp64_wavefront")
_, pu, pv, _ = ref.execute(im1, im2)
ref.close()
first = {}
for mode in ("fp32_hybrid", "fp32_wavefront"):
    for tuning in ("throughput", "latency"):
        plan = pyflow.FlowPlan(h, w, c, nSOR=60, colType=1, mode=mode, tuning=tuning)
        for r in range(reps):
            _, u, v, _ = plan.execute(im1, im2)
            e = np.hypot(u - pu, v - pv)
            key = (mode, tuning)
            same = "first" if key not in first else ("same" if np.array_equal(u, first[key][0]) and np.array_equal(v, first[key][1]) else
                                                      "DIFFERENT (%d px)" % int(((u != first[key][0]) | (v != first[key][1])).sum()))
            first.setdefault(key, (u.copy(), v.copy()))
            print("%s %-10s run %d: max EPE %.4f, px > 0.5: %d, mean %.6f  [%s]" % (mode, tuning, r, e.max(), int((e > 0.5).sum()), e.mean(), same), flush=True)
        plan.close()
    a, b = first[(mode, "throughput")], first[(mode, "latency")]
    print("%s: throughput == latency tuning: %s" % (mode, np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])), flush=True)
