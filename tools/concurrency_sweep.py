"""Throughput of B concurrent resident solves on one GPU (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyflow
from conftest import load_frame
w = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
a, b = load_frame(w, 1), load_frame(w, 2)
plans = []
for B in (1, 4, 8, 12, 16, 24, 32):
    while len(plans) < B:
        p = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode="fp32_redblack"); p.upload(a, b); p.solve(2); plans.append(p)
    pyflow.multi_solve(plans[:B], 1)
    ms = pyflow.multi_solve(plans[:B], 4)
    print("B=%d  %.2f ms per pair  %.1f pairs/s" % (B, ms / (4 * B), 1000 * 4 * B / ms))
