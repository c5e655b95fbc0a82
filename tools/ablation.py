"""Where the saturated-GPU time of a solve goes: throughput of 16 concurrent resident solves (1920x1080 RGB)
with parts of the iteration switched off through the public parameters (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyflow
from conftest import load_frame
a, b = load_frame(1920, 1), load_frame(1920, 2)
B = 16
def run(tag, **kw):
    plans = []
    for _ in range(B):
        p = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode="fp32_redblack", **kw); p.upload(a, b); p.solve(1); plans.append(p)
    pyflow.multi_solve(plans, 1)
    ms = pyflow.multi_solve(plans, 4)
    print("%-28s %.3f ms per pair  %.1f pairs/s" % (tag, ms / (4 * B), 1000 * 4 * B / ms), flush=True)
    for p in plans: p.close() if hasattr(p, "close") else None
run("full 7/1/30")
run("nSOR=0", nSOR=0)
run("nOuter=0 (pyramid+export)", nOuter=0)
