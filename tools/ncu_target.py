"""Small eager (un-graphed) workload for ncu: ONE pyramid level at full 1920x1080 resolution, so
the first launches of every kernel are the level-0 ones.  usage: python tools/ncu_target.py [width]"""
import os, sys
os.environ["PF_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyflow
from conftest import load_frame
w = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
a, b = load_frame(w, 1), load_frame(w, 2)
plan = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, levels=1, nOuter=2, mode="fp32_redblack")
plan.upload(a, b)
print("ms", plan.solve(1))
