"""One full eager (un-graphed) 1920x1080 solve for an ncu launch list: the LAST solve's launches are the warm ones.
usage: python tools/ncu_solve.py [nsolves]"""
import os, sys
os.environ["PF_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyflow
from conftest import load_frame
a, b = load_frame(1920, 1), load_frame(1920, 2)
plan = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode="fp32_redblack")
plan.upload(a, b)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
print("ms", plan.solve(n) / n)
