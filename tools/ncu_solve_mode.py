"""One eager (un-graphed) 1920x1080 solve in a given mode / tuning for an ncu launch list: the LAST solve is the warm one.
usage: python tools/ncu_solve_mode.py [mode] [tuning] [nsolves]"""
import os, sys
os.environ["PF_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyflow
from conftest import load_frame
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32_redblack"
tuning = sys.argv[2] if len(sys.argv) > 2 else "throughput"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
a, b = load_frame(1920, 1), load_frame(1920, 2)
plan = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode=mode, tuning=tuning)
plan.upload(a, b)
print("ms", plan.solve(n) / n)
