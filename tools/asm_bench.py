"""Time of the fused assembly kernel at level 0 (developer tool): eager profile of a 1-level plan."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyflow
from conftest import load_frame
a, b = load_frame(1920, 1), load_frame(1920, 2)
plan = pyflow.FlowPlan(1080, 1920, 3, levels=1, nOuter=10, mode="fp32_redblack")
plan.upload(a, b); plan.solve(1); plan.profile(); t, cnt = plan.profile()
print("asm %.1f us/call  sor %.1f us/solve  upd %.1f us/call  (10 outer iterations, 1920x1080)" % (t[6]*100, t[7]*100, t[8]*1000/11))
