"""Developer tool: per-level phase times of the synthetic 3840x2160 gray pair (BASELINE config 5), one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import pyflow
from synth4k import make
im1, im2, _, _ = make()
plan = pyflow.FlowPlan(2160, 3840, 1, alpha=0.012, ratio=0.75, minWidth=20, nOuter=7, nInner=1, nSOR=60, colType=1, mode="fp32_redblack")
plan.upload(im1, im2); plan.solve(2)
print("graph solve %.2f ms" % (plan.solve(3) / 3))
plan.profile(); t, cnt = plan.profile(); lt = plan.level_timings()
names = ["tot", "pyr", "feat", "getDxs", "phi", "psi", "asm", "SOR", "upd", "post"]
print("lvl " + " ".join("%8s" % n for n in names[1:]))
for k in range(plan.levels):
    print("%3d " % k + " ".join("%8.3f" % lt[k][i] for i in range(1, 10)))
print("sum " + " ".join("%8.3f" % t[i] for i in range(1, 10)), " eager total %.2f" % t[12])
