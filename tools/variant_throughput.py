"""Developer tool: resident throughput (16 pairs in flight, 1920x1080 RGB, defaults) with variant builds of the library
(tools/build_variant.sh).  usage: python tools/variant_throughput.py base name1 name2 ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2:
    for n in sys.argv[1:]:
        subprocess.run([sys.executable, __file__, n])
    sys.exit(0)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from papteam_opticalflow_b200 import _lib
name = sys.argv[1]
if name != "base":
    _lib.LIB_PATH = os.path.join(ROOT, "tools", "bin", "lib_%s.so" % name)
import pyflow
from conftest import load_frame
a, b = load_frame(1920, 1), load_frame(1920, 2)
B = 16
plans = []
for _ in range(B):
    p = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode="fp32_redblack"); p.upload(a, b); p.solve(1); plans.append(p)
single = plans[0].solve(3) / 3
pyflow.multi_solve(plans, 2)
ms = pyflow.multi_solve(plans, 6)
print("%-12s PF_SOR_FUSE=%s  %.3f ms per pair  %.1f pairs/s   single pair %.2f ms" % (name, os.environ.get("PF_SOR_FUSE", "auto"), ms / (6 * B), 1000 * 6 * B / ms, single), flush=True)
