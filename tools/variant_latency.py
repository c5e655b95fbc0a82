"""Developer tool: single-pair latency (latency-tuned plan, graph replay) and the eager assembly / SOR / update time of the
levels >= 5 with variant builds of the library (tools/build_variant.sh), plus a checksum of the flow.
usage: python tools/variant_latency.py base name1 name2 ..."""
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
import hashlib, numpy as np, pyflow
from conftest import load_frame
a, b = load_frame(1920, 1), load_frame(1920, 2)
plan = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode="fp32_redblack", tuning="latency")
plan.upload(a, b); plan.solve(3)
ms = min(plan.solve(5) / 5 for _ in range(3))
_, u, v, wi = plan.execute(a, b)
plan.profile(); plan.profile()
lt = plan.level_timings()
coarse = lt[5:].sum(axis=0)
print("%-10s %.3f ms per pair | levels >= 5, eager: assemble %.3f SOR %.3f update %.3f ms | flow sha1 %s"
      % (name, ms, coarse[6], coarse[7], coarse[8], hashlib.sha1(u.tobytes() + v.tobytes()).hexdigest()[:12]), flush=True)
