"""Developer tool: single-pair latency of a latency-tuned plan with / without programmatic dependent launch (PF_PDL),
and equality of the results.  usage: python tools/pdl_check.py [width]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) < 3:
    w = sys.argv[1] if len(sys.argv) > 1 else "1920"
    for pdl in ("0", "1"):
        subprocess.run([sys.executable, __file__, w, pdl], env=dict(os.environ, PF_PDL=pdl))
    import numpy as np
    a, b = np.load("/tmp/pdl_0.npz"), np.load("/tmp/pdl_1.npz")
    print("results identical:", all(np.array_equal(a[k], b[k]) for k in ("u", "v", "w")))
    sys.exit(0)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyflow
from conftest import load_frame
w = int(sys.argv[1])
a, b = load_frame(w, 1), load_frame(w, 2)
for tuning in ("latency", "throughput"):
    plan = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode="fp32_redblack", tuning=tuning)
    plan.upload(a, b); plan.solve(3)
    ms = min(plan.solve(5) / 5 for _ in range(3))
    _, u, v, wi = plan.execute(a, b)
    print("PF_PDL=%s %-10s plan: %.3f ms per %d-wide pair (graph replay)" % (os.environ.get("PF_PDL"), tuning, ms, w), flush=True)
    if tuning == "latency":
        np.savez("/tmp/pdl_%s.npz" % os.environ.get("PF_PDL"), u=u, v=v, w=wi)
    plan.close()
