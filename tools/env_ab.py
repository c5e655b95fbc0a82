"""Developer tool: single-pair latency (graph replay) and per-level SOR time of latency- and throughput-tuned plans under two
values of an environment switch, and equality of the results.
usage: python tools/env_ab.py VAR value_a value_b [width]      e.g.  PF_SOR_SMALL_WARPS 16 4   |   PF_PDL 0 1"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
var, va, vb = sys.argv[1:4]
w = sys.argv[4] if len(sys.argv) > 4 else "1920"
if os.environ.get("_ENV_AB_CHILD") is None:
    for val in (va, vb):
        subprocess.run([sys.executable, __file__, var, va, vb, w], env=dict(os.environ, _ENV_AB_CHILD=val, **{var: val}))
    import numpy as np
    a, b = np.load("/tmp/envab_%s.npz" % va), np.load("/tmp/envab_%s.npz" % vb)
    print("results identical:", all(np.array_equal(a[k], b[k]) for k in ("u", "v", "w")))
    sys.exit(0)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyflow
from conftest import load_frame
val = os.environ["_ENV_AB_CHILD"]
a, b = load_frame(int(w), 1), load_frame(int(w), 2)
for tuning in ("latency", "throughput"):
    plan = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode="fp32_redblack", tuning=tuning)
    plan.upload(a, b); plan.solve(3)
    ms = min(plan.solve(5) / 5 for _ in range(3))
    _, u, v, wi = plan.execute(a, b)
    plan.profile(); plan.profile()
    lt = plan.level_timings()
    print("%s=%s %-10s plan: %.3f ms per %s-wide pair (graph replay); eager SOR ms of the three coarsest levels: %s"
          % (var, val, tuning, ms, w, " ".join("%.3f" % lt[k][7] for k in range(plan.levels - 3, plan.levels))), flush=True)
    if tuning == "latency":
        np.savez("/tmp/envab_%s.npz" % val, u=u, v=v, w=wi)
    plan.close()
