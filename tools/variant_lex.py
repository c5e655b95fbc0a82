"""Developer tool: single-pair time of the modes that use k_sor_lex in FP32 (fp32_wavefront, fp32_hybrid) with variant builds
(tools/build_variant.sh name -DPF_LEX_NS=...), plus a checksum of the flow.  usage: python tools/variant_lex.py base name1 ..."""
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
out = []
for w, mode in ((960, "fp32_wavefront"), (1920, "fp32_wavefront"), (1920, "fp32_hybrid")):
    a, b = load_frame(w, 1), load_frame(w, 2)
    plan = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode=mode, tuning="latency")
    plan.upload(a, b); plan.solve(2)
    ms = min(plan.solve(3) / 3 for _ in range(2))
    _, u, v, wi = plan.execute(a, b)
    out.append("%d %s %.2f ms (%s)" % (w, mode, ms, hashlib.sha1(u.tobytes() + v.tobytes()).hexdigest()[:8]))
    plan.close()
print("%-8s %s" % (name, " | ".join(out)), flush=True)
