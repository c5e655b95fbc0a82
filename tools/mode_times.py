"""Single-pair solve time of every mode at a given width (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyflow
from conftest import load_frame
for w in (240, 960, 1920):
    a, b = load_frame(w, 1), load_frame(w, 2)
    row = []
    for mode in ("fp32_redblack", "fp64_redblack", "fp32_wavefront", "fp64_wavefront"):
        p = pyflow.FlowPlan(a.shape[0], a.shape[1], 3, mode=mode); p.upload(a, b); p.solve(1)
        row.append("%s %.1f ms" % (mode, p.solve(2) / 2)); p.close()
    print(w, " | ".join(row))
