import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, pyflow
from conftest import load_frame
from synth4k import make
cases = [("960 rgb", load_frame(960, 1), load_frame(960, 2), (0.012, 0.75, 20, 7, 1, 30, 0), 50000)]
im1, im2, _, _ = make(1080, 1920)
cases.append(("1920 gray sor60", im1, im2, (0.012, 0.75, 20, 7, 1, 60, 1), 400000))
for name, a, b, args, thr in cases:
    u0, v0, w0 = pyflow.coarse2fine_flow(a, b, *args, mode="fp32_redblack")
    for devs in ([0, 0], [0, 0, 0], [0] * 5):
        u, v, w2, st = pyflow.coarse2fine_flow_multigpu(a, b, *args, devices=devs, split_min_pixels=thr)
        d = np.abs(u - u0)
        ys = np.where(d.max(axis=1) > 0)[0]
        print(name, len(devs), "bands: identical", np.array_equal(u, u0) and np.array_equal(v, v0), "max|du|", d.max(),
              "rows differing", (ys.min(), ys.max(), len(ys)) if len(ys) else None, "split solves", st["split_solves"])
