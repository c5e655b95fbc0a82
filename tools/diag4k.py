import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, pyflow
from synth4k import make
g = np.load(os.path.join(ROOT, "tests/golden/synth4k_L18_sor60_s16.npz")); s = int(g["stride"])
im1, im2, gu, gv = make()
args = (0.012, 0.75, 20, 7, 1, 60, 1)
ref = None
for mode in ("fp64_wavefront", "fp64_redblack", "fp32_wavefront", "fp32_redblack"):
    u, v, w2 = pyflow.coarse2fine_flow(im1, im2, *args, mode=mode)
    if ref is None: ref = (u, v)
    e = np.hypot(u - ref[0], v - ref[1])
    iy, ix = np.unravel_index(np.argmax(e), e.shape)
    print("%-15s full-res EPE vs fp64_wavefront: mean %.6f p99.9 %.5f max %.4f at (y=%d,x=%d)  n>0.5: %d  n>0.1: %d" %
          (mode, e.mean(), np.quantile(e, 0.999), e.max(), iy, ix, (e > 0.5).sum(), (e > 0.1).sum()))
    if mode != "fp64_wavefront":
        ys, xs = np.where(e > 0.5)
        if len(ys): print("    >0.5 px region: y %d..%d x %d..%d ; flow there u=%.2f v=%.2f (ref u=%.2f v=%.2f)" % (ys.min(), ys.max(), xs.min(), xs.max(), u[iy, ix], v[iy, ix], ref[0][iy, ix], ref[1][iy, ix]))
