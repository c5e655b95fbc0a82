"""How sensitive is the reference's own result (the FP64 lexicographic mode is bit-identical to it) on the 1920-wide
HoChiMinh pairs?  Compares against it: (a) the same arithmetic order in FP32, (b) red-black order in FP64, (c) the
reference order and arithmetic on inputs perturbed by 1e-7 / 1e-10 (far below the 1/255 quantisation of the frames),
(d) the hybrid fast mode."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_frame
import pyflow
pairs = [(1, 2), (50, 51), (101, 102)] if len(sys.argv) < 2 else [(int(sys.argv[1]), int(sys.argv[1]) + 1)]


def stat(tag, u, v, ru, rv):
    e = np.hypot(u - ru, v - rv)
    print("  %-34s EPE mean %.5f p99.9 %.4f max %7.3f  n>0.5 %6d (%.4f%%)" % (tag, e.mean(), np.quantile(e, 0.999), e.max(), (e > 0.5).sum(), 100 * (e > 0.5).mean()), flush=True)
    return e


for a, b in pairs:
    im1, im2 = load_frame(1920, a), load_frame(1920, b)
    print("pair %d-%d" % (a, b), flush=True)
    _, ru, rv, _ = pyflow.coarse2fine_flow(im1, im2, 15, 1, mode="fp64_wavefront")
    for mode in ("fp32_wavefront", "fp64_redblack", "fp32_redblack"):
        _, u, v, _ = pyflow.coarse2fine_flow(im1, im2, 15, 1, mode=mode)
        e = stat(mode, u, v, ru, rv)
        if mode == "fp32_redblack":
            ys, xs = np.nonzero(e > 0.5)
            if len(ys):
                print("     outliers: rows %d..%d cols %d..%d; |ref flow| there mean %.2f max %.2f" % (ys.min(), ys.max(), xs.min(), xs.max(),
                      np.hypot(ru, rv)[e > 0.5].mean(), np.hypot(ru, rv)[e > 0.5].max()))
    rng = np.random.default_rng(0)
    for eps in (1e-7, 1e-10, 1e-13):
        p1 = np.clip(im1 + eps * rng.standard_normal(im1.shape), 0, 1)
        _, u, v, _ = pyflow.coarse2fine_flow(p1, im2, 15, 1, mode="fp64_wavefront")
        stat("reference order, input + %.0e noise" % eps, u, v, ru, rv)
