"""Which pyramid levels must run in lexicographic order for the FP32 fast mode to meet max EPE <= 0.5 px on the 1920-wide
HoChiMinh pairs (BASELINE configs 3/4)?  PF_LEX_FROM=k runs levels >= k with k_sor_lex.
usage: python tools/hybrid_pairs.py ref | run <k or -1>"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_frame
PAIRS = [(1, 2), (50, 51), (101, 102)]
if sys.argv[1] == "ref":
    import pyflow
    for a, b in PAIRS:
        t = time.time()
        _, u, v, _ = pyflow.coarse2fine_flow(load_frame(1920, a), load_frame(1920, b), 15, 1, mode="fp64_wavefront")
        print("fp64_wavefront pair %d: %.2f s" % (a, time.time() - t), flush=True)
        np.savez("/tmp/ref1920_%d.npz" % a, u=u, v=v)
else:
    k = int(sys.argv[2])
    if k >= 0:
        os.environ["PF_LEX_FROM"] = str(k)
    import pyflow
    for a, b in PAIRS:
        r = np.load("/tmp/ref1920_%d.npz" % a)
        im1, im2 = load_frame(1920, a), load_frame(1920, b)
        pyflow.coarse2fine_flow(im1, im2, 15, 1, mode="fp32_redblack")
        t = time.time()
        _, u, v, _ = pyflow.coarse2fine_flow(im1, im2, 15, 1, mode="fp32_redblack")
        dt = time.time() - t
        e = np.hypot(u - r["u"], v - r["v"])
        print("lex_from=%2d pair %3d: EPE mean %.5f p99.9 %.4f max %.3f  n>0.5 %d  (%.1f ms/call)"
              % (k, a, e.mean(), np.quantile(e, 0.999), e.max(), (e > 0.5).sum(), dt * 1e3), flush=True)
