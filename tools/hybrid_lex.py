"""Experiment: which pyramid levels must run in the reference's lexicographic order for the FP32 fast mode to meet
max EPE <= 0.5 px on BASELINE config 5?  PF_LEX_FROM=k runs levels >= k with the lexicographic kernel.
usage: python tools/hybrid_lex.py ref | run <k or -1>"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from synth4k import make
ARGS = (0.012, 0.75, 20, 7, 1, 60, 1)
REF = "/tmp/ref4k.npz"
if sys.argv[1] == "ref":
    import pyflow
    im1, im2, gu, gv = make()
    t = time.time()
    u, v, _ = pyflow.coarse2fine_flow(im1, im2, *ARGS, mode="fp64_wavefront")
    print("fp64_wavefront 4K: %.1f s" % (time.time() - t))
    np.savez(REF, u=u, v=v)
else:
    k = int(sys.argv[2])
    if k >= 0:
        os.environ["PF_LEX_FROM"] = str(k)
    mode = sys.argv[3] if len(sys.argv) > 3 else "fp32_redblack"
    import pyflow
    im1, im2, gu, gv = make()
    r = np.load(REF)
    t = time.time()
    u, v, _ = pyflow.coarse2fine_flow(im1, im2, *ARGS, mode=mode)
    dt = time.time() - t
    e = np.hypot(u - r["u"], v - r["v"])
    print("lex_from=%d %s: EPE mean %.5f p99.9 %.4f max %.3f  frac>0.5 %.5f%%  n>0.5 %d  (%.1f s)"
          % (k, mode, e.mean(), np.quantile(e, 0.999), e.max(), 100 * (e > 0.5).mean(), (e > 0.5).sum(), dt), flush=True)
