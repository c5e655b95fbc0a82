"""k_sor_lex (band-march) against the grid-synchronised wavefront kernel: bit-exact in FP64 on random systems of many
shapes, then time per solve of both and of the red-black kernel on the level sizes of a 1920-wide pyramid.
usage: python tools/lex_check.py [check|time|all]"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from papteam_opticalflow_b200 import _lib
import gpu_util as G
L = _lib.lib()
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def system(h, w, seed):
    r = np.random.default_rng(seed)
    phi = r.random((h, w)) * 40 + 0.5
    dxy = r.standard_normal((h, w)) * 0.2
    dx2 = r.random((h, w)) + 0.05
    dy2 = r.random((h, w)) + 0.05
    bu = r.standard_normal((h, w))
    bv = r.standard_normal((h, w))
    return phi, dxy, dx2, dy2, bu, bv


if what in ("check", "all"):
    bad = 0
    shapes = [(1, 1), (1, 7), (5, 1), (3, 5), (19, 34), (31, 33), (32, 32), (33, 65), (34, 61), (45, 81), (64, 64), (70, 100),
              (107, 192), (192, 341), (270, 480), (131, 67)]
    for (h, w) in shapes:
        for nsor in (1, 2, 7, 8, 9, 17, 30, 72):
            if h * w > 40000 and nsor not in (9, 30):
                continue
            s = system(h, w, h * 1000 + w + nsor)
            os.environ["PF_LEX_IMPL"] = "coop"
            ru, rv = G.sor(*s, 0.012, nsor, G.F64)
            os.environ["PF_LEX_IMPL"] = "band"
            t = time.time()
            try:
                gu, gv = G.sor(*s, 0.012, nsor, G.F64)
            except Exception as e:
                print("%dx%d nsor=%d: FAILED %s" % (w, h, nsor, e), flush=True)
                bad += 1
                continue
            ok = np.array_equal(gu, ru) and np.array_equal(gv, rv)
            if not ok:
                d = np.abs(gu - ru)
                ij = np.unravel_index(np.argmax(d), d.shape)
                nb = int((d > 0).sum())
                first = np.argwhere(d > 0)[0]
                print("%dx%d nsor=%d: MISMATCH max %.3g at %s, %d px differ, first at %s" % (w, h, nsor, d.max(), ij, nb, first), flush=True)
                bad += 1
            else:
                print("%dx%d nsor=%d: identical (%.2f s)" % (w, h, nsor, time.time() - t), flush=True)
            os.environ["PF_LEX_IMPL"] = "coop"
            r32 = G.sor(*s, 0.012, nsor, G.F32LEX)
            os.environ["PF_LEX_IMPL"] = "band"
            g32 = G.sor(*s, 0.012, nsor, G.F32LEX)
            d32 = max(np.abs(g32[0] - r32[0]).max(), np.abs(g32[1] - r32[1]).max())
            if d32 > 1e-3 * max(1.0, np.abs(r32[0]).max()):
                print("   fp32 band vs coop: max diff %.3g (scale %.3g)" % (d32, np.abs(r32[0]).max()), flush=True)
                bad += 1
    print("CHECK:", "all identical" if bad == 0 else "%d failures" % bad, flush=True)

if what in ("time", "all"):
    sizes = [(1920, 1080, 30), (1440, 810, 33), (1080, 607, 36), (810, 455, 39), (607, 341, 42), (455, 256, 45), (341, 192, 48), (256, 143, 51),
             (192, 107, 54), (144, 81, 57), (108, 60, 60), (81, 45, 63), (60, 33, 66), (46, 26, 69), (34, 19, 72), (960, 540, 30), (3840, 2160, 60)]
    for w, h, nsor in sizes:
        row = []
        for name, mode, impl in (("rb32", 1, "band"), ("lex32 band", 3, "band"), ("lex64 band", 0, "band"), ("lex32 coop", 3, "coop"), ("lex64 coop", 0, "coop")):
            if impl == "coop" and w > 1000:
                continue
            os.environ["PF_LEX_IMPL"] = impl
            ms = C.c_double(); ln = C.c_double()
            rc = L.pf_bench_sor(h, w, nsor, 4, mode, 0, C.byref(ms), C.byref(ln))
            row.append("%s %8.1f" % (name, ms.value * 1000 if rc == 0 else -1))
        print("%4dx%-4d nsor=%2d us/solve  " % (w, h, nsor) + " | ".join(row), flush=True)
