"""Timing of k_sor_lex under its developer knobs (publication interval, fence kind, barrier kind)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from papteam_opticalflow_b200 import _lib
L = _lib.lib()
for w, h, nsor in ((341, 20, 8), (341, 192, 48), (34, 19, 72), (960, 540, 30)):
    for pub, opt in ((1, 1), (1, 0), (4, 0), (8, 0), (16, 0), (4, 2), (4, 4), (4, 6), (1000000, 2)):
        os.environ["PF_LEX_PUB"] = str(pub); os.environ["PF_LEX_OPT"] = str(opt)
        ms = C.c_double(); ln = C.c_double()
        rc = L.pf_bench_sor(h, w, nsor, 4, 3, 0, C.byref(ms), C.byref(ln))
        print("%4dx%-4d nsor=%2d pub=%-7d opt=%d: %8.1f us/solve (rc %d)" % (w, h, nsor, pub, opt, ms.value * 1000, rc), flush=True)
