"""Per-CTA timeline of k_sor_lex on one level size (developer aid).  usage: python tools/lex_stats.py W H NSOR [mode]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from papteam_opticalflow_b200 import _lib
L = _lib.lib()
w, h, nsor = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ms = C.c_double(); ln = C.c_double()
L.pf_bench_sor(h, w, nsor, 3, mode, 0, C.byref(ms), C.byref(ln))
print("plain: %.1f us/solve" % (ms.value * 1000))
os.environ["PF_LEX_STATS"] = "1"
L.pf_bench_sor(h, w, nsor, 2, mode, 0, C.byref(ms), C.byref(ln))
