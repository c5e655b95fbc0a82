"""One size of k_sor_lex, a few solves (ncu target).  usage: python tools/lex_one.py W H NSOR [mode]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from papteam_opticalflow_b200 import _lib
L = _lib.lib()
w, h, nsor = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ms = C.c_double(); ln = C.c_double()
L.pf_bench_sor(h, w, nsor, 3, mode, 0, C.byref(ms), C.byref(ln))
print("%.1f us/solve" % (ms.value * 1000))
