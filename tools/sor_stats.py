"""Developer tool: per-CTA phase cycles of k_sor_rb_tma from the instrumented build
(tools/build_variant.sh stats -DPF_SOR_STATS=1).  usage: python tools/sor_stats.py [w h nsor [fuse]]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from papteam_opticalflow_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "bin", "lib_stats.so")
w, h, nsor = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1920, 1080, 30)
if len(sys.argv) > 4:
    os.environ["PF_SOR_FUSE"] = sys.argv[4]
L = _lib.lib()
ms = C.c_double(); ln = C.c_double()
rc = L.pf_bench_sor(h, w, nsor, 2, 1, 0, C.byref(ms), C.byref(ln))
L.pf_last_error.restype = C.c_char_p
print("rc %d: %.1f us per solve, %d launches %s" % (rc, ms.value * 1000, int(ln.value), L.pf_last_error() if rc else ""), flush=True)
