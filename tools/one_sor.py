import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
from papteam_opticalflow_b200 import _lib
L = _lib.lib()
w, h, nsor, fuse = [int(x) for x in sys.argv[1:5]]
os.environ["PF_SOR_FUSE"] = str(fuse)
ms = C.c_double(); ln = C.c_double()
rc = L.pf_bench_sor(h, w, nsor, 6, 1, 0, C.byref(ms), C.byref(ln))
print(w, h, nsor, fuse, "rc", rc, "us", ms.value * 1000, L.pf_last_error())
