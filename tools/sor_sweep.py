"""SOR kernel timing vs fused sweeps per launch for the level sizes of a 1920-wide pyramid."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from papteam_opticalflow_b200 import _lib
L = _lib.lib()
sizes = [(1920,1080,30),(1440,810,33),(1080,607,36),(810,455,39),(607,341,42),(455,256,45),(341,192,48),(256,143,51),(192,107,54),(144,81,57),(108,60,60),(81,45,63),(60,33,66)]
if len(sys.argv) > 1: sizes = sizes[:int(sys.argv[1])]
for w,h,nsor in sizes:
    row = []
    for fuse in (0,1,2,3,4,5,6,7,8,10):
        os.environ["PF_SOR_FUSE"] = str(fuse)
        ms = C.c_double(); ln = C.c_double()
        rc = L.pf_bench_sor(h, w, nsor, 6, 1, 0, C.byref(ms), C.byref(ln))
        row.append("%s:%6.1f(%d)" % ("auto" if fuse == 0 else "T%d" % fuse, ms.value*1000 if rc == 0 else -1, int(ln.value)))
    print("%4dx%-4d nsor=%d us/solve  " % (w,h,nsor) + " ".join(row))
