"""Developer tool: time of one SOR solve (pf_bench_sor, resident synthetic coefficients) for every forced number of fused
sweeps per pass against the latency fit's own choice.  usage: python tools/sor_fuse_sweep.py w h nsor [w h nsor ...]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PF_SOR_TUNE"] = "latency"
from papteam_opticalflow_b200 import _lib
L = _lib.lib()
a = [int(x) for x in sys.argv[1:]]
for w, h, nsor in zip(a[0::3], a[1::3], a[2::3]):
    out = []
    for fuse in ["auto"] + [str(f) for f in range(3, 16)]:
        if fuse == "auto": os.environ.pop("PF_SOR_FUSE", None)
        else: os.environ["PF_SOR_FUSE"] = fuse
        ms = C.c_double(); ln = C.c_double()
        rc = L.pf_bench_sor(h, w, nsor, 8, 1, 0, C.byref(ms), C.byref(ln))
        out.append("%s:%.1f(%d)" % (fuse, ms.value * 1000 if rc == 0 else -1, int(ln.value)))
    print("%dx%d nsor %d  us per solve (passes): %s" % (w, h, nsor, "  ".join(out)), flush=True)
