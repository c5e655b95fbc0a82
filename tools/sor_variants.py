"""Developer tool: time the SOR solve of the finest levels with variant builds of the library
(tools/build_variant.sh).  usage: python tools/sor_variants.py name1 name2 ... ; env PF_SOR_FUSE applies."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2:
    for n in sys.argv[1:]:
        subprocess.run([sys.executable, __file__, n])
    sys.exit(0)
sys.path.insert(0, ROOT)
from papteam_opticalflow_b200 import _lib
name = sys.argv[1]
if name != "base":
    _lib.LIB_PATH = os.path.join(ROOT, "tools", "bin", "lib_%s.so" % name)
L = _lib.lib()
sizes = [(1920, 1080, 30), (1440, 810, 33), (810, 455, 39)]
out = []
for w, h, nsor in sizes:
    for fuse in (os.environ.get("SORV_FUSES", "0,3,5,7")).split(","):
        os.environ["PF_SOR_FUSE"] = fuse
        ms = C.c_double(); ln = C.c_double()
        rc = L.pf_bench_sor(h, w, nsor, 8, 1, 0, C.byref(ms), C.byref(ln))
        out.append("%dx%d T%s:%6.1f(%d)" % (w, h, fuse, ms.value * 1000 if rc == 0 else -1, int(ln.value)))
print("%-10s " % name + "  ".join(out), flush=True)
