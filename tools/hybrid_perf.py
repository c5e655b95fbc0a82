"""fp32_hybrid on the headline pair: single-pair latency and resident throughput with 24 pairs in flight (bench.py's legs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import pyflow
from conftest import load_frame
a, b = load_frame(1920, 1), load_frame(1920, 2)
for mode in sys.argv[1:] or ["fp32_hybrid", "fp32_redblack"]:
    lat = pyflow.FlowPlan(1080, 1920, 3, mode=mode, tuning="latency")
    lat.upload(a, b); lat.solve(2)
    ms1 = lat.solve(4) / 4
    lat.close()
    B = 24
    plans = [pyflow.FlowPlan(1080, 1920, 3, mode=mode) for _ in range(B)]
    for p in plans:
        p.upload(a, b)
    pyflow.multi_solve(plans, 1)
    ms = pyflow.multi_solve(plans, 3)
    for p in plans:
        p.close()
    print("%s: single pair %.2f ms; %d in flight: %.1f pairs/s" % (mode, ms1, B, 3 * B / (ms / 1000.0)), flush=True)
