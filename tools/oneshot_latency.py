import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, pyflow
from conftest import load_frame
a, b = load_frame(1920, 1), load_frame(1920, 2)
for i in range(6):
    t = time.perf_counter(); r = pyflow.coarse2fine_flow(a, b, 0.012, 0.75, 20, 7, 1, 30, 0); dt = time.perf_counter() - t
    print("one-shot upstream call %d: %.1f ms" % (i, dt * 1e3), flush=True)
for i in range(3):
    t = time.perf_counter(); r = pyflow.coarse2fine_flow(a, b, 15, 1); dt = time.perf_counter() - t
    print("one-shot fork call %d: %.1f ms  (dict total %s)" % (i, dt * 1e3, r[0].get("Total C++ Execution")), flush=True)
