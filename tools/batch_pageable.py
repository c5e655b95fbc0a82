"""Developer tool: end-to-end batch throughput with PAGEABLE numpy buffers (what a Python caller passes), stager on/off.
usage: python tools/batch_pageable.py [npairs]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyflow
from conftest import load_frame
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
fr = [load_frame(1920, i) for i in (1, 2, 3)]
pairs = [(fr[i % 2], fr[i % 2 + 1]) for i in range(n)]
outs = [(np.zeros((1080, 1920)), np.zeros((1080, 1920)), np.zeros((1080, 1920, 3))) for _ in range(n)]
for o in outs:
    for a in o: a.fill(0)          # touch the pages: a real caller reuses or has written its buffers
pyflow.coarse2fine_flow_batch(pairs[:16], 0.012, 0.75, 20, 7, 1, 30, 0, outs=outs[:16])
t = time.perf_counter()
pyflow.coarse2fine_flow_batch(pairs, 0.012, 0.75, 20, 7, 1, 30, 0, outs=outs)
dt = time.perf_counter() - t
print("PF_STAGER=%s  batch of %d pairs from pageable buffers: %.1f pairs/s" % (os.environ.get("PF_STAGER", "1"), n, n / dt), flush=True)
