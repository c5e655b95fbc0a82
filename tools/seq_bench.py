"""Sequence-mode throughput (uint8 frames in, float32 flows out) vs the pairwise batch call, 1920x1080 RGB."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyflow

def frames_1080(n):
    rng = np.random.default_rng(0)
    base = rng.random((1080 // 8 + 2, 1920 // 8 + 2, 3))
    base = np.kron(base, np.ones((8, 8, 1)))
    out = []
    for t in range(n):
        f = base[t % 8: t % 8 + 1080, (2 * t) % 8: (2 * t) % 8 + 1920]
        out.append(np.ascontiguousarray((f * 255).astype(np.uint8)))
    return out

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65
fr = frames_1080(n)
for streams in ("4", "8", "16"):
    os.environ["PF_BATCH_STREAMS"] = streams
    pyflow.sequence_flow(fr[:streams and int(streams) * 2 + 1], mode="fp32_redblack", devices=[0])
    t0 = time.perf_counter()
    flows, secs = pyflow.sequence_flow(fr, mode="fp32_redblack", devices=[0])
    dt = time.perf_counter() - t0
    print("streams", streams, "sequence: %.1f pairs/s (wall %.3f s, lib %.3f s)" % ((n - 1) / dt, dt, secs), flush=True)
os.environ["PF_BATCH_STREAMS"] = "8"
for output in ("float32", "u16", "bgr8"):
    pyflow.sequence_flow(fr[:17], mode="fp32_redblack", devices=[0], output=output)
    t0 = time.perf_counter()
    pyflow.sequence_flow(fr, mode="fp32_redblack", devices=[0], output=output)
    dt = time.perf_counter() - t0
    print("output %-8s %.1f pairs/s" % (output, (n - 1) / dt), flush=True)
