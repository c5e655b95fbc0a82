"""Golden fixture for the reference's flow file format, written by the REFERENCE's own
OpticalFlow::SaveOpticalFlow (oracle/_ref, built from /root/reference; run in the build container):
  tests/golden/flowfile_37x29.npz  -- the float64 input flow (with out-of-range, NaN-free edge values)
  tests/golden/flowfile_37x29.bin  -- the file the reference wrote for it (type tag bytes 1..15 zeroed: the
                                      reference leaves them uninitialised)
usage: python tests/golden/make_golden_flowfile.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref

h, w = 29, 37
rng = np.random.default_rng(7)
flow = rng.normal(0, 6, (h, w, 2))
flow[0, :5, 0] = [-250.0, -200.0, 200.0, 250.0, 0.0]          # clamp edges
flow[1, :4, 1] = [199.99999, -199.99999, 1 / 160, -1 / 160]   # truncation edges
flow[2, :3, 0] = [0.003124, 0.003126, 12.345678]
here = os.path.join(ROOT, "tests", "golden")
path = os.path.join(here, "flowfile_37x29.bin")
R = ref.serial()
R.save_optical_flow(flow, path)
raw = bytearray(open(path, "rb").read())
raw[1:16] = bytes(15)
open(path, "wb").write(bytes(raw))
back = R.load_optical_flow(path, h, w)
np.savez_compressed(os.path.join(here, "flowfile_37x29.npz"), flow=flow, decoded=back)
print("wrote", path, len(raw), "bytes; tag byte", raw[0:1])
