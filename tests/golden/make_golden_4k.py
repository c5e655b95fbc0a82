"""Golden fixture for BASELINE config 5 (3840x2160 gray, colType=1, nSORIterations=60, 18 levels):
runs the UNMODIFIED reference through the parameterised harness (about 5 minutes, one core) and
stores a stride-16 subsample of (vx, vy, warpI2) plus checksums.  python tests/golden/make_golden_4k.py"""
import os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
from oracle import ref
from synth4k import make

im1, im2, gu, gv = make()
t = time.time()
vx, vy, wi = ref.serial().coarse2fine_flow(im1, im2, 0.012, 0.75, 20, 7, 1, 60, 1)
secs = time.time() - t
s = 16
epe = np.hypot(vx - gu, vy - gv)
print("reference 4K: %.1fs, EPE vs ground truth mean %.4f" % (secs, epe.mean()))
np.savez_compressed(os.path.join(HERE, "synth4k_L18_sor60_s16.npz"), stride=s, vx=vx[::s, ::s].copy(), vy=vy[::s, ::s].copy(),
                    warpI2=wi[::s, ::s].copy(), sums=np.array([vx.sum(), vy.sum(), wi.sum()]), ref_seconds=secs,
                    in_sums=np.array([im1.sum(), im2.sum()]), gt_epe_mean=epe.mean())
