"""BASELINE config 4 spot checks: pairs 50 and 101 of images_New/HoChiMinhTraffic_10FPS_1920 (pair p = frames
p -> p+1, the pairing rule of Par/InputCreation/TestImagePairGenerator.py:151-171) through the UNMODIFIED
reference (oracle/_ref, Serial build), stride-8 subsamples + full-array sums.  Pair 1 is hcm1920_L15_s8.npz
(make_golden.py).

Run in the build container (needs /root/reference):   python tests/golden/make_golden_config4.py
"""
import os
import shutil
import sys
import time

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

SRC = "/root/reference/images_New/HoChiMinhTraffic_10FPS_1920/frame_%05d.jpg"


def load(idx):
    return np.array(Image.open(os.path.join(HERE, "frames", "hcm1920_%05d.jpg" % idx))).astype(float) / 255.


def main():
    r = ref.serial()
    s = 8
    for p in (50, 101):
        for i in (p, p + 1):
            dst = os.path.join(HERE, "frames", "hcm1920_%05d.jpg" % i)
            if not os.path.exists(dst):
                shutil.copyfile(SRC % i, dst)
                os.chmod(dst, 0o644)
        a, b = load(p), load(p + 1)
        t = time.time()
        tm, vx, vy, wi = r.coarse2fine_flow_levels(a, b, 15)
        print("ref pair %d: %.1fs" % (p, time.time() - t))
        np.savez_compressed(os.path.join(HERE, "hcm1920_p%d_L15_s8.npz" % p), stride=s,
                            vx=np.ascontiguousarray(vx[::s, ::s]), vy=np.ascontiguousarray(vy[::s, ::s]),
                            warpI2=np.ascontiguousarray(wi[::s, ::s]),
                            sums=np.array([vx.sum(), vy.sum(), wi.sum()]),
                            minmax=np.array([vx.min(), vx.max(), vy.min(), vy.max()]),
                            ref_seconds=float(tm["Total C++ Execution"]))


if __name__ == "__main__":
    main()
