"""Golden vectors of the alternative solver branches, made by RUNNING THE UNMODIFIED
REFERENCE (oracle/_ref, Serial build) with its two public statics assigned through the harness
(OpticalFlow::interpolation / OpticalFlow::noiseModel, S/OpticalFlow.h:19-27).

Run in the build container (needs /root/reference):   python tests/golden/make_golden_variants.py
Output (committed): variants_128x96.npz
  input        rows 0..95, columns 0..127 of HoChiMinhTraffic_10FPS_240 frames 1 and 2
  bicubic_*    interpolation = Bicubic, noiseModel = Lap: fork entry (4 levels) and upstream-shaped entry
               (minWidth 20 -> 6 levels, 5/1/20 iterations)
  gmix_*       noiseModel = GMixture.  The reference's mixture branch is numerically unstable (flows of +-45 px on
               this 128-px crop with the default 7/1/30 iterations; a 1e-13 input perturbation grows to 0.6 px within
               4 levels x 3 outer iterations), so the vectors use short horizons on which parity is meaningful:
               1 level x 3 outer iterations and 2 levels x 2 outer iterations, 10 SOR sweeps; mixture parameters
               left by the last estGaussianMixture call are recorded too.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
from conftest import load_frame  # noqa: E402


def crop():
    a, b = load_frame(240, 1), load_frame(240, 2)
    return np.ascontiguousarray(a[:96, :128]), np.ascontiguousarray(b[:96, :128])


def main():
    r = ref.serial()
    a, b = crop()
    out = {}
    try:
        r.set_variant("bicubic", "lap")
        _, vx, vy, wi = r.coarse2fine_flow_levels(a, b, 4)
        out.update(bicubic_fork_vx=vx, bicubic_fork_vy=vy, bicubic_fork_warp=wi)
        vx, vy, wi = r.coarse2fine_flow(a, b, 0.012, 0.75, 20, 5, 1, 20, 0)
        out.update(bicubic_up_vx=vx, bicubic_up_vy=vy, bicubic_up_warp=wi)
        # gray through the upstream-shaped entry
        ag, bg = np.ascontiguousarray(a.mean(axis=2, keepdims=True)), np.ascontiguousarray(b.mean(axis=2, keepdims=True))
        vx, vy, wi = r.coarse2fine_flow(ag, bg, 0.012, 0.75, 20, 4, 2, 15, 1)
        out.update(bicubic_gray_vx=vx, bicubic_gray_vy=vy, bicubic_gray_warp=wi)
        for interp in ("bilinear", "bicubic"):
            r.set_variant(interp, "gmixture")
            for tag, mw, no in (("l1o3", 90, 3), ("l2o2", 70, 2)):
                # minWidth 90 -> 1 level, 70 -> 2 levels on a 128-wide image (S/GaussianPyramid.cpp:53)
                vx, vy, wi = r.coarse2fine_flow(a, b, 0.012, 0.75, mw, no, 1, 10, 0)
                al, sg, be = r.gm_get(5)
                k = "gmix_%s_%s_" % (interp, tag)
                out.update({k + "vx": vx, k + "vy": vy, k + "alpha": al, k + "sigma": sg, k + "beta": be})
    finally:
        r.set_variant("bilinear", "lap")
    np.savez_compressed(os.path.join(HERE, "variants_128x96.npz"), **out)
    for k in sorted(out):
        print(k, out[k].shape, float(np.abs(out[k]).max()))


if __name__ == "__main__":
    main()
