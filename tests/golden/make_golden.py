"""Generates tests/golden/* by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref, Serial build).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
Outputs (committed):
  frames/hcm<w>_<idx>.jpg      a few input frames of the reference's own test data
                               (images_New/HoChiMinhTraffic_10FPS_<w>), byte copies, decoded with PIL
                               exactly as the reference driver does (Par/OpticalFlowCalculation.py:66-71)
  hcm240_L8.npz                full (vx, vy, warpI2) float64, fork call shape pyramidLevels=8
  hcm240_p30_31_L8.npz         same for frames 30->31
  hcm240_gray_params.npz       gray (h,w,1) input through the parameterised driver with non-default
                               alpha/ratio/minWidth/iterations (exercises colType=1 and nInner=2)
  hcm480_L11_s4.npz, hcm960_L13_s8.npz, hcm1920_L15_s8.npz
                               strided subsamples of the outputs + full-array sums (the full arrays are
                               too large to commit; GPU tests also re-run oracle/_ref live when present)
  stages_96x64.npz             per-stage dumps through the reference's public static functions
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

SRC = "/root/reference/images_New/HoChiMinhTraffic_10FPS_%d/frame_%05d.jpg"
FRAMES = {240: [1, 2, 30, 31], 480: [1, 2], 960: [1, 2], 1920: [1, 2, 3]}


def load(w, idx):
    return np.array(Image.open(os.path.join(HERE, "frames", "hcm%d_%05d.jpg" % (w, idx)))).astype(float) / 255.


def sub(a, s):
    return np.ascontiguousarray(a[::s, ::s])


def main():
    os.makedirs(os.path.join(HERE, "frames"), exist_ok=True)
    for w, idxs in FRAMES.items():
        for i in idxs:
            dst = os.path.join(HERE, "frames", "hcm%d_%05d.jpg" % (w, i))
            if not os.path.exists(dst):
                shutil.copyfile(SRC % (w, i), dst)
                os.chmod(dst, 0o644)
    r = ref.serial()

    a, b = load(240, 1), load(240, 2)
    _, vx, vy, wi = r.coarse2fine_flow_levels(a, b, 8)
    np.savez_compressed(os.path.join(HERE, "hcm240_L8.npz"), vx=vx, vy=vy, warpI2=wi)
    a, b = load(240, 30), load(240, 31)
    _, vx, vy, wi = r.coarse2fine_flow_levels(a, b, 8)
    np.savez_compressed(os.path.join(HERE, "hcm240_p30_31_L8.npz"), vx=vx, vy=vy, warpI2=wi)

    # gray, non-default parameters, two inner iterations
    a, b = load(240, 1), load(240, 2)
    ga = (a @ np.array([.299, .587, .114]))[..., None]
    gb = (b @ np.array([.299, .587, .114]))[..., None]
    params = dict(alpha=0.02, ratio=0.6, minWidth=30, nOuter=3, nInner=2, nSOR=12, colType=1)
    vx, vy, wi = r.coarse2fine_flow(ga, gb, **params)
    np.savez_compressed(os.path.join(HERE, "hcm240_gray_params.npz"), vx=vx, vy=vy, warpI2=wi,
                        im1=ga, im2=gb, **params)

    for w, lv, s in ((480, 11, 4), (960, 13, 8), (1920, 15, 8)):
        a, b = load(w, 1), load(w, 2)
        t = time.time()
        tm, vx, vy, wi = r.coarse2fine_flow_levels(a, b, lv)
        print("ref %dw L%d: %.1fs" % (w, lv, time.time() - t), tm.get("Phase5_SOR"))
        np.savez_compressed(os.path.join(HERE, "hcm%d_L%d_s%d.npz" % (w, lv, s)), stride=s,
                            vx=sub(vx, s), vy=sub(vy, s), warpI2=sub(wi, s),
                            sums=np.array([vx.sum(), vy.sum(), wi.sum()]),
                            abs_sums=np.array([np.abs(vx).sum(), np.abs(vy).sum(), np.abs(wi).sum()]),
                            minmax=np.array([vx.min(), vx.max(), vy.min(), vy.max()]),
                            ref_seconds=float(tm["Total C++ Execution"]),
                            ref_sor_seconds=float(tm["Phase5_SOR"]))

    # per-stage dumps on a 96x64 crop of the 240w pair
    a, b = load(240, 1)[30:94, 100:196], load(240, 2)[30:94, 100:196]
    pyr = r.pyramid(a, ratio=0.75, levels=5)
    f1, f2 = r.im2feature(a), r.im2feature(b)
    dx, dy, dt = r.getdxs(f1, f2)
    rng = np.random.default_rng(7)
    u = rng.normal(size=a.shape[:2]) * 1.5
    v = rng.normal(size=a.shape[:2]) * 1.5
    wgt = rng.random(a.shape[:2]) + 0.1
    warp = r.warpfl(f1, f2, u, v)
    lap = r.laplacian(u, wgt)
    up = r.resize_to(u, 85, 128, 1 / 0.75)
    bic = r.bicubic(a, b, u, v)
    w0, u1, v1 = r.smoothflow_sor(f1, f2, f2, np.zeros_like(u), np.zeros_like(v), 0.012, 3, 1, 10, 3)
    out = dict(im1=a, im2=b, f1=f1, f2=f2, imdx=dx, imdy=dy, imdt=dt, u=u, v=v, wgt=wgt, warp=warp,
               lap=lap, up=up, bicubic=bic, sor_warp=w0, sor_u=u1, sor_v=v1)
    for k, p in enumerate(pyr):
        out["pyr%d" % k] = p
    np.savez_compressed(os.path.join(HERE, "stages_96x64.npz"), **out)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
