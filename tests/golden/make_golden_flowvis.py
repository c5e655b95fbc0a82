"""Golden vectors for the flow visualisation, produced by the reference driver's own
statements (Par/OpticalFlowCalculation.py:152-160) running on cv2 (4.13.0 in this image):
  tests/golden/flowvis.npz: for a few float32 flows (stored as float32) the HSV image and the BGR image;
                            plus the full 8-bit HSV->BGR table for S = 255 (hue byte 0..255 x value byte 0..255),
                            taken from a 256-wide image so that every pixel goes through cv2's vector body
                            (its scalar tail -- the last w mod 32 columns of a row on this build -- rounds
                            instead of truncating, see oracle/flowvis.py).
usage: python tests/golden/make_golden_flowvis.py"""
import os
import cv2
import numpy as np

def driver_visualisation(flow, shape3):
    hsv = np.zeros(shape3, dtype=np.uint8)
    hsv[:, :, 0] = 255
    hsv[:, :, 1] = 255
    mag, ang = cv2.cartToPolar(flow[..., 0], flow[..., 1])
    hsv[..., 0] = ang * 180 / np.pi / 2
    hsv[..., 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX)
    return hsv, cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)

here = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(11)
out = {}
cases = {"a": (64, 128, 3.0), "b": (45, 200, 0.4), "c": (33, 77, 25.0), "zero": (8, 64, 0.0)}
for name, (h, w, sd) in cases.items():
    flow = rng.normal(0, 1, (h, w, 2)) * sd
    if name == "a":
        flow[0, :6] = [[1, 0], [0, 1], [-1, 0], [0, -1], [-1, -1e-30], [1e-20, 1e-20]]   # axes, hue wrap, tiny
    f32 = flow.astype(np.float32)
    hsv, bgr = driver_visualisation(f32.astype(np.float64), (h, w, 3))
    out["flow_" + name], out["hsv_" + name], out["bgr_" + name] = f32, hsv, bgr
hh, vv = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
hsv = np.zeros((256, 256, 3), np.uint8); hsv[..., 0] = hh; hsv[..., 1] = 255; hsv[..., 2] = vv
out["lut_s255"] = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)
out["cv2_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(here, "flowvis.npz"), **out)
print("cv2", cv2.__version__, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})
