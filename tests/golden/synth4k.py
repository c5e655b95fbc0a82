"""BASELINE config 5 input: synthetic 3840x2160 gray pair with known affine motion.
Fixed generator (seed 0): 4096x2304 canvas of uniform noise summed over 4 octaves (Gaussian blurred
sigma = 1,2,4,8 px, weights 1, 1/2, 1/4, 1/8), normalised to [0,1]; im1 = central crop; im2 = the same
canvas resampled (bicubic, from the oversize canvas) under x' = 1.002x + 0.003y + 2.5,
y' = -0.003x + 0.998y - 1.5 about the image centre.  Returns (im1, im2, gt_u, gt_v) with images (h,w,1)."""
import numpy as np


def make(h=2160, w=3840, seed=0):
    from scipy import ndimage
    ch, cw = h + 144, w + 256
    rng = np.random.default_rng(seed)
    canvas = np.zeros((ch, cw))
    for s, wt in ((1, 1.0), (2, 0.5), (4, 0.25), (8, 0.125)):
        canvas += wt * ndimage.gaussian_filter(rng.random((ch, cw)), s, mode="wrap")
    canvas = (canvas - canvas.min()) / (canvas.max() - canvas.min())
    oy, ox = (ch - h) // 2, (cw - w) // 2
    im1 = canvas[oy:oy + h, ox:ox + w]
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    xc, yc = xx - (w - 1) / 2.0, yy - (h - 1) / 2.0
    # im2(x) = im1(x - flow): a point at p in im1 moves to p + flow(p); sample the canvas backwards
    a = np.array([[1.002, 0.003], [-0.003, 0.998]]); t = np.array([2.5, -1.5])
    ai = np.linalg.inv(a)
    sx = ai[0, 0] * (xc - t[0]) + ai[0, 1] * (yc - t[1]) + (w - 1) / 2.0
    sy = ai[1, 0] * (xc - t[0]) + ai[1, 1] * (yc - t[1]) + (h - 1) / 2.0
    im2 = ndimage.map_coordinates(canvas, [sy + oy, sx + ox], order=3, mode="nearest")
    gt_u = (a[0, 0] - 1) * xc + a[0, 1] * yc + t[0]
    gt_v = a[1, 0] * xc + (a[1, 1] - 1) * yc + t[1]
    return (np.ascontiguousarray(np.clip(im1, 0, 1)[..., None]), np.ascontiguousarray(np.clip(im2, 0, 1)[..., None]),
            gt_u, gt_v)
