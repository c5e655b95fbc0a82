"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference driver's flow visualisation
(SURVEY.md 8f row f1), Par/OpticalFlowCalculation.py:143-162 (generateOutputFlowImageFile):

    hsv = zeros(imDimensions, uint8); hsv[..., 0] = 255; hsv[..., 1] = 255
    mag, ang = cv2.cartToPolar(flow[..., 0], flow[..., 1])
    hsv[..., 0] = ang * 180 / np.pi / 2
    hsv[..., 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX)
    rgb = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)

The arithmetic lives in a third-party dependency that is not in /root/reference: OpenCV (cv2 4.13.0 in this
image; the reference pins no version).  Its algorithms are restated here and PINNED by running cv2 itself
(tests/test_flowvis_oracle.py, bit-exact on every case tried) and by golden vectors made with it
(tests/golden/make_golden_flowvis.py):
  * cartToPolar on float64 input works in float32: mag = sqrt(fma(x, x, y*y)); angle = fastAtan32f, a degree-
    valued odd polynomial in c = min(|x|,|y|) / (max(|x|,|y|) + (float)DBL_EPSILON) evaluated with FMAs,
    coefficients (float)k * (float)(180/pi), folded into the quadrant, then * (float)(pi/180);
  * normalize(NORM_MINMAX, 0..255) on float64: scale = 255 * (1 / (max - min)) (0 when max - min <= DBL_EPSILON),
    shift = 0 - min * scale, dst = src * scale + shift;
  * assigning float64 to a uint8 array truncates toward zero;
  * 8-bit HSV2BGR (hue range 180): float32 h = H * (6/180), s = S/255, v = V/255, sector = floor(h), f = h - sector,
    tab = [v, v(1-s), v(1-s f), v(1-s(1-f))], (b, g, r) picked per sector, each TRUNCATED from x * 255.
Nothing under papteam_opticalflow_b200/ imports this module."""
import numpy as np

f32, f64 = np.float32, np.float64
_K = f32(180.0 / np.pi)
_P1, _P3 = f32(0.9997878412794807) * _K, f32(-0.3258083974640975) * _K
_P5, _P7 = f32(0.1555786518463281) * _K, f32(-0.04432655554792128) * _K
_EPS = f32(2.220446049250313e-16)
_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])


def _fma32(a, b, c):
    # exact product of two float32 in float64, one rounding of the sum to float64, then to float32
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)


def cart_to_polar(x, y):
    """cv2.cartToPolar(x, y) for float64 (or float32) arrays -> (mag, ang) as float64 holding float32 values."""
    x = np.asarray(x).astype(f32); y = np.asarray(y).astype(f32)
    mag = np.sqrt(_fma32(x, x, (y * y).astype(f32))).astype(f32)
    ax, ay = np.abs(x), np.abs(y)
    with np.errstate(invalid="ignore", divide="ignore"):
        c = (np.minimum(ax, ay) / (np.maximum(ax, ay) + _EPS)).astype(f32)
    c2 = (c * c).astype(f32)
    a = _fma32(c2, _P7, _P5); a = _fma32(a, c2, _P3); a = _fma32(a, c2, _P1); a = (a * c).astype(f32)
    a = np.where(ax >= ay, a, f32(90) - a).astype(f32)
    a = np.where(x < 0, f32(180) - a, a).astype(f32)
    a = np.where(y < 0, f32(360) - a, a).astype(f32)
    ang = (a * f32(np.pi / 180)).astype(f32)
    return mag.astype(f64), ang.astype(f64)


def normalize_minmax_255(src):
    src = np.asarray(src, f64)
    smin, smax = src.min(), src.max()
    scale = 255.0 * (1.0 / (smax - smin) if smax - smin > np.finfo(f64).eps else 0.0)
    shift = 0.0 - smin * scale
    return src * scale + shift


def hsv_to_bgr_u8(hsv):
    h = hsv[..., 0].astype(f32) * f32(6.0 / 180.0)
    s = hsv[..., 1].astype(f32) * f32(1.0 / 255.0)
    v = hsv[..., 2].astype(f32) * f32(1.0 / 255.0)
    h = np.where(h >= 6, h - f32(6), h).astype(f32)       # hue bytes 180..255 wrap once (255*6/180 = 8.5)
    sec = np.floor(h).astype(np.int32)
    f = (h - sec.astype(f32)).astype(f32)
    tab = np.stack([v, v * (f32(1) - s), v * (f32(1) - s * f), v * (f32(1) - s * (f32(1) - f))], 0).astype(f32)
    out = np.zeros(hsv.shape, np.uint8)
    for ch in range(3):
        val = np.take_along_axis(tab, _SECTOR[sec, ch][None], 0)[0]
        out[..., ch] = np.trunc((val * f32(255)).astype(f32)).astype(np.uint8)
    return out


def flow_to_hsv(flow):
    flow = np.asarray(flow, f64)
    mag, ang = cart_to_polar(flow[..., 0], flow[..., 1])
    hsv = np.zeros(flow.shape[:2] + (3,), np.uint8)
    hsv[..., 1] = 255
    hsv[..., 0] = (ang * 180 / np.pi / 2).astype(np.uint8)
    hsv[..., 2] = normalize_minmax_255(mag).astype(np.uint8)
    return hsv


def flow_to_bgr(flow):
    """(h, w, 2) flow -> (h, w, 3) uint8 BGR image, as generateOutputFlowImageFile builds it before imwrite."""
    return hsv_to_bgr_u8(flow_to_hsv(flow))
