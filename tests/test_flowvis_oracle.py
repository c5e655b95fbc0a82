"""oracle/flowvis.py (numpy restatement of the reference driver's flow visualisation)
against golden vectors made by the driver's own cv2 statements, and against cv2 itself when it is importable."""
import numpy as np
import pytest

from conftest import golden


def _check_bgr(got, want):
    w = want.shape[1]
    body = w & ~63          # cv2's scalar row tail (w mod 16/32/64 columns, build dependent) rounds instead of truncating
    assert np.array_equal(got[:, :body], want[:, :body])
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1


def test_flowvis_oracle_matches_cv2_golden():
    from oracle import flowvis as fv
    g = golden("flowvis.npz")
    for name in ("a", "b", "c", "zero"):
        flow = g["flow_" + name].astype(np.float64)
        assert np.array_equal(fv.flow_to_hsv(flow), g["hsv_" + name]), name
        _check_bgr(fv.flow_to_bgr(flow), g["bgr_" + name])
    hh, vv = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    hsv = np.zeros((256, 256, 3), np.uint8); hsv[..., 0] = hh; hsv[..., 1] = 255; hsv[..., 2] = vv
    assert np.array_equal(fv.hsv_to_bgr_u8(hsv), g["lut_s255"])


def test_flowvis_oracle_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    from oracle import flowvis as fv
    rng = np.random.default_rng(2)
    for h, w, sd in ((40, 192, 2.0), (37, 131, 0.2), (64, 64, 40.0)):
        flow = (rng.normal(0, sd, (h, w, 2))).astype(np.float32).astype(np.float64)
        mag, ang = cv2.cartToPolar(flow[..., 0], flow[..., 1])
        m2, a2 = fv.cart_to_polar(flow[..., 0], flow[..., 1])
        assert np.array_equal(mag, m2) and np.array_equal(ang, a2)
        hsv = np.zeros((h, w, 3), np.uint8); hsv[..., 1] = 255
        hsv[..., 0] = ang * 180 / np.pi / 2
        hsv[..., 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX)
        assert np.array_equal(fv.flow_to_hsv(flow), hsv)
        _check_bgr(fv.flow_to_bgr(flow), cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR))
