"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/liboracle.so (oracle/pyflow_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.  All arrays are HWC float64, like the
reference's numpy-facing boundary (Par/pyflow.pyx:31-52).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None

LEX, REDBLACK = 0, 1


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        i, d = C.c_int, C.c_double
        L.oracle_filter_h.argtypes = [_dp, _dp, i, i, i, _dp, i]
        L.oracle_filter_v.argtypes = [_dp, _dp, i, i, i, _dp, i]
        L.oracle_gaussian.argtypes = [_dp, _dp, i, i, i, d, i]
        L.oracle_resize_ratio.argtypes = [_dp, _dp, i, i, i, d]
        L.oracle_resize_to.argtypes = [_dp, _dp, i, i, i, i, i, d]
        L.oracle_levels_from_min_width.argtypes = [i, d, i]
        L.oracle_levels_from_min_width.restype = i
        L.oracle_level_geometry.argtypes = [i, i, d, i, _ip, _ip, _ip, _ip, _dp, _dp]
        L.oracle_level_geometry.restype = i
        L.oracle_pyramid.argtypes = [_dp, i, i, i, d, i, _dp]
        L.oracle_im2feature.argtypes = [_dp, _dp, i, i, i, i]
        L.oracle_im2feature.restype = i
        L.oracle_getdxs.argtypes = [_dp, _dp, _dp, _dp, _dp, i, i, i]
        L.oracle_warpfl.argtypes = [_dp, _dp, _dp, _dp, _dp, i, i, i]
        L.oracle_laplacian.argtypes = [_dp, _dp, _dp, i, i]
        L.oracle_est_laplacian_noise.argtypes = [_dp, _dp, i, i, i, _dp]
        L.oracle_sor_solve.argtypes = [_dp] * 8 + [i, i, d, d, i, i]
        L.oracle_assemble.argtypes = [_dp] * 8 + [i, i, i, d] + [_dp] * 6
        L.oracle_smoothflow_sor.argtypes = [_dp] * 6 + [i, i, i, d, i, i, i, i]
        L.oracle_bicubic_warp.argtypes = [_dp] * 5 + [i, i, i]
        L.oracle_coarse2fine_flow.argtypes = [_dp] * 5 + [d, d, i, i, i, i, i, i, i, i, i, i]
        L.oracle_coarse2fine_flow.restype = i
        L.oracle_set_variant.argtypes = [i, i]
        L.oracle_set_variant.restype = None
        L.oracle_gm_get.argtypes = [_dp, _dp, _dp, i]
        L.oracle_gm_get.restype = None
        L.oracle_gm_reset.restype = None
        L.oracle_est_gaussian_mixture.argtypes = [_dp, _dp, i, i, i]
        L.oracle_est_gaussian_mixture.restype = None
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def set_variant(interpolation="bilinear", noise_model="lap"):
    """Process-global switch like the reference's statics OpticalFlow::interpolation / noiseModel
    (S/OpticalFlow.h:19-27)."""
    lib().oracle_set_variant({"bilinear": 0, "bicubic": 1}[interpolation], {"gmixture": 0, "lap": 1}[noise_model])


def gm_get(c):
    a = np.zeros(c); s = np.zeros(c); b = np.zeros(c)
    lib().oracle_gm_get(_p(a), _p(s), _p(b), c)
    return a, s, b


def est_gaussian_mixture(im1, im2, reset=True):
    """OpticalFlow::estGaussianMixture (S/OpticalFlow.cpp:539-591) from freshly reset parameters."""
    im1, im2 = _hwc(im1), _hwc(im2)
    h, w, c = im1.shape
    if reset:
        lib().oracle_gm_reset()
    lib().oracle_est_gaussian_mixture(_p(im1), _p(im2), w, h, c)
    return gm_get(c)


def _p(a):
    return a.ctypes.data_as(_dp)


def _hwc(a):
    a = _c(a)
    return a[..., None] if a.ndim == 2 else a


def filter_h(src, taps):
    src = _hwc(src); taps = _c(taps); h, w, c = src.shape
    out = np.zeros_like(src)
    lib().oracle_filter_h(_p(src), _p(out), w, h, c, _p(taps), len(taps) // 2)
    return out


def filter_v(src, taps):
    src = _hwc(src); taps = _c(taps); h, w, c = src.shape
    out = np.zeros_like(src)
    lib().oracle_filter_v(_p(src), _p(out), w, h, c, _p(taps), len(taps) // 2)
    return out


def gaussian(src, sigma, fsize):
    src = _hwc(src); h, w, c = src.shape
    out = np.zeros_like(src)
    lib().oracle_gaussian(_p(src), _p(out), w, h, c, sigma, fsize)
    return out


def resize_ratio(src, ratio):
    src = _hwc(src); h, w, c = src.shape
    dw, dh = int(float(w) * ratio), int(float(h) * ratio)
    out = np.zeros((dh, dw, c))
    lib().oracle_resize_ratio(_p(src), _p(out), w, h, c, ratio)
    return out


def resize_to(src, dh, dw, scale=1.0):
    src = _hwc(src); h, w, c = src.shape
    out = np.zeros((dh, dw, c))
    lib().oracle_resize_to(_p(src), _p(out), w, h, c, dw, dh, scale)
    return out


def levels_from_min_width(width, ratio, min_width):
    return lib().oracle_levels_from_min_width(width, ratio, min_width)


def level_geometry(w0, h0, ratio, nlevels):
    ws = (C.c_int * 64)(); hs = (C.c_int * 64)(); sl = (C.c_int * 64)(); fs = (C.c_int * 64)()
    sg = (C.c_double * 64)(); rt = (C.c_double * 64)()
    lib().oracle_level_geometry(w0, h0, ratio, nlevels, ws, hs, sl, fs, sg, rt)
    return [dict(w=ws[k], h=hs[k], src=sl[k], fsize=fs[k], sigma=sg[k], rate=rt[k])
            for k in range(nlevels)]


def pyramid(im, ratio, nlevels):
    im = _hwc(im); h, w, c = im.shape
    geo = level_geometry(w, h, ratio, nlevels)
    total = sum(g["w"] * g["h"] * c for g in geo)
    buf = np.zeros(total)
    lib().oracle_pyramid(_p(im), w, h, c, ratio, nlevels, _p(buf))
    out, off = [], 0
    for g in geo:
        n = g["w"] * g["h"] * c
        out.append(buf[off:off + n].reshape(g["h"], g["w"], c).copy())
        off += n
    return out


def im2feature(im, swap_luma=0):
    im = _hwc(im); h, w, c = im.shape
    fc = {1: 3, 3: 5}.get(c, c)
    out = np.zeros((h, w, fc))
    lib().oracle_im2feature(_p(im), _p(out), w, h, c, swap_luma)
    return out


def getdxs(im1, im2):
    im1, im2 = _hwc(im1), _hwc(im2); h, w, c = im1.shape
    dx, dy, dt = np.zeros_like(im1), np.zeros_like(im1), np.zeros_like(im1)
    lib().oracle_getdxs(_p(dx), _p(dy), _p(dt), _p(im1), _p(im2), w, h, c)
    return dx, dy, dt


def warpfl(im1, im2, vx, vy):
    im1, im2, vx, vy = _hwc(im1), _hwc(im2), _c(vx), _c(vy); h, w, c = im1.shape
    out = np.zeros_like(im1)
    lib().oracle_warpfl(_p(out), _p(im1), _p(im2), _p(vx), _p(vy), w, h, c)
    return out


def laplacian(x, weight):
    x, weight = _c(x), _c(weight); h, w = x.shape
    out = np.zeros_like(x)
    lib().oracle_laplacian(_p(out), _p(x), _p(weight), w, h)
    return out


def est_laplacian_noise(im1, im2):
    im1, im2 = _hwc(im1), _hwc(im2); h, w, c = im1.shape
    para = np.zeros(16)
    lib().oracle_est_laplacian_noise(_p(im1), _p(im2), w, h, c, _p(para))
    return para[:c]


def sor_solve(phi, dxy, dx2, dy2, bu, bv, alpha, nsor, order=LEX, omega=1.8, du=None, dv=None):
    phi, dxy, dx2, dy2, bu, bv = map(_c, (phi, dxy, dx2, dy2, bu, bv)); h, w = phi.shape
    du = np.zeros((h, w)) if du is None else _c(du).copy()
    dv = np.zeros((h, w)) if dv is None else _c(dv).copy()
    lib().oracle_sor_solve(_p(du), _p(dv), _p(phi), _p(dxy), _p(dx2), _p(dy2), _p(bu), _p(bv), w,
                           h, alpha, omega, nsor, order)
    return du, dv


def assemble(imdx, imdy, imdt, u, v, du, dv, lap, alpha):
    imdx, imdy, imdt = _hwc(imdx), _hwc(imdy), _hwc(imdt); h, w, c = imdx.shape
    u, v, du, dv = map(_c, (u, v, du, dv))
    lap16 = np.zeros(16); lap16[:len(lap)] = lap
    outs = [np.zeros((h, w)) for _ in range(6)]
    lib().oracle_assemble(_p(imdx), _p(imdy), _p(imdt), _p(u), _p(v), _p(du), _p(dv), _p(lap16),
                          w, h, c, alpha, *[_p(o) for o in outs])
    return dict(zip(("phi", "dxy", "dx2", "dy2", "bu", "bv"), outs))


def smoothflow_sor(f1, f2, warp, u, v, alpha, n_outer, n_inner, n_sor, lap=None, order=LEX):
    f1, f2 = _hwc(f1), _hwc(f2); h, w, c = f1.shape
    warp, u, v = _hwc(warp).copy(), _c(u).copy(), _c(v).copy()
    lap16 = np.full(16, 0.02) if lap is None else _c(lap).copy()
    lib().oracle_smoothflow_sor(_p(f1), _p(f2), _p(warp), _p(u), _p(v), _p(lap16), w, h, c, alpha,
                                n_outer, n_inner, n_sor, order)
    return warp, u, v


def bicubic_warp(ref, im2, vx, vy):
    ref, im2, vx, vy = _hwc(ref), _hwc(im2), _c(vx), _c(vy); h, w, c = im2.shape
    out = np.zeros_like(im2)
    lib().oracle_bicubic_warp(_p(out), _p(ref), _p(im2), _p(vx), _p(vy), w, h, c)
    return out


def flow_encode_u16(flow):
    """S/OpticalFlow.cpp:993-1003 on any float64 array -> uint16 of the same shape."""
    flow = _c(flow)
    q = np.zeros(flow.shape, dtype=np.uint16)
    lib().oracle_flow_encode_u16(q.ctypes.data_as(C.POINTER(C.c_ushort)), _p(flow), C.c_long(flow.size))
    return q


def flow_decode_u16(q):
    """S/OpticalFlow.cpp:962-975."""
    q = np.ascontiguousarray(q, dtype=np.uint16)
    flow = np.zeros(q.shape)
    lib().oracle_flow_decode_u16(_p(flow), q.ctypes.data_as(C.POINTER(C.c_ushort)), C.c_long(q.size))
    return flow


def coarse2fine_flow(im1, im2, alpha=0.012, ratio=0.75, minWidth=20, nOuter=7, nInner=1, nSOR=30,
                     colType=0, levels=0, order=LEX):
    """Upstream call shape (north_star); pass levels>0 for the fork's pyramidLevels shape."""
    im1, im2 = _hwc(im1), _hwc(im2); h, w, c = im1.shape
    vx, vy, wi = np.zeros((h, w)), np.zeros((h, w)), np.zeros((h, w, c))
    n = lib().oracle_coarse2fine_flow(_p(vx), _p(vy), _p(wi), _p(im1), _p(im2), alpha, ratio,
                                      minWidth, levels, nOuter, nInner, nSOR, colType, h, w, c,
                                      order)
    if n <= 0:
        raise ValueError("oracle: invalid level count for width %d / minWidth %d" % (w, minWidth))
    return vx, vy, wi
