"""ctypes helpers for the single-stage C-ABI entry points (GPU tests only)."""
import numpy as np

from papteam_opticalflow_b200 import _lib
from papteam_opticalflow_b200._lib import check, dp

F64, F32, F64RB, F32LEX = 0, 1, 2, 3


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _hwc(a):
    a = _c(a)
    return a[..., None] if a.ndim == 2 else a


def _p(a):
    return a.ctypes.data_as(dp) if a is not None else None


def pyramid(im, ratio, levels, mode):
    im = _hwc(im); h, w, c = im.shape
    import ctypes as C
    ws = (C.c_int * 64)(); hs = (C.c_int * 64)()
    check(_lib.lib().pf_level_geometry(w, h, ratio, levels, ws, hs))
    buf = np.zeros(sum(ws[k] * hs[k] * c for k in range(levels)))
    check(_lib.lib().pf_stage_pyramid(_p(buf), _p(im), h, w, c, ratio, levels, mode, 0))
    out, off = [], 0
    for k in range(levels):
        n = ws[k] * hs[k] * c
        out.append(buf[off:off + n].reshape(hs[k], ws[k], c).copy()); off += n
    return out


def im2feature(im, mode, swap=0):
    im = _hwc(im); h, w, c = im.shape
    out = np.zeros((h, w, {1: 3, 3: 5}.get(c, c)))
    check(_lib.lib().pf_stage_im2feature(_p(out), _p(im), h, w, c, swap, mode, 0))
    return out


def getdxs(a, b, mode):
    a, b = _hwc(a), _hwc(b); h, w, c = a.shape
    o = [np.zeros_like(a) for _ in range(3)]
    check(_lib.lib().pf_stage_getdxs(_p(o[0]), _p(o[1]), _p(o[2]), _p(a), _p(b), h, w, c, mode, 0))
    return o


def warpfl(a, b, u, v, mode):
    a, b, u, v = _hwc(a), _hwc(b), _c(u), _c(v); h, w, c = a.shape
    o = np.zeros_like(a)
    check(_lib.lib().pf_stage_warpfl(_p(o), _p(a), _p(b), _p(u), _p(v), h, w, c, mode, 0))
    return o


def resize_to(src, dh, dw, scale, mode):
    src = _hwc(src); h, w, c = src.shape
    o = np.zeros((dh, dw, c))
    check(_lib.lib().pf_stage_resize_to(_p(o), _p(src), h, w, c, dh, dw, scale, mode, 0))
    return o


def bicubic(ref, im2, u, v, mode):
    ref, im2, u, v = _hwc(ref), _hwc(im2), _c(u), _c(v); h, w, c = im2.shape
    o = np.zeros_like(im2)
    check(_lib.lib().pf_stage_bicubic(_p(o), _p(ref), _p(im2), _p(u), _p(v), h, w, c, mode, 0))
    return o


def assemble(imdx, imdy, imdt, u, v, du, dv, lap, alpha, mode):
    imdx, imdy, imdt = _hwc(imdx), _hwc(imdy), _hwc(imdt); h, w, c = imdx.shape
    u, v = _c(u), _c(v)
    du = None if du is None else _c(du)
    dv = None if dv is None else _c(dv)
    lap = None if lap is None else _c(lap)
    o = [np.zeros((h, w)) for _ in range(6)]
    check(_lib.lib().pf_stage_assemble(*[_p(x) for x in o], _p(imdx), _p(imdy), _p(imdt), _p(u), _p(v),
                                       _p(du), _p(dv), _p(lap), alpha, h, w, c, mode, 0))
    return dict(zip(("phi", "dxy", "dx2", "dy2", "bu", "bv"), o))


def sor(phi, dxy, dx2, dy2, bu, bv, alpha, nsor, mode):
    phi, dxy, dx2, dy2, bu, bv = map(_c, (phi, dxy, dx2, dy2, bu, bv)); h, w = phi.shape
    du, dv = np.zeros((h, w)), np.zeros((h, w))
    check(_lib.lib().pf_stage_sor(_p(du), _p(dv), _p(phi), _p(dxy), _p(dx2), _p(dy2), _p(bu), _p(bv),
                                  alpha, nsor, h, w, mode, 0))
    return du, dv
