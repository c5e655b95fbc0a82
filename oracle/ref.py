"""TEST INFRASTRUCTURE ONLY: ctypes loader for the unmodified reference built into oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (papteam_opticalflow_b200/) never does.

`serial()` is the correctness oracle (Code/Serial, deterministic); `parallel()` is the OpenMP
build (racy for nCores>1) and is a timing baseline only.
"""
import contextlib
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a):
    return a.ctypes.data_as(_dp)


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


@contextlib.contextmanager
def quiet_stdout():
    """The reference prints `P[k](t s)` progress to fd 1 (S/OpticalFlow.cpp:787-788,832-834);
    divert fd 1 to stderr for the duration so callers that own stdout (bench.py) stay clean."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


class _Ref:
    def __init__(self, path, parallel):
        self.lib = C.CDLL(path)
        self.is_parallel = parallel
        L = self.lib
        L.ref_coarse2fine_flow_levels.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int,
                                                  C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.ref_coarse2fine_flow_levels.restype = C.c_int
        L.ref_stage_pyramid.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                        C.c_int, _ip, _ip, _dp]
        L.ref_stage_pyramid.restype = C.c_int
        L.ref_set_variant.argtypes = [C.c_int, C.c_int]
        L.ref_set_variant.restype = None
        L.ref_gm_get.argtypes = [_dp, _dp, _dp, C.c_int]
        L.ref_gm_get.restype = None
        if not parallel:
            L.ref_coarse2fine_flow.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double,
                                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_int]
            L.ref_stage_im2feature.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int]
            L.ref_stage_im2feature.restype = C.c_int
            L.ref_stage_getdxs.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int]
            L.ref_stage_warpfl.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int]
            L.ref_stage_laplacian.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int]
            L.ref_stage_resize_to.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_double]
            L.ref_stage_gaussian.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double,
                                             C.c_int]
            L.ref_stage_bicubic.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int]
            L.ref_save_optical_flow.argtypes = [_dp, C.c_int, C.c_int, C.c_char_p]
            L.ref_save_optical_flow.restype = C.c_int
            L.ref_load_optical_flow.argtypes = [_dp, C.c_int, C.c_int, C.c_char_p]
            L.ref_load_optical_flow.restype = C.c_int
            L.ref_stage_smoothflow_sor.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int,
                                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                   C.c_int]

    # -- alternative solver branches: the reference's public statics ------
    def set_variant(self, interpolation="bilinear", noise_model="lap"):
        """OpticalFlow::interpolation / OpticalFlow::noiseModel (S/OpticalFlow.h:19-27)."""
        self.lib.ref_set_variant({"bilinear": 0, "bicubic": 1}[interpolation], {"gmixture": 0, "lap": 1}[noise_model])

    def gm_get(self, c):
        a = np.zeros(c); s = np.zeros(c); b = np.zeros(c)
        self.lib.ref_gm_get(_p(a), _p(s), _p(b), c)
        return a, s, b

    # -- the fork's entry point (pyramidLevels, nCores) ---------------------------------------
    def coarse2fine_flow_levels(self, im1, im2, levels, ncores=1):
        im1, im2 = _c(im1), _c(im2)
        h, w, c = im1.shape
        vx = np.zeros((h, w)); vy = np.zeros((h, w)); wi = np.zeros((h, w, c))
        buf = C.create_string_buffer(4096)
        with quiet_stdout():
            self.lib.ref_coarse2fine_flow_levels(_p(vx), _p(vy), _p(wi), _p(im1), _p(im2),
                                                 levels, ncores, h, w, c, buf, 4096)
        timing = dict(line.split("=", 1) for line in buf.value.decode().splitlines() if "=" in line)
        return timing, vx, vy, wi

    # -- upstream-shaped parameterised entry (Serial only) ------------------------------------
    def coarse2fine_flow(self, im1, im2, alpha=0.012, ratio=0.75, minWidth=20, nOuter=7,
                         nInner=1, nSOR=30, colType=0):
        im1, im2 = _c(im1), _c(im2)
        h, w, c = im1.shape
        vx = np.zeros((h, w)); vy = np.zeros((h, w)); wi = np.zeros((h, w, c))
        with quiet_stdout():
            self.lib.ref_coarse2fine_flow(_p(vx), _p(vy), _p(wi), _p(im1), _p(im2), alpha, ratio,
                                          minWidth, nOuter, nInner, nSOR, colType, h, w, c)
        return vx, vy, wi

    def pyramid(self, im, ratio=0.75, minWidth=None, levels=None):
        im = _c(im)
        h, w, c = im.shape
        use_mw = minWidth is not None
        arg = minWidth if use_mw else levels
        ws = (C.c_int * 64)(); hs = (C.c_int * 64)()
        n = self.lib.ref_stage_pyramid(_p(im), h, w, c, ratio, int(use_mw), arg, ws, hs, None)
        total = sum(ws[k] * hs[k] * c for k in range(n))
        data = np.zeros(total)
        self.lib.ref_stage_pyramid(_p(im), h, w, c, ratio, int(use_mw), arg, ws, hs, _p(data))
        out, off = [], 0
        for k in range(n):
            sz = ws[k] * hs[k] * c
            out.append(data[off:off + sz].reshape(hs[k], ws[k], c).copy())
            off += sz
        return out

    def im2feature(self, im):
        im = _c(im)
        h, w, c = im.shape
        fc = {1: 3, 3: 5}.get(c, c)
        feat = np.zeros((h, w, fc))
        with quiet_stdout():
            self.lib.ref_stage_im2feature(_p(feat), _p(im), h, w, c)
        return feat

    def getdxs(self, im1, im2):
        im1, im2 = _c(im1), _c(im2)
        h, w, c = im1.shape
        dx = np.zeros((h, w, c)); dy = np.zeros((h, w, c)); dt = np.zeros((h, w, c))
        self.lib.ref_stage_getdxs(_p(dx), _p(dy), _p(dt), _p(im1), _p(im2), h, w, c)
        return dx, dy, dt

    def warpfl(self, im1, im2, vx, vy):
        im1, im2, vx, vy = _c(im1), _c(im2), _c(vx), _c(vy)
        h, w, c = im1.shape
        out = np.zeros((h, w, c))
        self.lib.ref_stage_warpfl(_p(out), _p(im1), _p(im2), _p(vx), _p(vy), h, w, c)
        return out

    def laplacian(self, x, weight):
        x, weight = _c(x), _c(weight)
        h, w = x.shape
        out = np.zeros((h, w))
        self.lib.ref_stage_laplacian(_p(out), _p(x), _p(weight), h, w)
        return out

    def resize_to(self, src, dh, dw, scale=1.0):
        src = _c(src)
        if src.ndim == 2:
            src = src[..., None]
        h, w, c = src.shape
        out = np.zeros((dh, dw, c))
        self.lib.ref_stage_resize_to(_p(out), _p(src), h, w, c, dh, dw, scale)
        return out

    def gaussian(self, src, sigma, fsize):
        src = _c(src)
        h, w, c = src.shape
        out = np.zeros((h, w, c))
        self.lib.ref_stage_gaussian(_p(out), _p(src), h, w, c, sigma, fsize)
        return out

    def save_optical_flow(self, flow, path):
        """OpticalFlow::SaveOpticalFlow (S/OpticalFlow.cpp:993-1003) on an (h, w, 2) float64 flow."""
        flow = _c(flow)
        h, w, _ = flow.shape
        if not self.lib.ref_save_optical_flow(_p(flow), h, w, path.encode()):
            raise IOError("reference SaveOpticalFlow failed")

    def load_optical_flow(self, path, h, w):
        """OpticalFlow::LoadOpticalFlow (S/OpticalFlow.cpp:962-975) -> (h, w, 2) float64."""
        flow = np.zeros((h, w, 2))
        if not self.lib.ref_load_optical_flow(_p(flow), h, w, path.encode()):
            raise IOError("reference LoadOpticalFlow failed")
        return flow

    def bicubic(self, ref, im2, vx, vy):
        ref, im2, vx, vy = _c(ref), _c(im2), _c(vx), _c(vy)
        h, w, c = im2.shape
        out = np.zeros((h, w, c))
        self.lib.ref_stage_bicubic(_p(out), _p(ref), _p(im2), _p(vx), _p(vy), h, w, c)
        return out

    def smoothflow_sor(self, f1, f2, warp, u, v, alpha, nOuter, nInner, nSOR, lap_init_channels):
        f1, f2 = _c(f1), _c(f2)
        warp, u, v = _c(warp).copy(), _c(u).copy(), _c(v).copy()
        h, w, c = f1.shape
        with quiet_stdout():
            self.lib.ref_stage_smoothflow_sor(_p(f1), _p(f2), _p(warp), _p(u), _p(v), alpha,
                                              nOuter, nInner, nSOR, h, w, c, lap_init_channels)
        return warp, u, v


_cache = {}


def available():
    return os.path.exists(os.path.join(_REF_DIR, "libpyflow_ref_serial.so"))


def serial():
    if "s" not in _cache:
        _cache["s"] = _Ref(os.path.join(_REF_DIR, "libpyflow_ref_serial.so"), False)
    return _cache["s"]


def parallel():
    if "p" not in _cache:
        _cache["p"] = _Ref(os.path.join(_REF_DIR, "libpyflow_ref_parallel.so"), True)
    return _cache["p"]
