"""Host-side mirror of the reference's Cython module `pyflow` (Par/pyflow.pyx:31-70).

    coarse2fine_flow(Im1, Im2, pyramidLevels[, nCores])                      -> (timing, vx, vy, warpI2)
        the fork's call shape (Par/pyflow.pyx:31-33; serial flavour Ser/pyflow.pyx:31-33 has no nCores);
        alpha=0.012 ratio=0.75 7/1/30 are the fork's hard-coded constants (S/OpticalFlow.cpp:747-751).
    coarse2fine_flow(im1, im2, alpha, ratio, minWidth, nOuterFPIterations,
                     nInnerFPIterations, nSORIterations, colType)               -> (u, v, im2W)
        upstream pyflow's call shape, the one BASELINE.json's north_star names (parameter list at
        Par/pyflow.pyx:36-41).

Inputs are float64, C-contiguous, (h, w, c) arrays in [0,1], gray passed as (h, w, 1); outputs are
freshly allocated float64 arrays -- exactly the reference's contract.  Keyword-only extras:
    mode    'fp32_redblack' (fast, default) | 'fp64_wavefront' (matches the reference to <=1e-6)
            | 'fp64_redblack' | 'fp32_wavefront' | 'fp32_hybrid' (fast mode with the reference's sweep order on the
            pyramid levels <= 400 px wide)            (env PYFLOW_B200_MODE overrides the default)
    device  CUDA device index (default 0)
    profile True: run eagerly with CUDA events per phase and fill every timing key
"""
import ctypes as C
import os
import threading

import numpy as np

from . import _lib
from ._lib import PyflowB200Error, check, dp  # noqa: F401

MODES = {"fp64_wavefront": 0, "fp32_redblack": 1, "fp64_redblack": 2, "fp32_wavefront": 3, "fp32_hybrid": 4}
_FORK_DEFAULTS = dict(alpha=0.012, ratio=0.75, minWidth=20, nOuter=7, nInner=1, nSOR=30, colType=0)
# reference timing-map keys (S/OpticalFlow.cpp:850-860) in PF_T_* order, then the GPU-only legs
_TIMING_KEYS = ["Total C++ Execution", "Construction", "Allocation", "Phase1_Generate",
                "Phase2_Derivatives", "Phase3_PsiData", "Phase4_LinearSystem", "Phase5_SOR",
                "Phase6_Update", "PostProcessing", "H2D", "D2H", "Solve"]


TUNINGS = {"throughput": 0, "latency": 1}
INTERPOLATIONS = {"bilinear": 0, "bicubic": 1}
NOISE_MODELS = {"gmixture": 0, "lap": 1}


def set_solver_variant(interpolation="bilinear", noise_model="lap"):
    """Selects the alternative solver branches the reference keeps behind two process-global public statics,
    `OpticalFlow::interpolation` and `OpticalFlow::noiseModel` (S/OpticalFlow.h:19-27; defaults Bilinear / Lap at
    S/OpticalFlow.cpp:33-34).  Like there, the switch is process-global; it is sampled when a plan is created or a
    one-shot / batch / sequence call starts."""
    if interpolation not in INTERPOLATIONS:
        raise ValueError("unknown interpolation %r (expected one of %s)" % (interpolation, sorted(INTERPOLATIONS)))
    if noise_model not in NOISE_MODELS:
        raise ValueError("unknown noise_model %r (expected one of %s)" % (noise_model, sorted(NOISE_MODELS)))
    check(_lib.lib().pf_set_solver_variant(INTERPOLATIONS[interpolation], NOISE_MODELS[noise_model]))


def get_solver_variant():
    a, b = C.c_int(), C.c_int()
    check(_lib.lib().pf_get_solver_variant(C.byref(a), C.byref(b)))
    inv = lambda d, v: [k for k, x in d.items() if x == v][0]
    return inv(INTERPOLATIONS, a.value), inv(NOISE_MODELS, b.value)


def _mode_id(mode):
    if mode is None:
        mode = os.environ.get("PYFLOW_B200_MODE", "fp32_redblack")
    if isinstance(mode, str):
        if mode not in MODES:
            raise ValueError("unknown mode %r (expected one of %s)" % (mode, sorted(MODES)))
        return MODES[mode]
    if mode in MODES.values():
        return int(mode)
    raise ValueError("unknown mode %r" % (mode,))


def _check_image(name, a):
    """Same acceptance rules as Cython's `np.ndarray[double, ndim=3, mode="c"] X not None`."""
    if a is None:
        raise TypeError("Argument '%s' must not be None" % name)
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
    if a.ndim != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % a.ndim)
    if a.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double' but got '%s'" % a.dtype.name)
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("ndarray is not C-contiguous")
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(dp)


def _check_outputs(out, h, w, c):
    """Caller-supplied outputs are written by native code as raw double*: require exactly what the module itself would
    allocate (Par/pyflow.pyx:47-52: float64, C-contiguous, (h, w), (h, w), (h, w, c)).  Any of the three may be None:
    that output is then not copied back from the device."""
    try:
        vx, vy, wi = out
    except (TypeError, ValueError):
        raise ValueError("outputs must be a (vx, vy, warpI2) triple")
    for name, a, shape in (("vx", vx, (h, w)), ("vy", vy, (h, w)), ("warpI2", wi, (h, w, c))):
        if a is None:
            continue
        if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.shape != shape or not a.flags["C_CONTIGUOUS"] \
                or not a.flags["WRITEABLE"]:
            raise ValueError("output %s must be a writeable C-contiguous float64 array of shape %s" % (name, shape))
    return vx, vy, wi


class FlowPlan(object):
    """Device arena + captured CUDA graph for one image shape / parameter set (pf_plan_*)."""

    def __init__(self, h, w, c, alpha=0.012, ratio=0.75, minWidth=20, nOuter=7, nInner=1, nSOR=30,
                 colType=0, levels=0, mode=None, device=0, tuning="throughput"):
        """tuning: 'throughput' (plans that run concurrently with others) or 'latency' (one pair at a time); it only
        selects how many SOR sweeps a launch fuses per level, the results are bit-identical."""
        if tuning not in TUNINGS:
            raise ValueError("unknown tuning %r (expected one of %s)" % (tuning, sorted(TUNINGS)))
        self.shape = (int(h), int(w), int(c))
        self.mode = _mode_id(mode)
        self.device = int(device)
        self._h = C.c_void_p()
        self._lock = threading.Lock()
        check(_lib.lib().pf_plan_create_tuned(C.byref(self._h), h, w, c, float(alpha), float(ratio), int(minWidth),
                                              int(levels), int(nOuter), int(nInner), int(nSOR), int(colType),
                                              self.mode, self.device, TUNINGS[tuning]))
        self.levels = _lib.lib().pf_plan_levels(self._h)

    def close(self):
        lock = getattr(self, "_lock", None)
        if lock is None:
            return
        with lock:                           # never under a native call of another thread
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().pf_plan_destroy(self._h)
                self._h = C.c_void_p()

    def _handle(self):
        if not self._h.value:
            raise ValueError("plan is closed")
        return self._h

    __del__ = close

    def _outputs(self):
        h, w, c = self.shape
        return np.zeros((h, w)), np.zeros((h, w)), np.zeros((h, w, c))

    def _check_pair(self, im1, im2):
        _check_image("Im1", im1)
        _check_image("Im2", im2)
        if im1.shape != self.shape or im2.shape != self.shape:
            raise ValueError("image shapes %s / %s do not match the plan's %s" % (im1.shape, im2.shape, self.shape))

    def execute(self, im1, im2, out=None):
        """H2D + solve + D2H.  Returns (timings_ms ndarray, vx, vy, warpI2)."""
        self._check_pair(im1, im2)
        vx, vy, wi = _check_outputs(out, *self.shape) if out is not None else self._outputs()
        t = np.zeros(_lib.PF_NUM_TIMINGS)
        with self._lock:
            check(_lib.lib().pf_plan_execute(self._handle(), _ptr(vx), _ptr(vy), _ptr(wi), _ptr(im1), _ptr(im2), _ptr(t)))
        return t, vx, vy, wi

    def upload(self, im1, im2):
        self._check_pair(im1, im2)
        with self._lock:
            check(_lib.lib().pf_plan_upload(self._handle(), _ptr(im1), _ptr(im2)))

    def solve(self, repeats=1):
        """Device-only solve on the resident inputs; returns total milliseconds (CUDA events)."""
        ms = C.c_double()
        with self._lock:
            check(_lib.lib().pf_plan_solve(self._handle(), int(repeats), C.byref(ms)))
        return ms.value

    def download(self, out=None):
        vx, vy, wi = _check_outputs(out, *self.shape) if out is not None else self._outputs()
        with self._lock:
            check(_lib.lib().pf_plan_download(self._handle(), _ptr(vx), _ptr(vy), _ptr(wi)))
        return vx, vy, wi

    def profile(self):
        """One eager solve with events around every phase: (timings_ms, counters)."""
        t = np.zeros(_lib.PF_NUM_TIMINGS)
        cnt = np.zeros(8)
        with self._lock:
            check(_lib.lib().pf_plan_profile(self._handle(), _ptr(t), _ptr(cnt)))
        return t, cnt

    def mixture_params(self):
        """(alpha, sigma, beta) per feature channel after the last solve (Gaussian-mixture noise model only)."""
        a, s, b = np.zeros(16), np.zeros(16), np.zeros(16)
        n = _lib.lib().pf_plan_mixture_params(self._handle(), _ptr(a), _ptr(s), _ptr(b), 16)
        if n < 0:
            check(n)
        return a[:n].copy(), s[:n].copy(), b[:n].copy()

    def level_timings(self):
        """(levels, PF_NUM_TIMINGS) array of per-level phase milliseconds from the last profile()."""
        out = np.zeros((self.levels, _lib.PF_NUM_TIMINGS))
        with self._lock:
            check(_lib.lib().pf_plan_level_timings(self._handle(), _ptr(out), self.levels))
        return out


def multi_solve(plans, repeats=1):
    """Concurrent device-only solves of several resident plans (one stream each) on one device;
    returns total milliseconds from CUDA events (pf_multi_solve)."""
    arr = (C.c_void_p * len(plans))(*[p._h for p in plans])
    ms = C.c_double()
    check(_lib.lib().pf_multi_solve(arr, len(plans), int(repeats), C.byref(ms)))
    return ms.value


def _timing_dict(t_ms):
    # the reference stringifies seconds with std::to_string (6 decimals)
    return {k: "%.6f" % (t_ms[i] / 1000.0) for i, k in enumerate(_TIMING_KEYS)}


def coarse2fine_flow(Im1, Im2, *args, **kwargs):
    mode = kwargs.pop("mode", None)
    device = kwargs.pop("device", 0)
    profile = kwargs.pop("profile", False)
    if kwargs:
        raise TypeError("coarse2fine_flow() got unexpected keyword arguments %s" % sorted(kwargs))
    n = len(args)
    if n in (1, 2):
        fork = True
        levels = int(args[0])
        if n == 2:
            int(args[1])                      # nCores: accepted, type-checked, ignored
        if levels < 1:
            raise ValueError("pyramidLevels must be >= 1")
        p = dict(_FORK_DEFAULTS)
    elif n == 7:
        fork = False
        levels = 0
        p = dict(alpha=float(args[0]), ratio=float(args[1]), minWidth=int(args[2]), nOuter=int(args[3]),
                 nInner=int(args[4]), nSOR=int(args[5]), colType=int(args[6]))
    else:
        raise TypeError("coarse2fine_flow() takes (Im1, Im2, pyramidLevels[, nCores]) or (im1, im2, alpha, ratio, "
                        "minWidth, nOuterFPIterations, nInnerFPIterations, nSORIterations, colType); got %d positional "
                        "arguments" % (n + 2))
    _check_image("Im1", Im1)
    _check_image("Im2", Im2)
    if Im1.shape != Im2.shape:
        raise ValueError("Im1 and Im2 must have the same shape, got %s and %s" % (Im1.shape, Im2.shape))
    h, w, c = Im1.shape
    mid = _mode_id(mode)
    # One pair at a time goes through the library's own one-shot entry points: they draw an idle plan (arena + captured
    # graph, latency-tuned) from the native pool, which never hands a plan to two callers and never destroys one in use.
    L = _lib.lib()
    vx, vy, wi = np.zeros((h, w)), np.zeros((h, w)), np.zeros((h, w, c))
    t = np.zeros(_lib.PF_NUM_TIMINGS)
    if fork:
        check(L.pf_coarse2fine_flow_levels(_ptr(vx), _ptr(vy), _ptr(wi), _ptr(Im1), _ptr(Im2), levels, 1, h, w, c, mid,
                                           int(device), _ptr(t)))
        if profile:
            plan = FlowPlan(h, w, c, levels=levels, mode=mid, device=device, tuning="latency", **p)   # private to this call
            try:
                plan.upload(Im1, Im2)
                tp, _ = plan.profile()
            finally:
                plan.close()
            for i in range(1, 10):
                t[i] = tp[i]
        return _timing_dict(t), vx, vy, wi
    check(L.pf_coarse2fine_flow(_ptr(vx), _ptr(vy), _ptr(wi), _ptr(Im1), _ptr(Im2), p["alpha"], p["ratio"], p["minWidth"],
                                p["nOuter"], p["nInner"], p["nSOR"], p["colType"], h, w, c, mid, int(device), _ptr(t)))
    return vx, vy, wi


_SEQ_OUTPUTS = {"float32": (np.float32, C.c_float, 2, "pf_sequence_flow_u8"),
                "u16": (np.uint16, C.c_ushort, 2, "pf_sequence_flow_u8_u16"),
                "bgr8": (np.uint8, C.c_ubyte, 3, "pf_sequence_flow_u8_bgr")}


def sequence_flow(frames, alpha=0.012, ratio=0.75, minWidth=20, nOuterFPIterations=7, nInnerFPIterations=1,
                  nSORIterations=30, colType=0, levels=0, mode=None, devices=None, output="float32", outs=None):
    """Flows of the consecutive pairs of a frame sequence.

    frames: list of (h, w, c) uint8 C-contiguous arrays (what PIL decodes; the reference driver converts
    them with astype(float)/255. before calling pyflow -- here that happens on the device).
    Returns ([out_0, ..., out_{n-2}], seconds), out_t belonging to pair (t, t+1):
      output="float32": (h, w, 2) float32 (u, v), equal to coarse2fine_flow(frames[t]/255., frames[t+1]/255., ...)
                        cast to float32;
      output="u16":     (h, w, 2) uint16 in the reference's flow-file encoding (S/OpticalFlow.cpp:993-1003; see
                        decode_flow_u16 / save_flow_u16);
      output="bgr8":    (h, w, 3) uint8, the HSV flow image the reference driver writes with cv2.imwrite
                        (Par/OpticalFlowCalculation.py:143-162), see flow_to_bgr.
    outs: optional preallocated (e.g. pinned) output arrays."""
    L = _lib.lib()
    if output not in _SEQ_OUTPUTS:
        raise ValueError("output must be one of %s" % sorted(_SEQ_OUTPUTS))
    dt, ct, nch, fname = _SEQ_OUTPUTS[output]
    n = len(frames)
    if n < 2:
        return [], 0.0
    for f in frames:
        if not isinstance(f, np.ndarray) or f.dtype != np.uint8 or f.ndim != 3 or not f.flags["C_CONTIGUOUS"]:
            raise ValueError("frames must be C-contiguous (h, w, c) uint8 arrays")
        if f.shape != frames[0].shape:
            raise ValueError("all frames of a sequence must share one shape")
    if devices is None:
        devices = list(range(max(1, L.pf_device_count())))
    h, w, c = frames[0].shape
    if outs is None:
        res = [np.zeros((h, w, nch), dtype=dt) for _ in range(n - 1)]
    else:
        res = list(outs)
        if len(res) != n - 1 or any(o.dtype != dt or o.shape != (h, w, nch) or not o.flags["C_CONTIGUOUS"] for o in res):
            raise ValueError("outs must be %d C-contiguous (h, w, %d) %s arrays" % (n - 1, nch, np.dtype(dt).name))
    fp = (C.POINTER(C.c_ubyte) * n)(*[f.ctypes.data_as(C.POINTER(C.c_ubyte)) for f in frames])
    op = (C.POINTER(ct) * (n - 1))(*[f.ctypes.data_as(C.POINTER(ct)) for f in res])
    dev = (C.c_int * len(devices))(*devices)
    secs = C.c_double()
    check(getattr(L, fname)(n, fp, op, float(alpha), float(ratio), int(minWidth), int(levels), int(nOuterFPIterations),
                            int(nInnerFPIterations), int(nSORIterations), int(colType), h, w, c, _mode_id(mode), dev,
                            len(devices), C.byref(secs)))
    return res, secs.value


def flow_to_bgr(flow, v=None, device=0):
    """The reference driver's flow visualisation on the device (Par/OpticalFlowCalculation.py:143-162 up to, not
    including, cv2.imwrite): flow_to_bgr(flow(h, w, 2)) or flow_to_bgr(u, v) -> (h, w, 3) uint8 BGR.
    The flow is taken in float32 (cv2.cartToPolar itself computes in float32)."""
    if v is not None:
        flow = np.stack([np.asarray(flow), np.asarray(v)], axis=-1)
    flow = np.ascontiguousarray(flow, dtype=np.float32)
    if flow.ndim != 3 or flow.shape[2] != 2:
        raise ValueError("flow must be (h, w, 2)")
    h, w, _ = flow.shape
    out = np.zeros((h, w, 3), dtype=np.uint8)
    check(_lib.lib().pf_flow_to_bgr(flow.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_ubyte)), h, w,
                                    int(device)))
    return out


# ---- the reference's flow file format -- host-side container I/O only; the encoding
#      itself is produced on the device by sequence_flow(encoded=True) ------------------------------------
def decode_flow_u16(q):
    """OpticalFlow::LoadOpticalFlow (S/OpticalFlow.cpp:962-975): f = (double)q / 160 - 200."""
    return np.asarray(q, dtype=np.float64) / 160 - 200


def save_flow_u16(path, q):
    """Write an encoded (h, w, 2) uint16 flow as the reference's Image<unsigned short>::saveImage container
    (S/Image.h:825-836): 16 bytes of type tag (g++'s typeid(unsigned short).name() = "t", zero padded here --
    the reference leaves the 14 trailing bytes uninitialised), int32 width, height, channels, one bool
    IsDerivativeImage (0), then the samples, all little endian."""
    q = np.ascontiguousarray(q, dtype=np.uint16)
    if q.ndim != 3 or q.shape[2] != 2:
        raise ValueError("encoded flow must be (h, w, 2) uint16")
    with open(path, "wb") as f:
        f.write(b"t".ljust(16, b"\0"))
        f.write(np.array([q.shape[1], q.shape[0], 2], dtype="<i4").tobytes())
        f.write(b"\0")
        f.write(q.astype("<u2").tobytes())


def load_flow_u16(path):
    """Read a file written by save_flow_u16 or by the reference's SaveOpticalFlow; returns (h, w, 2) uint16."""
    with open(path, "rb") as f:
        f.read(16)                                  # type tag: only its first byte(s) are defined
        w, h, c = np.frombuffer(f.read(12), dtype="<i4")
        f.read(1)
        q = np.frombuffer(f.read(int(w) * int(h) * int(c) * 2), dtype="<u2")
    if c != 2 or q.size != w * h * c:
        raise ValueError("not an encoded two-channel flow file")
    return q.reshape(int(h), int(w), 2).copy()


def coarse2fine_flow_multigpu(im1, im2, alpha=0.012, ratio=0.75, minWidth=20, nOuterFPIterations=7,
                              nInnerFPIterations=1, nSORIterations=30, colType=0, levels=0, devices=None,
                              split_min_pixels=-1):
    """ONE pair solved by several GPUs: the SOR solve of the fine levels is split into row bands with
    NVLink peer-to-peer halo exchange (pf_multigpu_flow; FP32 red-black mode, bit-identical to the
    single-GPU fast mode).  Returns (u, v, im2W, stats) with stats = dict(ms, halo_bytes, gather_bytes,
    split_solves, graph, flag_solves)."""
    L = _lib.lib()
    _check_image("Im1", im1)
    _check_image("Im2", im2)
    if im1.shape != im2.shape:
        raise ValueError("Im1 and Im2 must have the same shape")
    if devices is None:
        devices = list(range(max(1, L.pf_device_count())))
    h, w, c = im1.shape
    vx, vy, wi = np.zeros((h, w)), np.zeros((h, w)), np.zeros((h, w, c))
    dev = (C.c_int * len(devices))(*devices)
    st = np.zeros(8)
    check(L.pf_multigpu_flow(_ptr(vx), _ptr(vy), _ptr(wi), _ptr(im1), _ptr(im2), float(alpha), float(ratio), int(minWidth),
                             int(levels), int(nOuterFPIterations), int(nInnerFPIterations), int(nSORIterations),
                             int(colType), h, w, c, dev, len(devices), int(split_min_pixels), _ptr(st)))
    return vx, vy, wi, dict(ms=st[0], halo_bytes=int(st[1]), gather_bytes=int(st[2]), split_solves=int(st[3]),
                            graph=bool(st[4]), flag_solves=int(st[5]))


def coarse2fine_flow_batch(pairs, alpha=0.012, ratio=0.75, minWidth=20, nOuterFPIterations=7,
                           nInnerFPIterations=1, nSORIterations=30, colType=0, levels=0, mode=None,
                           devices=None, outs=None):
    """Solve a list of independent (im1, im2) pairs, pair p on devices[p % len(devices)], no
    collective.  Returns ([(u, v, im2W), ...], wall_seconds)."""
    L = _lib.lib()
    if devices is None:
        devices = list(range(max(1, L.pf_device_count())))
    n = len(pairs)
    if n == 0:
        return [], 0.0
    for a, b in pairs:
        _check_image("Im1", a)
        _check_image("Im2", b)
        if a.shape != pairs[0][0].shape or b.shape != a.shape:
            raise ValueError("all images of a batch must share one shape")
    h, w, c = pairs[0][0].shape
    if outs is None:
        outs = [(np.zeros((h, w)), np.zeros((h, w)), np.zeros((h, w, c))) for _ in range(n)]
    else:
        if len(outs) != n:
            raise ValueError("outs must hold one (vx, vy, warpI2) triple per pair")
        outs = [_check_outputs(o, h, w, c) for o in outs]     # entries may be None: that output is not copied back
    arr = lambda xs: (dp * n)(*[_ptr(x) for x in xs])  # noqa: E731
    dev = (C.c_int * len(devices))(*devices)
    secs = C.c_double()
    check(L.pf_batch_flow(n, arr([o[0] for o in outs]), arr([o[1] for o in outs]), arr([o[2] for o in outs]),
                          arr([p[0] for p in pairs]), arr([p[1] for p in pairs]), float(alpha), float(ratio),
                          int(minWidth), int(levels), int(nOuterFPIterations), int(nInnerFPIterations),
                          int(nSORIterations), int(colType), h, w, c, _mode_id(mode), dev, len(devices),
                          C.byref(secs)))
    return outs, secs.value


def batch_last_stats():
    """Event-timed legs of the last coarse2fine_flow_batch call (pf_batch_last_stats): dict(pairs, workers, seconds,
    h2d_ms, solve_ms, d2h_ms, call_ms) -- means per pair."""
    st = np.zeros(8)
    check(_lib.lib().pf_batch_last_stats(_ptr(st)))
    return dict(pairs=int(st[0]), workers=int(st[1]), seconds=st[2], h2d_ms=st[3], solve_ms=st[4], d2h_ms=st[5], call_ms=st[6])
