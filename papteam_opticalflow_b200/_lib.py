"""ctypes binding of include/pyflow_b200.h.  Loading fails loudly when the shared library has not
been built (python __graft_entry__.py build, or make -C papteam_opticalflow_b200/csrc)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpyflow_b200.so")

PF_NUM_TIMINGS = 16
(T_TOTAL, T_CONSTRUCTION, T_ALLOCATION, T_PHASE1, T_PHASE2, T_PHASE3, T_PHASE4, T_PHASE5, T_PHASE6,
 T_POST, T_H2D, T_D2H, T_SOLVE) = range(13)

PF_OK, PF_EINVAL, PF_ENODEVICE, PF_ECUDA, PF_ENOMEM, PF_EUNSUPPORTED = 0, -1, -2, -3, -4, -5

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
dpp = C.POINTER(dp)
_lib = None


class PyflowB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("pyflow_b200 error %d: %s" % (code, message))
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python __graft_entry__.py` or "
            "`make -C papteam_opticalflow_b200/csrc` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    # hardware work queues for the one-stream-per-pair concurrency; read by the driver at context creation
    # (the library's load-time constructor does the same -- this covers interpreters that fork workers)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    L = C.CDLL(LIB_PATH)
    i, d, v = C.c_int, C.c_double, C.c_void_p
    sig = {
        "pf_last_error": (C.c_char_p, []),
        "pf_version": (C.c_char_p, []),
        "pf_device_count": (i, []),
        "pf_host_alloc": (v, [C.c_size_t]),
        "pf_host_free": (None, [v]),
        "pf_pyramid_levels": (i, [i, d, i]),
        "pf_level_geometry": (i, [i, i, d, i, ip, ip]),
        "pf_coarse2fine_flow": (i, [dp, dp, dp, dp, dp, d, d, i, i, i, i, i, i, i, i, i, i, dp]),
        "pf_coarse2fine_flow_levels": (i, [dp, dp, dp, dp, dp, i, i, i, i, i, i, i, dp]),
        "pf_pool_clear": (i, []),
        "pf_set_solver_variant": (i, [i, i]),
        "pf_get_solver_variant": (i, [ip, ip]),
        "pf_plan_create": (i, [C.POINTER(v), i, i, i, d, d, i, i, i, i, i, i, i, i]),
        "pf_plan_create_tuned": (i, [C.POINTER(v), i, i, i, d, d, i, i, i, i, i, i, i, i, i]),
        "pf_plan_destroy": (i, [v]),
        "pf_plan_levels": (i, [v]),
        "pf_plan_execute": (i, [v, dp, dp, dp, dp, dp, dp]),
        "pf_plan_upload": (i, [v, dp, dp]),
        "pf_plan_solve": (i, [v, i, dp]),
        "pf_plan_download": (i, [v, dp, dp, dp]),
        "pf_plan_profile": (i, [v, dp, dp]),
        "pf_plan_level_timings": (i, [v, dp, i]),
        "pf_plan_mixture_params": (i, [v, dp, dp, dp, i]),
        "pf_multi_solve": (i, [C.POINTER(v), i, i, dp]),
        "pf_batch_flow": (i, [i, dpp, dpp, dpp, dpp, dpp, d, d, i, i, i, i, i, i, i, i, i, i, ip, i, dp]),
        "pf_batch_last_stats": (i, [dp]),
        "pf_sequence_flow_u8": (i, [i, C.POINTER(C.POINTER(C.c_ubyte)), C.POINTER(C.POINTER(C.c_float)), d, d, i, i, i, i, i, i, i, i, i, i,
                                    ip, i, dp]),
        "pf_sequence_flow_u8_u16": (i, [i, C.POINTER(C.POINTER(C.c_ubyte)), C.POINTER(C.POINTER(C.c_ushort)), d, d, i, i, i, i, i, i, i, i, i, i,
                                        ip, i, dp]),
        "pf_sequence_flow_u8_bgr": (i, [i, C.POINTER(C.POINTER(C.c_ubyte)), C.POINTER(C.POINTER(C.c_ubyte)), d, d, i, i, i, i, i, i, i, i, i, i,
                                        ip, i, dp]),
        "pf_flow_to_bgr": (i, [C.POINTER(C.c_float), C.POINTER(C.c_ubyte), i, i, i]),
        "pf_multigpu_flow": (i, [dp, dp, dp, dp, dp, d, d, i, i, i, i, i, i, i, i, i, ip, i, C.c_longlong, dp]),
        "pf_stage_pyramid": (i, [dp, dp, i, i, i, d, i, i, i]),
        "pf_stage_im2feature": (i, [dp, dp, i, i, i, i, i, i]),
        "pf_stage_getdxs": (i, [dp, dp, dp, dp, dp, i, i, i, i, i]),
        "pf_stage_warpfl": (i, [dp, dp, dp, dp, dp, i, i, i, i, i]),
        "pf_stage_resize_to": (i, [dp, dp, i, i, i, i, i, d, i, i]),
        "pf_stage_bicubic": (i, [dp, dp, dp, dp, dp, i, i, i, i, i]),
        "pf_stage_assemble": (i, [dp] * 14 + [d, i, i, i, i, i]),
        "pf_stage_sor": (i, [dp] * 8 + [d, i, i, i, i, i]),
        "pf_bench_sor": (i, [i, i, i, i, i, i, dp, dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)      # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(code):
    if code < 0:
        raise PyflowB200Error(code, lib().pf_last_error().decode("utf-8", "replace"))
    return code
