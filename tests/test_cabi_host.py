"""CPU-only checks of the product's host side: the C-ABI library loads, exports every symbol that
include/pyflow_b200.h declares, evaluates the pyramid geometry exactly like the reference, validates
arguments like the reference's Cython signature, and FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from papteam_opticalflow_b200 import _lib
import pyflow


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pyflow_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    L = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert _lib.lib().pf_version().startswith(b"pyflow_b200")


def test_geometry_matches_oracle(oracle_mod):
    L = _lib.lib()
    for w, h, ratio, mw in [(240, 135, 0.75, 20), (480, 270, 0.75, 20), (960, 540, 0.75, 20), (1920, 1080, 0.75, 20),
                            (3840, 2160, 0.75, 20), (640, 480, 0.5, 30), (640, 480, 0.6, 16), (333, 77, 0.9, 25),
                            (640, 480, 0.3, 20), (100, 100, 0.99, 40)]:
        n = L.pf_pyramid_levels(w, ratio, mw)
        assert n == oracle_mod.levels_from_min_width(w, ratio, mw)
        if n < 1:
            continue
        ws = (C.c_int * 64)(); hs = (C.c_int * 64)()
        assert L.pf_level_geometry(w, h, ratio, n, ws, hs) == 0
        geo = oracle_mod.level_geometry(w, h, ratio, n)
        assert [(ws[k], hs[k]) for k in range(n)] == [(g["w"], g["h"]) for g in geo]
    assert L.pf_pyramid_levels(1920, 0.75, 20) == 15 and L.pf_pyramid_levels(3840, 0.75, 20) == 18


def test_argument_validation_mirrors_cython_signature():
    a = np.zeros((8, 9, 3))
    with pytest.raises(TypeError):
        pyflow.coarse2fine_flow(None, a, 3)
    with pytest.raises(TypeError):
        pyflow.coarse2fine_flow([[1.0]], a, 3)
    with pytest.raises(ValueError):
        pyflow.coarse2fine_flow(a.astype(np.float32), a, 3)
    with pytest.raises(ValueError):
        pyflow.coarse2fine_flow(a[:, :, 0], a[:, :, 0], 3)
    with pytest.raises(ValueError):
        pyflow.coarse2fine_flow(np.asfortranarray(a), a, 3)
    with pytest.raises(ValueError):
        pyflow.coarse2fine_flow(a, np.zeros((8, 10, 3)), 3)
    with pytest.raises(TypeError):
        pyflow.coarse2fine_flow(a, a)                      # neither call shape
    with pytest.raises(TypeError):
        pyflow.coarse2fine_flow(a, a, 1, 2, 3, 4)
    with pytest.raises(ValueError):
        pyflow.coarse2fine_flow(a, a, 3, mode="fp16")
    with pytest.raises(TypeError):
        pyflow.coarse2fine_flow(a, a, 3, bogus=1)


def test_no_cpu_fallback_without_gpu():
    L = _lib.lib()
    if L.pf_device_count() > 0:
        pytest.skip("a GPU is present")
    a = np.zeros((16, 24, 3))
    with pytest.raises(_lib.PyflowB200Error) as e:
        pyflow.coarse2fine_flow(a, a, 2)
    assert e.value.code == _lib.PF_ENODEVICE and "no CPU fallback" in str(e.value)
    out = np.zeros((16, 24, 5))
    rc = L.pf_stage_im2feature(out.ctypes.data_as(_lib.dp), a.ctypes.data_as(_lib.dp), 16, 24, 3, 0, 0, 0)
    assert rc == _lib.PF_ENODEVICE
    # the sequence / visualisation entry points refuse just as loudly
    f = np.zeros((16, 24, 3), np.uint8)
    for output in ("float32", "u16", "bgr8"):
        with pytest.raises(_lib.PyflowB200Error) as e:
            pyflow.sequence_flow([f, f, f], devices=[0], output=output)
        assert e.value.code == _lib.PF_ENODEVICE
    with pytest.raises(_lib.PyflowB200Error) as e:
        pyflow.flow_to_bgr(np.zeros((16, 24, 2), np.float32))
    assert e.value.code == _lib.PF_ENODEVICE


def test_sequence_argument_validation_on_the_host():
    f = np.zeros((16, 24, 3), np.uint8)
    assert pyflow.sequence_flow([]) == ([], 0.0) and pyflow.sequence_flow([f]) == ([], 0.0)
    with pytest.raises(ValueError):
        pyflow.sequence_flow([f, f.astype(np.float64)])
    with pytest.raises(ValueError):
        pyflow.sequence_flow([f, f[:-1]])
    with pytest.raises(ValueError):
        pyflow.sequence_flow([f, f], output="png")
    with pytest.raises(ValueError):
        pyflow.sequence_flow([f, f], outs=[np.zeros((16, 24, 2), np.float64)])
    with pytest.raises(ValueError):
        pyflow.flow_to_bgr(np.zeros((16, 24, 3), np.float32))
    with pytest.raises(ValueError):
        pyflow.save_flow_u16("/dev/null", np.zeros((4, 4, 3), np.uint16))


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "papteam_opticalflow_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "oracle/", "oracle.", "liboracle", "_ref/", "pyflow_ref"):
                    assert bad not in txt, (os.path.join(dirpath, f), bad)


def test_solver_variant_switch_is_host_only_and_validated():
    """pf_set_solver_variant mirrors the reference's process-global statics (S/OpticalFlow.h:19-27); no GPU needed."""
    import pyflow
    assert pyflow.get_solver_variant() == ("bilinear", "lap")   # the reference's defaults, S/OpticalFlow.cpp:33-34
    try:
        pyflow.set_solver_variant("bicubic", "gmixture")
        assert pyflow.get_solver_variant() == ("bicubic", "gmixture")
        with pytest.raises(ValueError):
            pyflow.set_solver_variant("nearest", "lap")
        with pytest.raises(ValueError):
            pyflow.set_solver_variant("bilinear", "gaussian")
        assert pyflow.get_solver_variant() == ("bicubic", "gmixture")
        from papteam_opticalflow_b200 import _lib
        assert _lib.lib().pf_set_solver_variant(7, 1) == _lib.PF_EINVAL
    finally:
        pyflow.set_solver_variant("bilinear", "lap")
