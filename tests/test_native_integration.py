"""Native-level integration: the reference's OWN Cython module --
Par/pyflow.pyx + Par/coarse2Fine.pxd + P/Coarse2FineFlowWrapper.h, unmodified, cythonized and compiled where they lie --
linked against integration/Coarse2FineFlowWrapper.cpp, the shim over libpyflow_b200.so (integration/Makefile; built by
__graft_entry__.build() when /root/reference is present, the .so travels to the GPU box).

CPU: the module imports, Cython's own buffer checks reject what the reference rejects, and a call stops at the
library's PF_ENODEVICE (no CPU fallback) without taking the interpreter down.
GPU: the call through the reference's binding equals pf_coarse2fine_flow_levels bit for bit, in both modes."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden, load_frame

BUILD = os.path.join(ROOT, "integration", "_build")


def _ext():
    so = glob.glob(os.path.join(BUILD, "pyflow.*.so"))
    if not so:
        pytest.skip("integration/_build not built (reference tree absent at build time)")
    return so[0]


def _run(code, env=None, timeout=600):
    # a fresh interpreter whose `import pyflow` finds the Cython extension, not this repository's pyflow.py
    full = ("import sys; sys.path.insert(0, %r); sys.path = [p for p in sys.path if p not in ('', %r)]\n"
            "import pyflow, numpy as np, json\nassert pyflow.__file__.endswith('.so'), pyflow.__file__\n" % (BUILD, ROOT)) + code
    e = dict(os.environ)
    e.pop("PYTHONPATH", None)
    e.update(env or {})
    return subprocess.run([sys.executable, "-c", full], capture_output=True, text=True, timeout=timeout, env=e, cwd="/tmp")


def test_reference_cython_module_builds_and_binds_the_shim():
    so = _ext()
    # the extension links the product library and nothing of the reference's solver
    out = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "libpyflow_b200.so" in out
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    assert "PyInit_pyflow" in syms and "Coarse2FineFlowWrapper" in syms
    assert "OpticalFlow" not in syms and "GaussianPyramid" not in syms
    r = _run("a = np.zeros((12, 16, 3))\n"
             "for bad in (a.astype(np.float32), a[..., 0], None):\n"
             "    try:\n"
             "        pyflow.coarse2fine_flow(bad, a, 3, 1); print('accepted')\n"
             "    except (ValueError, TypeError) as e:\n"
             "        print(type(e).__name__)\n")
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.split() == ["ValueError", "ValueError", "TypeError"]


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present: the call succeeds (see the gpu test)")
def test_call_without_a_device_reports_enodevice_the_reference_way():
    _ext()
    r = _run("a = np.random.default_rng(0).random((24, 32, 3))\n"
             "t, vx, vy, w = pyflow.coarse2fine_flow(a, a, 3, 1)\n"
             "print(json.dumps({'err': t.get('error', ''), 'zero': bool((vx == 0).all() and (w == 0).all()), 'keys': sorted(t)}))\n")
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert "no usable CUDA device" in res["err"] and res["zero"] and "Total C++ Execution" in res["keys"]
    assert "pyflow_b200:" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32_redblack", "fp64_wavefront"])
def test_reference_binding_on_gpu_equals_the_c_abi(mode, tmp_path):
    _ext()
    import pyflow as ours
    a, b = load_frame(240, 1), load_frame(240, 2)
    np.save(tmp_path / "a.npy", a); np.save(tmp_path / "b.npy", b)
    r = _run("a, b = np.load(%r), np.load(%r)\n"
             "t, vx, vy, w = pyflow.coarse2fine_flow(a, b, 8, 4)\n"
             "assert 'error' not in t and float(t['Total C++ Execution']) > 0, t\n"
             "np.savez(%r, vx=vx, vy=vy, w=w)\n" % (str(tmp_path / "a.npy"), str(tmp_path / "b.npy"), str(tmp_path / "out.npz")),
             env={"PYFLOW_B200_MODE": mode})
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.load(tmp_path / "out.npz")
    _, vx, vy, wi = ours.coarse2fine_flow(a, b, 8, 4, mode=mode)
    assert np.array_equal(got["vx"], vx) and np.array_equal(got["vy"], vy) and np.array_equal(got["w"], wi)
    if mode == "fp64_wavefront":                      # and therefore the reference itself, to 1e-6
        g = golden("hcm240_L8.npz")
        assert np.abs(got["vx"] - g["vx"]).max() <= 1e-6 and np.abs(got["w"] - g["warpI2"]).max() <= 1e-6
