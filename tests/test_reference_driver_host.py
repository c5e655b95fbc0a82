"""SURVEY.md 8f row f1 (driver compatibility), host side: the reference's own driver module
(Par/OpticalFlowCalculation.py, imported UNMODIFIED from /root/reference when that tree is present) is run against this
repository's `pyflow` module.  Without a GPU the call must get through the driver's own preprocessing and our argument
validation and stop exactly at the device check (PF_ENODEVICE): the drop-in accepts what the driver passes
(`im.astype(float) / 255.` arrays, pyramidLevels, numCores -- Par/OpticalFlowCalculation.py:66-74).
On a machine with a GPU the same call runs the solve and the driver writes its outputs."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN

DRIVERS = {"parallel": "/root/reference/Code/Parallel/OpticalFlowCalculation.py",   # coarse2fine_flow(im1, im2, levels, nCores)
           "serial": "/root/reference/Code/Serial/OpticalFlowCalculation.py"}       # coarse2fine_flow(im1, im2, levels)


class _Img:
    def __init__(self, path, idx):
        self.IMAGE_PATH = path
        self.IMAGE_PARENT = "HoChiMinhTraffic_10FPS_240"
        self.IMAGE_INDEX_STRING = "%05d" % idx
        self.IMAGE_INDEX = idx


class _Pair:
    def __init__(self):
        self.BEFORE = _Img(os.path.join(GOLDEN, "frames", "hcm240_00001.jpg"), 1)
        self.AFTER = _Img(os.path.join(GOLDEN, "frames", "hcm240_00002.jpg"), 2)


@pytest.mark.parametrize("flavour", ["parallel", "serial"])
def test_reference_driver_calls_the_drop_in_module(flavour, tmp_path, monkeypatch):
    DRIVER = DRIVERS[flavour]
    if not os.path.exists(DRIVER):
        pytest.skip("reference tree not present")
    call = (lambda drv: drv.CalculateOpticalFlow(_Pair(), 8, 4)) if flavour == "parallel" else (lambda drv: drv.CalculateOpticalFlow(_Pair(), 8))
    import pyflow                     # this repository's module
    from papteam_opticalflow_b200 import _lib
    assert "papteam_opticalflow_b200" in (pyflow.coarse2fine_flow.__module__ or "")
    # the driver imports matplotlib (unused on this path, not installed here) and its InputCreation package (optional)
    for name in ("matplotlib", "matplotlib.pyplot"):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("ref_driver_" + flavour, DRIVER)
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    assert drv.pyflow is pyflow       # the driver bound OUR module under the reference's name
    seen = {}
    real = pyflow.coarse2fine_flow

    def spy(*a, **k):
        seen["args"] = a
        return real(*a, **k)

    monkeypatch.setattr(drv.pyflow, "coarse2fine_flow", spy)
    monkeypatch.chdir(tmp_path)       # the driver writes under ./output
    if _lib.lib().pf_device_count() > 0:
        call(drv)
        assert os.path.isdir(tmp_path / "output")
    else:
        with pytest.raises(pyflow.PyflowB200Error) as e:
            call(drv)
        assert e.value.code == _lib.PF_ENODEVICE
    im1, im2, levels = seen["args"][:3]
    assert im1.dtype == np.float64 and im1.shape == (135, 240, 3) and im1.flags["C_CONTIGUOUS"] and im2.shape == im1.shape
    assert levels == 8 and seen["args"][3:] == ((4,) if flavour == "parallel" else ()) and 0.0 <= im1.min() and im1.max() <= 1.0
