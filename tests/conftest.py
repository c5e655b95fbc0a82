import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def load_frame(width, idx):
    """Decode a fixture frame exactly as the reference driver does
    (Par/OpticalFlowCalculation.py:66-71): PIL -> uint8 RGB -> float64 / 255."""
    from PIL import Image
    path = os.path.join(GOLDEN, "frames", "hcm%d_%05d.jpg" % (width, idx))
    return np.array(Image.open(path)).astype(float) / 255.


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def synthetic_pair(h, w, c=3, seed=0, shift=(1.5, -0.75)):
    """Smooth random texture and a translated copy (known motion), values in [0,1]."""
    rng = np.random.default_rng(seed)
    H, W = h + 16, w + 16
    base = rng.random((H, W, c))
    k = np.array([1, 4, 6, 4, 1], float); k /= k.sum()
    for _ in range(3):
        base = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 0, base)
        base = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, base)
    base = (base - base.min()) / (base.max() - base.min())
    yy, xx = np.mgrid[0:h, 0:w].astype(float)
    def sample(dx, dy):
        x = xx + 8 + dx; y = yy + 8 + dy
        x0 = np.floor(x).astype(int); y0 = np.floor(y).astype(int)
        fx = (x - x0)[..., None]; fy = (y - y0)[..., None]
        return ((1 - fx) * (1 - fy) * base[y0, x0] + fx * (1 - fy) * base[y0, x0 + 1]
                + (1 - fx) * fy * base[y0 + 1, x0] + fx * fy * base[y0 + 1, x0 + 1])
    im1 = np.ascontiguousarray(sample(0, 0))
    im2 = np.ascontiguousarray(sample(-shift[0], -shift[1]))
    return im1, im2


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def ref_serial():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return ref.serial()
