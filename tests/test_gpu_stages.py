"""Per-stage parity of the CUDA kernels (through the C ABI) against the plain-C oracle, which is
itself pinned bit-for-bit to the reference (tests/test_oracle_vs_ref.py).

FP64 instantiation: compiled with -fmad=false and written in the reference's operation order, so
the bar is bit-exact (np.array_equal) except where noted.  FP32 instantiation: tolerance 2e-5
relative to the stage's dynamic range (single-precision rounding of O(10) operations)."""
import numpy as np
import pytest

import gpu_util as G
from conftest import golden, synthetic_pair

pytestmark = pytest.mark.gpu

SHAPES = [((40, 56), 3), ((33, 47), 1), ((21, 130), 3), ((70, 23), 1), ((5, 9), 3)]


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-30, np.abs(np.asarray(b)).max())


def fields(shape, seed=1):
    rng = np.random.default_rng(seed)
    u = rng.normal(size=shape) * 3
    v = rng.normal(size=shape) * 3
    u[0, :] = -5
    v[-1, :] = 7           # samples that leave the image
    wt = rng.random(shape) + 0.05
    return u, v, wt


@pytest.mark.parametrize("shape,c", SHAPES)
def test_stages_fp64_bit_exact(oracle_mod, shape, c):
    o = oracle_mod
    h, w = shape
    im1, im2 = synthetic_pair(h, w, c, seed=h * w)
    u, v, wt = fields(shape)
    nl = 4 if min(h, w) > 12 else 2
    for a, b in zip(G.pyramid(im1, 0.75, nl, G.F64), o.pyramid(im1, 0.75, nl)):
        assert np.array_equal(a, b)
    f1 = o.im2feature(im1); f2 = o.im2feature(im2)
    assert np.array_equal(G.im2feature(im1, G.F64), f1)
    assert np.array_equal(G.im2feature(im1, G.F64, swap=1), o.im2feature(im1, 1))
    for a, b in zip(G.getdxs(f1, f2, G.F64), o.getdxs(f1, f2)):
        assert np.array_equal(a, b)
    assert np.array_equal(G.warpfl(f1, f2, u, v, G.F64), o.warpfl(f1, f2, u, v))
    dh, dw = int(h / 0.75), int(w / 0.75)
    assert np.array_equal(G.resize_to(u, dh, dw, 1 / 0.75, G.F64), o.resize_to(u, dh, dw, 1 / 0.75))
    assert np.array_equal(G.bicubic(im1, im2, u, v, G.F64), o.bicubic_warp(im1, im2, u, v))


@pytest.mark.parametrize("shape,c", SHAPES)
def test_assemble_and_sor_fp64(oracle_mod, shape, c):
    o = oracle_mod
    h, w = shape
    im1, im2 = synthetic_pair(h, w, c, seed=7 + h)
    f1, f2 = o.im2feature(im1), o.im2feature(im2)
    dx, dy, dt = o.getdxs(f1, f2)
    u, v, _ = fields(shape, 3)
    u *= 0.2; v *= 0.2
    rng = np.random.default_rng(5)
    du = rng.normal(size=shape) * 0.1; dv = rng.normal(size=shape) * 0.1
    fc = f1.shape[2]
    lap = np.full(fc, 0.02)
    for ddu, ddv in ((None, None), (du, dv)):
        want = o.assemble(dx, dy, dt, u, v, np.zeros(shape) if ddu is None else ddu,
                          np.zeros(shape) if ddv is None else ddv, lap, 0.012)
        got = G.assemble(dx, dy, dt, u, v, ddu, ddv, lap, 0.012, G.F64)
        for k in want:
            assert np.array_equal(got[k], want[k]), k
    # psi guard: a channel whose noise scale is < 1e-20 contributes nothing (S/OpticalFlow.cpp:399-400)
    lap0 = lap.copy(); lap0[0] = 0.0
    want = o.assemble(dx, dy, dt, u, v, np.zeros(shape), np.zeros(shape), lap0, 0.012)
    got = G.assemble(dx, dy, dt, u, v, None, None, lap0, 0.012, G.F64)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    # SOR, lexicographic wavefront == the reference's loop order, bit for bit
    s = o.assemble(dx, dy, dt, u, v, np.zeros(shape), np.zeros(shape), lap, 0.012)
    for nsor in (1, 2, 9):
        wdu, wdv = o.sor_solve(s["phi"], s["dxy"], s["dx2"], s["dy2"], s["bu"], s["bv"], 0.012, nsor, o.LEX)
        gdu, gdv = G.sor(s["phi"], s["dxy"], s["dx2"], s["dy2"], s["bu"], s["bv"], 0.012, nsor, G.F64)
        assert np.array_equal(gdu, wdu) and np.array_equal(gdv, wdv), nsor


@pytest.mark.parametrize("shape", [(40, 56), (33, 47), (64, 64), (65, 66), (70, 200), (150, 130), (9, 300), (300, 9)])
@pytest.mark.parametrize("nsor", [1, 3, 7, 20])
def test_sor_redblack_tile_kernel_fp64(oracle_mod, shape, nsor, monkeypatch):
    """The temporally blocked register kernel against the oracle's red-black ordering.  Same
    update order, so differences come only from the association of the 4-neighbour sum:
    tolerance 1e-12 relative.  Shapes straddle the 64-wide region and the tile halos; several fused
    sweep counts force multi-tile overlap handling."""
    o = oracle_mod
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    phi = rng.random(shape) * 30 + 0.5
    dxy = (rng.random(shape) - 0.5) * 0.2
    dx2 = rng.random(shape) + 0.1
    dy2 = rng.random(shape) + 0.1
    bu = rng.random(shape) - 0.5
    bv = rng.random(shape) - 0.5
    wdu, wdv = o.sor_solve(phi, dxy, dx2, dy2, bu, bv, 0.012, nsor, o.REDBLACK)
    for fuse in ("0", "1", "2", "3", "5"):
        monkeypatch.setenv("PF_SOR_FUSE", fuse)
        gdu, gdv = G.sor(phi, dxy, dx2, dy2, bu, bv, 0.012, nsor, G.F64RB)
        assert rel(gdu, wdu) < 1e-12 and rel(gdv, wdv) < 1e-12, fuse
    monkeypatch.setenv("PF_SOR_FUSE", "0")
    monkeypatch.setenv("PF_SOR_SIMPLE", "1")     # cross-check kernel: one half-sweep per launch
    gdu, gdv = G.sor(phi, dxy, dx2, dy2, bu, bv, 0.012, nsor, G.F64RB)
    assert rel(gdu, wdu) < 1e-12 and rel(gdv, wdv) < 1e-12
    monkeypatch.delenv("PF_SOR_SIMPLE")
    # FP32 instantiation of the same kernel: single-precision tolerance
    gdu, gdv = G.sor(phi, dxy, dx2, dy2, bu, bv, 0.012, nsor, G.F32)
    assert rel(gdu, wdu) < 5e-4 and rel(gdv, wdv) < 5e-4


@pytest.mark.parametrize("shape,c", SHAPES[:4])
def test_stages_fp32_tolerance(oracle_mod, shape, c):
    o = oracle_mod
    h, w = shape
    im1, im2 = synthetic_pair(h, w, c, seed=h * w)
    u, v, wt = fields(shape)
    tol = 2e-5
    for a, b in zip(G.pyramid(im1, 0.75, 3, G.F32), o.pyramid(im1, 0.75, 3)):
        assert rel(a, b) < tol
    f1 = o.im2feature(im1); f2 = o.im2feature(im2)
    assert rel(G.im2feature(im1, G.F32), f1) < tol
    for a, b in zip(G.getdxs(f1, f2, G.F32), o.getdxs(f1, f2)):
        assert np.abs(a - b).max() < tol
    assert rel(G.warpfl(f1, f2, u, v, G.F32), o.warpfl(f1, f2, u, v)) < tol
    assert rel(G.bicubic(im1, im2, u, v, G.F32), o.bicubic_warp(im1, im2, u, v)) < 1e-4


def test_stage_fixtures_from_reference_fp64():
    """Straight against the committed dumps of the reference's public statics."""
    g = golden("stages_96x64.npz")
    for k, p in enumerate(G.pyramid(g["im1"], 0.75, 5, G.F64)):
        assert np.array_equal(p, g["pyr%d" % k])
    assert np.array_equal(G.im2feature(g["im1"], G.F64), g["f1"])
    dx, dy, dt = G.getdxs(g["f1"], g["f2"], G.F64)
    assert np.array_equal(dx, g["imdx"]) and np.array_equal(dy, g["imdy"]) and np.array_equal(dt, g["imdt"])
    assert np.array_equal(G.warpfl(g["f1"], g["f2"], g["u"], g["v"], G.F64), g["warp"])
    assert np.array_equal(G.resize_to(g["u"], 85, 128, 1 / 0.75, G.F64)[..., 0], g["up"][..., 0])
    assert np.array_equal(G.bicubic(g["im1"], g["im2"], g["u"], g["v"], G.F64), g["bicubic"])


@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (19, 34), (31, 33), (33, 65), (45, 81), (70, 100), (107, 192), (131, 67), (192, 341)])
def test_sor_lexicographic_band_march_equals_grid_wavefront(shape, monkeypatch):
    """k_sor_lex (time-skewed band march: one warp per sweep, CTAs handing du/dv over through the planes with progress
    words) against k_sor_wavefront (one grid-wide barrier per anti-diagonal), which test_assemble_and_sor_fp64 pins
    to the oracle: the SAME bits in FP64 for every combination of one / several bands and one / several sweep groups,
    images smaller than a band, odd sizes, and more sweeps than rows (S/OpticalFlow.cpp:451-505)."""
    h, w = shape
    r = np.random.default_rng(h * 1000 + w)
    sysm = (r.random((h, w)) * 40 + 0.5, r.standard_normal((h, w)) * 0.2, r.random((h, w)) + 0.05, r.random((h, w)) + 0.05,
            r.standard_normal((h, w)), r.standard_normal((h, w)))
    for nsor in ((1, 7, 8, 9, 17, 30, 72) if h * w < 30000 else (9, 30)):
        monkeypatch.setenv("PF_LEX_IMPL", "coop")
        ru, rv = G.sor(*sysm, 0.012, nsor, G.F64)
        monkeypatch.setenv("PF_LEX_IMPL", "band")
        gu, gv = G.sor(*sysm, 0.012, nsor, G.F64)
        assert np.array_equal(gu, ru) and np.array_equal(gv, rv), (shape, nsor)
        monkeypatch.setenv("PF_LEX_IMPL", "coop")
        r32 = G.sor(*sysm, 0.012, nsor, G.F32LEX)
        monkeypatch.setenv("PF_LEX_IMPL", "band")
        g32 = G.sor(*sysm, 0.012, nsor, G.F32LEX)
        if np.isfinite(r32[0]).all() and np.abs(r32[0]).max() < 1e3:   # random systems may diverge under omega = 1.8
            assert np.allclose(g32[0], r32[0], rtol=0, atol=2e-3 * max(1.0, np.abs(r32[0]).max()))
