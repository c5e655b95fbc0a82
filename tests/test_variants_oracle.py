"""Alternative solver branches (Bicubic inner warp, Gaussian-mixture noise model): the plain-C
oracle against the golden vectors recorded from the unmodified reference (tests/golden/make_golden_variants.py) and,
when oracle/_ref is present, against the reference live.  CPU only."""
import numpy as np
import pytest

from conftest import golden, load_frame


def crop():
    a, b = load_frame(240, 1), load_frame(240, 2)
    return np.ascontiguousarray(a[:96, :128]), np.ascontiguousarray(b[:96, :128])


@pytest.fixture()
def variant(oracle_mod):
    yield oracle_mod.set_variant
    oracle_mod.set_variant("bilinear", "lap")


def test_oracle_bicubic_inner_warp_matches_reference_golden(oracle_mod, variant):
    g = golden("variants_128x96.npz")
    a, b = crop()
    variant("bicubic", "lap")
    vx, vy, wi = oracle_mod.coarse2fine_flow(a, b, levels=4)
    assert np.array_equal(vx, g["bicubic_fork_vx"]) and np.array_equal(vy, g["bicubic_fork_vy"])
    assert np.array_equal(wi, g["bicubic_fork_warp"])
    vx, vy, wi = oracle_mod.coarse2fine_flow(a, b, 0.012, 0.75, 20, 5, 1, 20, 0)
    assert np.array_equal(vx, g["bicubic_up_vx"]) and np.array_equal(vy, g["bicubic_up_vy"])
    assert np.array_equal(wi, g["bicubic_up_warp"])
    ag, bg = np.ascontiguousarray(a.mean(axis=2, keepdims=True)), np.ascontiguousarray(b.mean(axis=2, keepdims=True))
    vx, vy, wi = oracle_mod.coarse2fine_flow(ag, bg, 0.012, 0.75, 20, 4, 2, 15, 1)
    assert np.array_equal(vx, g["bicubic_gray_vx"]) and np.array_equal(vy, g["bicubic_gray_vy"])
    assert np.array_equal(wi, g["bicubic_gray_warp"])
    # the branch is really taken: the bilinear result differs
    variant("bilinear", "lap")
    vx2, _, _ = oracle_mod.coarse2fine_flow(a, b, levels=4)
    assert np.abs(vx2 - g["bicubic_fork_vx"]).max() > 1e-3


@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("tag,mw,no", [("l1o3", 90, 3), ("l2o2", 70, 2)])
def test_oracle_gaussian_mixture_matches_reference_golden(oracle_mod, variant, interp, tag, mw, no):
    g = golden("variants_128x96.npz")
    a, b = crop()
    variant(interp, "gmixture")
    vx, vy, _ = oracle_mod.coarse2fine_flow(a, b, 0.012, 0.75, mw, no, 1, 10, 0)
    k = "gmix_%s_%s_" % (interp, tag)
    assert np.array_equal(vx, g[k + "vx"]) and np.array_equal(vy, g[k + "vy"])
    al, sg, be = oracle_mod.gm_get(5)
    assert np.array_equal(al, g[k + "alpha"]) and np.array_equal(sg, g[k + "sigma"]) and np.array_equal(be, g[k + "beta"])


def test_oracle_variants_match_reference_live(oracle_mod, variant, ref_serial):
    a, b = crop()
    try:
        for interp, noise in (("bicubic", "lap"), ("bilinear", "gmixture"), ("bicubic", "gmixture")):
            ref_serial.set_variant(interp, noise)
            variant(interp, noise)
            _, rx, ry, rw = ref_serial.coarse2fine_flow_levels(a, b, 3)
            ox, oy, ow = oracle_mod.coarse2fine_flow(a, b, levels=3)
            assert np.array_equal(rx, ox) and np.array_equal(ry, oy) and np.array_equal(rw, ow), (interp, noise)
    finally:
        ref_serial.set_variant("bilinear", "lap")
