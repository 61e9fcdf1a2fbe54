"""Parity at BASELINE.json's full size (config 4: synthetic 3840x2160 grey, p=1000 random samples, m=999) on the GPU:
against the compact full-size golden written by the C/OpenMP fp64 oracle (tests/golden/make_golden_c4.py), and through
size-independent properties of the filter (gain 0 is the identity, the change z - y is linear in the gain, the spatial
cutoff and the fused epilogue do not change the answer, the run is bit-reproducible)."""
import os

import numpy as np
import pytest

import ipgl_b200 as gl
from oracle import oracle_np as o

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c4_full.npz")
W, H, P = 3840, 2160, 1000


@pytest.fixture(scope="module")
def ctx():
    c = gl.Context(0)
    c.set_synthetic_image(W, H, 1, 1234)
    yield c
    c.close()


def _run(ctx, **kw):
    prm = gl.default_params(sampling=gl.RANDOM, sample_size=P, seed=0, **kw)
    z = np.zeros((H, W), dtype=np.float32)
    r = ctx.run_resident(prm, z_out=z, want_eigvals=True)
    r["z"] = z
    return r


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def base(ctx):
    return _run(ctx)


def test_c4_against_full_size_oracle(ctx, base):
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/c4_full.npz not generated")
    g = np.load(GOLD)
    img = ctx.get_image()
    assert np.array_equal(img[::97, ::89], o.synthetic_image(W, H, 1, 1234)[::97, ::89])     # same input as the oracle saw
    assert np.array_equal(ctx.get_samples(), g["sample_indices"])                              # bit-exact samples
    assert base["p"] == P and base["m"] == P - 1
    err_mu = float(np.max(np.abs(base["mu"] - g["mu"]) / g["mu"]))
    idx = np.arange(0, W * H, int(g["stride"]))
    y = img.reshape(-1)[idx].astype(np.float64)
    z = base["z"].reshape(-1)[idx].astype(np.float64)
    zr = g["z_lattice"].astype(np.float64)
    err_z, err_dz = _rel(z, zr), _rel(z - y, zr - y)
    err_sum = abs(float(base["z"].astype(np.float64).sum()) - float(g["sum_z"])) / float(g["sum_z"])
    dz2 = float(((base["z"].astype(np.float64) - img) ** 2).sum())
    print(f"C4 full size: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e} err_sum={err_sum:.2e} "
          f"|z-y|^2 {dz2:.6e} vs {float(g['sum_dz2']):.6e}")
    assert err_mu <= 1e-4            # north_star: leading eigenvalues rel <= 1e-4
    assert err_z <= 1e-3             # north_star: filtered image rel L2 <= 1e-3
    assert err_dz <= 5e-3
    assert err_sum <= 1e-6
    assert abs(dz2 - float(g["sum_dz2"])) <= 1e-2 * float(g["sum_dz2"])


def test_c4_properties(ctx, base):
    img = ctx.get_image().astype(np.float64)
    mu = base["mu"]
    assert np.all(np.diff(mu) >= 0) and mu[0] > 0                     # ascending, L_A is positive definite
    assert base["z"].max() <= 255.0
    # gain 0: the filter is the identity
    z0 = _run(ctx, gain=0.0)["z"]
    assert np.array_equal(z0, img.astype(np.float32))
    # the change is linear in the gain wherever nothing was clipped
    z1 = _run(ctx, gain=1.5)["z"].astype(np.float64)
    z3 = base["z"].astype(np.float64)
    free = (z3 < 254.5) & (z1 < 254.5)
    assert _rel(2.0 * (z1 - img)[free], (z3 - img)[free]) < 1e-3
    # bit-reproducible
    again = _run(ctx)
    assert np.array_equal(again["z"], base["z"]) and np.array_equal(again["mu"], base["mu"])


def test_c4_variants_agree(ctx, base):
    ref = base["z"].astype(np.float64)
    img = ctx.get_image().astype(np.float64)
    # keep_phi=1: Phi written to HBM as well, which takes the blocked layout of K_B (default: patch layout, Phi consumed in the
    # epilogue only); kb_layout=blocked: the blocked layout with the fused filter; fuse_filter=0: stages apart; kb_cutoff=0: dense K_B
    for key, val, back in (("keep_phi", 1, 0), ("kb_layout", "blocked", "patch"), ("fuse_filter", 0, 1), ("kb_cutoff", 0, 1)):
        ctx.set_option(key, val)
        try:
            r = _run(ctx)
        finally:
            ctx.set_option(key, back)
        assert _rel(r["z"], ref) < 2e-5, key
        assert _rel(r["z"] - img, ref - img) < 2e-3, key
        assert np.max(np.abs(r["mu"] - base["mu"]) / base["mu"]) < 1e-6, key


def test_c5_against_full_size_oracle():
    """BASELINE.json config 5 (synthetic 8192x8192 colour, p=2000 random samples, m=1999) on ONE GPU (Phi would be 275 GB: it is
    consumed tile by tile and never stored), against the compact golden written by the CPU oracle (tests/golden/make_golden_c5.py)."""
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c5_full.npz")
    if not os.path.exists(gold):
        pytest.skip("tests/golden/c5_full.npz not generated")
    g = np.load(gold)
    W5, H5, C5, P5 = int(g["width"]), int(g["height"]), int(g["channels"]), int(g["p"])
    c = gl.Context(0)
    try:
        c.set_synthetic_image(W5, H5, C5, int(g["seed_img"]))
        prm = gl.default_params(sampling=gl.RANDOM, sample_size=P5, seed=int(g["seed_samples"]))
        z = np.zeros((H5, W5, C5), dtype=np.float32)
        r = c.run_resident(prm, z_out=z, want_eigvals=True)
        assert np.array_equal(c.get_samples(), g["sample_indices"])
        assert r["p"] == P5 and r["m"] == P5 - 1
        err_mu = float(np.max(np.abs(r["mu"] - g["mu"]) / g["mu"]))
        n = W5 * H5
        idx = np.arange(0, n, int(g["stride"]))
        img = c.get_image().reshape(n, C5)
        y = img[idx].astype(np.float64)
        zz = z.reshape(n, C5)[idx].astype(np.float64)
        zr = g["z_lattice"].astype(np.float64)
        err_z, err_dz = _rel(zz, zr), _rel(zz - y, zr - y)
        err_sum = abs(float(z.astype(np.float64).sum()) - float(g["sum_z"])) / float(g["sum_z"])
        print(f"C5 full size: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e} err_sum={err_sum:.2e}")
        assert err_mu <= 1e-4 and err_z <= 1e-3 and err_dz <= 5e-3 and err_sum <= 1e-6
    finally:
        c.close()
