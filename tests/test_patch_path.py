"""GPU parity of the PATCH layout (csrc/patch.cu: gathered sample lists per 64 x 16 pixel patch, K_B tiles of [128 pixels][32 slots],
extrapolation + filter in one tcgen05 kernel with a gathered MN-major W operand) -- the default path of the bilateral and spatial
affinities -- against the CPU oracle and against the blocked layout (csrc/affinity.cu + nystroem_gemm.cu), which computes the same
thing another way.  Tolerances as everywhere: eigenvalues 1e-4, z 1e-3, z - y 5e-3 (north_star + SURVEY H6)."""
import numpy as np
import pytest

import ipgl_b200 as gl
from oracle import oracle_c as oc
from oracle import oracle_np as o

pytestmark = pytest.mark.gpu
TOL_MU, TOL_Z, TOL_DZ = 1e-4, 1e-3, 5e-3


@pytest.fixture(scope="module")
def ctx():
    c = gl.Context(0)
    yield c
    c.close()


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b))


def _samples(W, H, p, method, seed=5):
    return oc.random_sampling(W, H, p, seed) if method == "random" else oc.uniform_sampling(W, H, p)


# shapes: ragged width (not a multiple of 64) and height (not a multiple of 16 / odd), several slot blocks per patch (small h_loc
# keeps one, the default h_loc = 40 on a small image needs many: more than 4 blocks streams the W rows per M tile), colour, spatial
CASES = [
    (301, 203, 1, 120, "bilateral", "random", 40.0),
    (640, 160, 1, 200, "bilateral", "random", 6.0),
    (193, 97, 1, 90, "spatial", "random", 9.0),
    (256, 131, 3, 90, "bilateral", "spatially_uniform", 12.0),
    (450, 300, 1, 400, "bilateral", "random", 40.0),
    (130, 70, 3, 60, "spatial", "random", 40.0),
]


@pytest.mark.parametrize("W,H,ch,p,kind,method,h_loc", CASES)
def test_patch_kb_and_sums_match_oracle(ctx, W, H, ch, p, kind, method, h_loc):
    img = o.synthetic_image(W, H, ch, seed=21)
    s = _samples(W, H, p, method)
    ctx.set_image(img)
    ctx.set_samples(s)
    K_A, K_B = ctx.affinity(kind, h_loc=h_loc)
    info = K_B.info
    assert info.layout == 1 and info.stored_pairs > 0
    kb = K_B.download()                                               # [pixels][p], from the patch tiles
    ref = oc.affinity_rows(img, s, np.arange(W * H, dtype=np.uint32), kind, h_loc, 30.0)      # [p][pixels]
    assert np.max(np.abs(kb.T - ref)) < 6e-4                          # fp16 storage; entries below 2^-25 are dropped
    D = K_B.rowsums()
    refD = ref.sum(axis=1)
    assert np.max(np.abs(D - refD) / refD) < 1e-5
    assert np.max(np.abs(K_A.download() - ref[:, s.astype(np.int64)])) < 1e-12
    # the blocked layout stores the same matrix
    ctx.set_option("kb_layout", "blocked")
    try:
        _, K_B2 = ctx.affinity(kind, h_loc=h_loc)
        assert K_B2.info.layout == 0
        kb2 = K_B2.download()
        assert np.max(np.abs(kb2 - kb)) < 5e-4                        # at most one fp16 ulp apart (the exponent is summed in another order)
        assert np.max(np.abs(K_B2.rowsums() - D) / D) < 1e-6
    finally:
        ctx.set_option("kb_layout", "patch")


@pytest.mark.parametrize("W,H,ch,p,kind,method,h_loc", CASES)
def test_patch_pipeline_matches_oracle_and_blocked(ctx, W, H, ch, p, kind, method, h_loc):
    img = o.synthetic_image(W, H, ch, seed=21)
    s = _samples(W, H, p, method)
    ctx.set_image(img)
    prm = gl.default_params(affinity=kind, sampling=method, sample_size=p, seed=5, h_loc=h_loc)
    shape = (H, W) if ch == 1 else (H, W, ch)
    z = np.zeros(shape, np.float32)
    r = ctx.run_resident(prm, z_out=z, want_eigvals=True)
    assert np.array_equal(ctx.get_samples(), s)
    ref = o.run_pipeline(img, s, kind=kind, h_loc=h_loc)
    err_mu = float(np.max(np.abs(r["mu"] - ref["mu"]) / ref["mu"]))
    err_z, err_dz = _rel(z, ref["z"]), _rel(z.astype(np.float64) - img, ref["z"] - img)
    print(f"patch {W}x{H}x{ch} p={p} {kind} h_loc={h_loc}: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_mu <= TOL_MU and err_z <= TOL_Z and err_dz <= TOL_DZ
    ctx.set_option("kb_layout", "blocked")
    try:
        z2 = np.zeros(shape, np.float32)
        r2 = ctx.run_resident(prm, z_out=z2, want_eigvals=True)
    finally:
        ctx.set_option("kb_layout", "patch")
    assert np.max(np.abs(r2["mu"] - r["mu"]) / r["mu"]) < 1e-6
    assert _rel(z2, z.astype(np.float64)) < 2e-5
    assert _rel(z2.astype(np.float64) - img, z.astype(np.float64) - img) < 2e-3
    # bit-reproducible
    z3 = np.zeros(shape, np.float32)
    ctx.run_resident(prm, z_out=z3)
    assert np.array_equal(z3, z)


@pytest.mark.parametrize("m", [20, 64, 100, 128, 200, 256, 300])
def test_patch_narrow_and_wide_spectra(ctx, m):
    """-num_eigvals m: N tiles of 64, 128 and 256 columns, one or several of them."""
    W, H, p = 320, 200, 320
    img = o.synthetic_image(W, H, 1, seed=4)
    s = oc.random_sampling(W, H, p, 9)
    ctx.set_image(img)
    prm = gl.default_params(sampling=gl.RANDOM, sample_size=p, seed=9, num_eigvals=m, h_loc=15.0)
    z = np.zeros((H, W), np.float32)
    r = ctx.run_resident(prm, z_out=z, want_eigvals=True)
    ref = o.run_pipeline(img, s, m=m, h_loc=15.0)
    assert r["m"] == m
    assert float(np.max(np.abs(r["mu"] - ref["mu"]) / ref["mu"])) <= TOL_MU
    assert _rel(z, ref["z"]) <= TOL_Z and _rel(z.astype(np.float64) - img, ref["z"] - img) <= TOL_DZ


def test_patch_staged_calls_and_lazy_blocked_storage(ctx):
    """The reference's stage sequence through the ABI: gl_affinity hands out the patch layout, gl_filter on the deferred Phi runs the
    fused patch kernel; a consumer that needs Phi itself (download) gets the blocked storage computed on demand."""
    W, H, p = 384, 160, 150
    img = o.synthetic_image(W, H, 1, seed=12)
    s = oc.random_sampling(W, H, p, 1)
    ctx.set_image(img)
    ctx.set_samples(s)
    K_A, K_B = ctx.affinity()
    assert K_B.info.layout == 1
    L_A, L_B = ctx.laplacian(K_A, K_B)
    U, mu, mu_inv = ctx.eigensolve(L_A, -1)
    phi = ctx.nystroem(L_B, U, mu_inv)
    z = ctx.filter(phi, mu).astype(np.float64)
    ref = o.run_pipeline(img, s, return_phi=True)
    assert _rel(z, ref["z"]) <= TOL_Z and _rel(z - img, ref["z"] - img) <= TOL_DZ
    P = phi.download()                                               # materialises Phi through the blocked path
    Ud = U.download()
    assert P.shape == (W * H, p - 1)
    assert np.max(np.abs(P[s.astype(np.int64)] - Ud)) <= 2 ** -11 * np.max(np.abs(Ud))      # sample rows are Phi_A
    y = img.reshape(-1).astype(np.float64)
    assert _rel(P @ (P.T @ y), ref["phi"] @ (ref["phi"].T @ y)) < 2e-2                       # projector parity (signs are arbitrary)
    z2 = ctx.filter(phi, mu).astype(np.float64)                      # now from the stored Phi
    assert _rel(z2, z) < 2e-5


@pytest.mark.parametrize("W,H,p,m,h_loc", [(640, 353, 400, -1, 30.0), (450, 300, 300, 256, 40.0), (700, 90, 500, 700, 8.0)])
def test_patch_dual_pipelines_match_the_general_kernel(ctx, W, H, p, m, h_loc):
    """Option pt_dual: the extrapolation as two pipelines per SM (k_patch_nystroem_dual: every patch resident, one channel) against
    the general kernel on the same K_B -- same MMAs, same order of the row dots, so the same bits; N tiles of 256 columns (one and
    several), odd M-tile counts at the band's lower edge, patches with two slot blocks (h_loc 40 on a small image)."""
    img = o.synthetic_image(W, H, 1, seed=21)
    ctx.set_image(img)
    prm = gl.default_params(sampling=gl.RANDOM, sample_size=p, seed=2, num_eigvals=m, h_loc=h_loc)
    z1, z0 = np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
    ctx.run_resident(prm, z_out=z1)
    ctx.set_option("pt_dual", 0)
    try:
        ctx.run_resident(prm, z_out=z0)
    finally:
        ctx.set_option("pt_dual", 1)
    assert np.array_equal(z1, z0)
    s = oc.random_sampling(W, H, p, 2)
    ref = o.run_pipeline(img, s, m=None if m < 0 else m, h_loc=h_loc)
    assert _rel(z1, ref["z"]) <= TOL_Z and _rel(z1.astype(np.float64) - img, ref["z"] - img) <= TOL_DZ


def test_u8_result_straight_into_pinned_host_memory(ctx):
    """gl_run with a u8 destination in pinned host memory: the fused patch kernel writes the filtered bytes itself (option z8_direct,
    no device-to-host copy afterwards); same bytes as through a pageable destination, which takes the copy."""
    W, H, p = 640, 353, 400
    img = o.synthetic_image(W, H, 1, seed=33)
    prm = gl.default_params(sampling=gl.RANDOM, sample_size=p, seed=5, h_loc=30.0)
    z8_page = np.zeros((H, W), np.uint8)
    ctx.run(img, prm, z_out=False, z8_out=z8_page, want_eigvals=False)
    pin = gl.PinnedArray((H, W), np.uint8)
    try:
        pin.array[...] = 7
        ctx.run(img, prm, z_out=False, z8_out=pin.array, want_eigvals=False)
        assert np.array_equal(pin.array, z8_page)
        ctx.set_option("z8_direct", 0)
        pin.array[...] = 9
        ctx.run(img, prm, z_out=False, z8_out=pin.array, want_eigvals=False)
        assert np.array_equal(pin.array, z8_page)
    finally:
        ctx.set_option("z8_direct", 1)
        pin.free()
    zf = np.zeros((H, W), np.float32)
    ctx.run(img, prm, z_out=zf, want_eigvals=False)
    assert np.array_equal(z8_page, np.clip(zf, 0, 255).astype(np.uint8))
