"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI against the golden
fixtures (reference python modules + oracle) and against the oracle on seeded synthetic inputs.

Tolerances (BASELINE.json north_star): sampled indices bit-exact; leading eigenvalues rel <= 1e-4;
filtered image rel L2 <= 1e-3.  The filtered image differs from the input by only ~0.2-1 %, so the
tests additionally bound the relative error of the CHANGE z - y."""
import os

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


def _src(g):
    img = g["image"]
    return np.repeat(img[:, :, None], 3, axis=2) if int(g["rgb"]) else img


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b))


# ---------------------------------------------------------------------------------------------
# a-1 sampling: bit-exact with the reference's python modules
# ---------------------------------------------------------------------------------------------
def test_sampling_bit_exact(ctx, golden):
    tab = golden("sampling")
    for key, ref in tab.items():
        parts = key.split("_")
        W, H = (int(v) for v in parts[1].split("x"))
        p = int(parts[2])
        ctx.set_synthetic_image(W, H, 1, 1)
        if parts[0] == "uniform":
            got = ctx.sampling(gl.SPATIALLY_UNIFORM, p)
        else:
            got = ctx.sampling(gl.RANDOM, p, seed=int(parts[3][1:]))
        assert got.dtype == np.uint32 and np.array_equal(got, ref), key


def test_random_sampling_with_collisions(ctx):
    # tiny image, many samples: duplicates in the stream are certain, the top-up path must run
    for (W, H, p, seed) in ((64, 64, 3000, 3), (40, 30, 1100, 11), (128, 64, 4000, 5)):
        ctx.set_synthetic_image(W, H, 1, 1)
        assert np.array_equal(ctx.sampling(gl.RANDOM, p, seed=seed), oc.random_sampling(W, H, p, seed))


def test_synthetic_image_matches_oracle(ctx):
    for (W, H, ch) in ((333, 222, 1), (257, 131, 3)):
        ctx.set_synthetic_image(W, H, ch, 1234)
        assert np.array_equal(ctx.get_image(), o.synthetic_image(W, H, ch, 1234))


# ---------------------------------------------------------------------------------------------
# a-2 / a-3 affinity + Laplacian against the reference python modules' K
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["test_uniform100", "cat_small_random50", "lion_photometric_h10", "test_spatial_h10",
                                 "barbara_uniform256"])
def test_affinity_and_laplacian(ctx, golden, tag):
    g = golden(tag)
    img, s = g["image"], g["sample_indices"]
    kind, h_loc, h_val = str(g["kind"]), float(g["h_loc"]), float(g["h_val"])
    ctx.set_image(img)
    ctx.set_samples(s)
    K_A, K_B = ctx.affinity(kind, h_loc, h_val)
    if "ref_K_A" in g:
        assert np.allclose(K_A.download(), g["ref_K_A"], rtol=1e-12, atol=1e-300)   # fp64, same formula as the reference
    D = K_B.rowsums()
    assert np.max(np.abs(D - g["ref_D"]) / g["ref_D"]) < 2e-6     # fp32 ex2 kernel values, fp64 final sum
    # K_B itself (fp16, stored pixel-major): compare a slab against the oracle
    n = img.size
    KB = K_B.download()
    assert KB.shape == (n, len(s))
    cols = np.arange(0, n, max(1, n // 4096))
    ref = o.affinity_rows(img, s, cols, kind, h_loc, h_val).T
    assert np.max(np.abs(KB[cols] - ref)) < 6e-4                  # fp16 rounding of values in [0,1]
    L_A, L_B = ctx.laplacian(K_A, K_B)
    alpha = 1.0 / g["ref_D"].mean()
    assert abs(L_B.info.scale + alpha) < 1e-6 * alpha             # L_B = -alpha K_B, no copy
    LA = L_A.download()
    LA_ref = alpha * (np.diag(g["ref_D"]) - o.affinity_rows(img, s, s, kind, h_loc, h_val))
    assert np.max(np.abs(LA - LA_ref)) < 3e-6 * np.max(np.abs(LA_ref))


# ---------------------------------------------------------------------------------------------
# a-4/5 eigensolver alone
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", [17, 50, 256, 529, 1000])
def test_eigensolver(ctx, p):
    rng = np.random.RandomState(p)
    # SPD, diagonally dominant-ish like L_A, with a few close eigenvalues
    B = rng.rand(p, p) * np.exp(-np.abs(np.subtract.outer(np.arange(p), np.arange(p))) / 3.0)
    A = np.diag(1.0 + rng.rand(p)) + 0.2 * (B + B.T) / 2
    A += np.eye(p) * max(0.0, 0.05 - np.linalg.eigvalsh(A)[0])
    L = ctx.upload(gl.MAT_KA, A)
    m = p - 1
    U, mu, mu_inv = ctx.eigensolve(L, m)
    w, V = np.linalg.eigh(A)
    got = mu.download()
    assert got.shape == (m,)
    assert np.all(np.diff(got) >= 0)
    err_mu = float(np.max(np.abs(got - w[:m]) / w[:m]))
    assert err_mu < 1e-7, err_mu                      # fp64 Rayleigh quotients on fp32 vectors (north_star asks for 1e-4)
    assert np.allclose(mu_inv.download(), 1.0 / got, rtol=1e-12)
    Ug = U.download()
    assert Ug.shape == (p, m)
    err_orth = float(np.max(np.abs(Ug.T @ Ug - np.eye(m))))
    err_res = float(np.max(np.abs(A @ Ug - Ug * got)) / np.max(np.abs(w)))
    assert err_orth < 1e-4 and err_res < 1e-4, (err_orth, err_res)   # the solver stops at a relative off-diagonal of 5e-5


# ---------------------------------------------------------------------------------------------
# whole path against the golden fixtures (BASELINE.json configs 1-3 and friends)
# ---------------------------------------------------------------------------------------------
CASES = ["test_uniform100", "cat_small_random50", "barbara_uniform256", "lion_rgb_photometric500",
         "lion_photometric_h10", "test_spatial_h10", "cat_small_uniform_m20"]


def _run_case(ctx, g, **over):
    src = _src(g)
    prm = gl.default_params(affinity=str(g["kind"]), sampling=gl.RANDOM if str(g["method"]) == "random" else gl.SPATIALLY_UNIFORM,
                            h_loc=float(g["h_loc"]), h_val=float(g["h_val"]), sample_size=int(g["p_req"]),
                            seed=int(g["seed"]), num_eigvals=int(g["m"]), **over)
    return src, ctx.run(src, prm)


@pytest.mark.parametrize("tag", CASES)
def test_pipeline_matches_golden(ctx, golden, tag):
    g = golden(tag)
    src, r = _run_case(ctx, g)
    assert r["p"] == len(g["sample_indices"]) and r["m"] == int(g["m"])
    assert np.array_equal(ctx.get_samples(), g["sample_indices"])              # bit-exact
    mu = g["mu"]
    err_mu = float(np.max(np.abs(r["mu"] - mu) / mu))
    z = g["z"].astype(np.float64)
    err_z, err_dz = _rel(r["z"], z), _rel(r["z"] - src, z - src)
    print(f"{tag}: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_mu <= TOL_MU, err_mu
    assert err_z <= TOL_Z, err_z
    assert err_dz <= TOL_DZ, err_dz
    assert r["z"].max() <= 255.0


@pytest.mark.parametrize("tag_py", ["pyref_test100", "pyref_lion_crop", "pyref_lion_crop_photometric", "pyref_test100_spatial",
                                    "pyref_lion_crop_random7"])
def test_against_the_reference_python_pipeline(ctx, golden, tag_py):
    """The CUDA path against the output of the reference's OWN image_processing(y) (python/image_processing.py:244-357;
    fixtures from tests/golden/make_golden_pyref.py): 1 % uniform samples, bilateral affinity, ALL p eigenpairs,
    z = y - Phi (mu + 5) Phi^T y.  f(mu) = mu + 5 is not a power, so the image is assembled from two filter applications
    on the same Phi: z = [y - Phi mu Phi^T y] + [y - 5 Phi Phi^T y] - y."""
    g = golden(tag_py)
    img, s, kind, seed = g["image"], g["sample_indices"], str(g["kind"]), int(g["seed"])
    # bandwidths hard-coded in the reference's plugins: bilateral 30 / 40, photometric 10, spatial 10
    h_loc, h_val = {"bilateral": (40.0, 30.0), "photometric": (40.0, 10.0), "spatial": (10.0, 30.0)}[kind]
    H, W = img.shape
    ctx.set_image(img)
    got = (ctx.sampling(gl.SPATIALLY_UNIFORM, int(W * H * 0.01)) if seed < 0
           else ctx.sampling(gl.RANDOM, int(W * H * 0.01), seed=seed))
    assert np.array_equal(got, s)                                   # the reference module's own sample list, bit for bit
    K_A, K_B = ctx.affinity(kind, h_loc, h_val)
    L_A, L_B = ctx.laplacian(K_A, K_B)
    U, mu, mu_inv = ctx.eigensolve(L_A, len(s))                     # every pair, like the prototype
    assert mu.info.rows == len(s)
    phi = ctx.nystroem(L_B, U, mu_inv)
    z1 = ctx.filter(phi, ctx.diag_pow(mu, 1.0), gain=-1.0).astype(np.float64)
    z0 = ctx.filter(phi, ctx.diag_pow(mu, 0.0), gain=-5.0).astype(np.float64)
    z = z1 + z0 - img
    ref = g["z"]
    free = (z1 < 254.9) & (z0 < 254.9) & (ref < 254.9)              # the C filter clips above 255 (display.c:76), Python does not
    assert free.mean() > 0.5                                       # test.png has a white (255) background: those pixels clip
    err_z = float(np.linalg.norm((z - ref)[free]) / np.linalg.norm(ref[free]))
    err_dz = float(np.linalg.norm(((z - img) - (ref - img))[free]) / np.linalg.norm((ref - img)[free]))
    print(f"{tag_py}: against the reference python pipeline err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_z <= TOL_Z and err_dz <= TOL_DZ


@pytest.mark.parametrize("tag", ["test_uniform100", "cat_small_random50"])
def test_tcgen05_gemm_matches_cuda_core_checker(ctx, golden, tag):
    g = golden(tag)
    ctx.set_option("gemm", "simple")
    try:
        _, a = _run_case(ctx, g)
    finally:
        ctx.set_option("gemm", "tcgen05")
    _, b = _run_case(ctx, g)
    assert _rel(b["z"], a["z"].astype(np.float64)) < 1e-6


@pytest.mark.parametrize("tag", ["test_uniform100", "cat_small_random50", "cat_small_uniform_m20"])
def test_gram_schmidt_stage(ctx, golden, tag):
    g = golden(tag)
    src, r = _run_case(ctx, g, gram_schmidt=1)
    z = g["z_gs"].astype(np.float64)
    err_z, err_dz = _rel(r["z"], z), _rel(r["z"] - src, z - src)
    print(f"gs {tag}: err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_z <= TOL_Z, err_z
    assert err_dz <= TOL_DZ, err_dz


@pytest.mark.parametrize("gram", ["tcgen05", "simple"])
def test_gram_schmidt_wide_phi(ctx, gram):
    """m_pad = 256: the Gram matrix comes from the tcgen05 kernel with MN-major operands (option gram=simple: the
    CUDA-core tiles); blocked Cholesky + triangular inverse; against the oracle's column-by-column Gram-Schmidt."""
    W, H, p = 320, 200, 200
    img = o.synthetic_image(W, H, 1, seed=11)
    s = oc.random_sampling(W, H, p, 4)
    ref = o.run_pipeline(img, s, orthonormalise=True, return_phi=True)
    _, nref = o.gram_schmidt(o.run_pipeline(img, s, return_phi=True)["phi"])
    ctx.set_option("gram", gram)
    try:
        ctx.set_image(img)
        ctx.set_samples(s)
        K_A, K_B = ctx.affinity()
        L_A, L_B = ctx.laplacian(K_A, K_B)
        U, mu, mu_inv = ctx.eigensolve(L_A, -1)
        phi = ctx.nystroem(L_B, U, mu_inv)
        assert phi.info.ld == 256
        norms = ctx.orthonormalise(phi)
        Q = phi.download()
        z = ctx.filter(phi, mu)
    finally:
        ctx.set_option("gram", "tcgen05")
    err_q = float(np.max(np.abs(Q.T @ Q - np.eye(Q.shape[1]))))
    err_n = float(np.max(np.abs(norms - nref) / nref))
    err_z, err_dz = _rel(z, ref["z"]), _rel(z - img, ref["z"] - img)
    print(f"gs wide ({gram}): orth={err_q:.2e} norms={err_n:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_q < 1e-3 and err_n < 2e-3
    assert err_z <= TOL_Z and err_dz <= TOL_DZ


@pytest.mark.parametrize("name", ["sq24", "rect"])
def test_nlm_affinity_against_the_reference_module(ctx, name):
    """The NLM patch affinity (python/affinity_methods/NLM.py) on the device against the reference module's own output
    (tests/golden/pyref_nlm_*.npz, columns mapped from its column-major pixel order to raster order, see
    tests/test_oracle.py) and against the oracle; then the whole path on it."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"pyref_nlm_{name}.npz"))
    img, s, K = g["image"], g["sample_indices"], g["K_AB"]
    M, N = img.shape
    r, c = np.divmod(np.arange(M * N), N)
    K_ref = K[:, c * M + r]                                   # [p][n] in raster columns
    ctx.set_image(img)
    ctx.set_samples(s)
    K_A, K_B = ctx.affinity(gl.NLM)                           # h = 3 (NLM.py:12)
    ka, kb, D = K_A.download(), K_B.download(), K_B.rowsums()
    assert np.max(np.abs(ka - K_ref[:, s])) < 1e-6            # fp64 on fp32 patch weights
    assert np.max(np.abs(kb.T - K_ref)) < 6e-4                # fp16 storage
    assert np.max(np.abs(D - K_ref.sum(axis=1)) / K_ref.sum(axis=1)) < 1e-5
    L_A, L_B = ctx.laplacian(K_A, K_B)
    U, mu, mu_inv = ctx.eigensolve(L_A, -1)
    z = ctx.filter(ctx.nystroem(L_B, U, mu_inv), mu).astype(np.float64)
    ref = o.run_pipeline(img, s, kind=o.NLM, h_val=3.0)
    err_mu = np.max(np.abs(mu.download() - ref["mu"]) / ref["mu"])
    err_z, err_dz = _rel(z, ref["z"]), _rel(z - img, ref["z"] - img)
    print(f"nlm {name}: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_mu <= TOL_MU and err_z <= TOL_Z and err_dz <= TOL_DZ
    # one call, through the params
    prm = gl.default_params(affinity=gl.NLM, sample_size=len(s))
    z2 = np.zeros(img.shape, np.float32)
    ctx.run_resident(prm, z_out=z2)
    assert _rel(z2, z) < 1e-5


def test_nlm_affinity_medium_image(ctx):
    """NLM on an image with several tiles and sample blocks (ragged last tile, p not a multiple of 64) against the oracle."""
    W, H, p = 173, 141, 150
    img = (o.synthetic_image(W, H, 1, seed=8) // 8 + 100).astype(np.uint8)     # low contrast: the h = 3 kernel stays alive
    s = oc.random_sampling(W, H, p, 2)
    ctx.set_image(img)
    ctx.set_samples(s)
    K_A, K_B = ctx.affinity("nlm", h_val=6.0)
    cols = np.arange(0, W * H, 7)
    ref = o.nlm_affinity_rows(img, s, cols, 6.0)
    kb = K_B.download()[cols]
    assert np.max(np.abs(kb.T - ref)) < 6e-4
    ref_D = sum(o.nlm_affinity_rows(img, s, np.arange(a, min(a + 4096, W * H)), 6.0).sum(axis=1) for a in range(0, W * H, 4096))
    assert np.max(np.abs(K_B.rowsums() - ref_D) / ref_D) < 1e-5
    assert np.max(np.abs(K_A.download() - o.nlm_affinity_rows(img, s, s, 6.0))) < 1e-6
    with pytest.raises(gl.GLError):
        ctx.full_affinity("nlm")                # the matrix-free full mode has the C program's three kinds only
    ctx.set_image(np.repeat(img[:, :, None], 3, axis=2))
    ctx.set_samples(s)
    with pytest.raises(gl.GLError):
        ctx.affinity("nlm")                     # one channel only


def test_phi_is_stored_only_on_request_and_when_it_fits(ctx):
    """gl_run consumes Phi in the GEMM epilogue without storing it (peak device memory stays below the size of Phi);
    option keep_phi=1 writes it as well (through the blocked layout of K_B: same z up to the summation order) -- unless it
    cannot be stored (config 5 on one GPU would need 275 GB; option phi_limit_mb forces that case at a small size)."""
    W, H, p = 1024, 768, 600
    img = o.synthetic_image(W, H, 1, seed=21)
    ctx.set_image(img)
    prm = gl.default_params(sampling=gl.RANDOM, sample_size=p, seed=4)
    z_a = np.zeros((H, W), np.float32)
    ctx.run_resident(prm, z_out=z_a)
    phi_bytes = W * H * 768 * 2                   # m = 599 -> 768 columns of fp16
    z_b = np.zeros((H, W), np.float32)
    ctx.memory_stats(reset_peak=True)
    r = ctx.run_resident(prm, z_out=z_b, want_eigvals=True)
    assert r["m"] == p - 1 and np.array_equal(z_a, z_b)
    assert ctx.memory_stats()["peak"] < phi_bytes    # the default does not store Phi
    ctx.set_option("keep_phi", 1)
    try:
        ctx.memory_stats(reset_peak=True)
        ctx.run_resident(prm, z_out=z_b)
        assert ctx.memory_stats()["peak"] > phi_bytes    # keep_phi=1 does
        assert _rel(z_b, z_a.astype(np.float64)) < 2e-5
        z_c = np.zeros((H, W), np.float32)
        ctx.set_option("phi_limit_mb", 64)               # ... unless it exceeds the limit
        ctx.memory_stats(reset_peak=True)
        ctx.run_resident(prm, z_out=z_c)
        assert ctx.memory_stats()["peak"] < phi_bytes and np.array_equal(z_c, z_b)   # blocked layout both times: bit for bit
    finally:
        ctx.set_option("keep_phi", 0)
        ctx.set_option("phi_limit_mb", 0)


def test_column_strip_download(ctx, golden):
    """gl_mat_download_cols (what WriteMatCol / WritePngMatCol read, hpc/display.c:85-126) against the full download."""
    g = golden("cat_small_random50")
    ctx.set_image(g["image"])
    ctx.set_samples(g["sample_indices"])
    K_A, K_B = ctx.affinity(str(g["kind"]), float(g["h_loc"]), float(g["h_val"]))
    L_A, L_B = ctx.laplacian(K_A, K_B)
    U, mu, mu_inv = ctx.eigensolve(L_A, -1)
    phi = ctx.nystroem(L_B, U, mu_inv)                   # deferred: the strip download has to materialise it
    strip = phi.download_cols(3, 4)
    full = phi.download()
    assert strip.shape == (g["image"].size, 4) and np.array_equal(strip, full[:, 3:7])
    assert np.array_equal(U.download_cols(0, 2), U.download()[:, :2])
    assert np.array_equal(L_A.download_cols(7, 1), L_A.download()[:, 7:8])
    with pytest.raises(gl.GLError):
        phi.download_cols(full.shape[1] - 1, 2)
    with pytest.raises(gl.GLError):
        K_B.download_cols(0, 1)


def test_stage_by_stage_phi_properties(ctx, golden):
    g = golden("cat_small_random50")
    img, s = g["image"], g["sample_indices"]
    ctx.set_image(img)
    ctx.set_samples(s)
    K_A, K_B = ctx.affinity()
    L_A, L_B = ctx.laplacian(K_A, K_B)
    U, mu, mu_inv = ctx.eigensolve(L_A, -1)
    phi = ctx.nystroem(L_B, U, mu_inv)
    P = phi.download()
    assert P.shape == (img.size, len(s) - 1)
    Ud = U.download()
    # sample rows of Phi are Phi_A (nystroem.c:25-34), up to fp16 storage
    assert np.max(np.abs(P[s.astype(np.int64)] - Ud)) <= 2 ** -11 * np.max(np.abs(Ud))
    G = P.T @ P
    assert np.linalg.norm(G - np.eye(G.shape[0])) < 3e-2          # nearly orthonormal before GS
    # projector parity with the oracle (eigenvector signs are arbitrary: never compare Phi entrywise)
    ref = o.run_pipeline(img, s, return_phi=True)
    y = img.reshape(-1).astype(np.float64)
    err_proj = _rel(P @ (P.T @ y), ref["phi"] @ (ref["phi"].T @ y))
    assert err_proj < 2e-2, err_proj     # Phi^T y cancels ~1e4-fold, so fp16 W shows here (the product path avoids it)
    norms = ctx.orthonormalise(phi)
    Q = phi.download()
    err_q = float(np.max(np.abs(Q.T @ Q - np.eye(Q.shape[1]))))
    assert err_q < 5e-4, err_q                                   # orthonormal up to fp16 storage
    _, nref = o.gram_schmidt(ref["phi"])
    err_n = float(np.max(np.abs(norms - nref) / nref))
    assert err_n < 2e-3, err_n
    z = ctx.filter(phi, mu)
    assert z.shape == img.shape and np.isfinite(z).all()


# ---------------------------------------------------------------------------------------------
# seeded synthetic inputs against the oracle, ragged sizes and edge cases
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("W,H,ch,p,kind,method", [
    (301, 203, 1, 77, "bilateral", "random"),       # ragged: n not a multiple of any tile
    (640, 360, 1, 300, "bilateral", "uniform"),
    (257, 129, 3, 130, "bilateral", "random"),      # colour (photometric RGB + spatial)
    (200, 150, 3, 260, "photometric", "uniform"),
    (129, 65, 1, 40, "spatial", "uniform"),
])
def test_synthetic_against_oracle(ctx, W, H, ch, p, kind, method):
    img = o.synthetic_image(W, H, ch, seed=W + H)
    prm = gl.default_params(affinity=kind, sampling=gl.RANDOM if method == "random" else gl.SPATIALLY_UNIFORM,
                            sample_size=p, seed=42)
    r = ctx.run(img, prm)
    s = oc.random_sampling(W, H, p, 42) if method == "random" else oc.uniform_sampling(W, H, p)
    assert np.array_equal(ctx.get_samples(), s)
    ref = oc.run_pipeline(img, s, kind=kind)
    err_mu = float(np.max(np.abs(r["mu"] - ref["mu"]) / ref["mu"]))
    err_z, err_dz = _rel(r["z"], ref["z"]), _rel(r["z"] - img, ref["z"] - img)
    print(f"synthetic {W}x{H}x{ch} p={p} {kind}: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_mu <= TOL_MU, err_mu
    assert err_z <= TOL_Z, err_z
    assert err_dz <= TOL_DZ, err_dz


def test_config5_shape_colour_p2000(ctx):
    """BASELINE config 5's shape on a small image: colour (5-dimensional features), p = 2000 random samples, m = 1999
    (m_pad = 2048: eight N tiles, three-channel fused filter, 250-panel eigensolve), against the fp64 CPU pipeline
    (OpenMP kernel rows + LAPACK eigh + BLAS)."""
    from oracle import cpu_pipeline as cp
    W, H, p = 480, 320, 2000
    img = o.synthetic_image(W, H, 3, seed=5)
    prm = gl.default_params(sampling=gl.RANDOM, sample_size=p, seed=9)
    r = ctx.run(img, prm)
    s = oc.random_sampling(W, H, p, 9)
    assert np.array_equal(ctx.get_samples(), s) and r["m"] == p - 1
    ref = cp.run(img, s)
    err_mu = float(np.max(np.abs(r["mu"] - ref["mu"]) / ref["mu"]))
    err_z, err_dz = _rel(r["z"], ref["z"]), _rel(r["z"] - img, ref["z"] - img)
    print(f"config-5 shape {W}x{H}x3 p={p}: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_mu <= TOL_MU and err_z <= TOL_Z and err_dz <= TOL_DZ
    # the stages apart: the wide three-channel Phi goes through the column-chunked warp apply (and the generic kernel)
    for impl in ("warp", "generic"):
        ctx.set_option("fuse_filter", 0)
        ctx.set_option("filter_apply", impl)
        try:
            r2 = ctx.run(img, prm)
        finally:
            ctx.set_option("fuse_filter", 1)
            ctx.set_option("filter_apply", "warp")
        assert _rel(r2["z"], ref["z"]) <= TOL_Z and _rel(r2["z"] - img, ref["z"] - img) <= TOL_DZ, impl
        assert _rel(r2["z"], r["z"].astype(np.float64)) < 5e-5, impl


@pytest.mark.parametrize("W,H,p,method", [
    (16, 16, 4, "uniform"),        # smaller than one 512-pixel tile, one eigenpair kept of ... 4 samples
    (33, 17, 9, "random"),
    (700, 2, 12, "random"),        # two image rows
    (2, 300, 7, "random"),         # two image columns
    (64, 64, 2, "random"),         # the minimum: p = 2, m = 1
])
def test_tiny_and_degenerate_shapes(ctx, W, H, p, method):
    img = o.synthetic_image(W, H, 1, seed=W * 7 + H)
    prm = gl.default_params(sampling=gl.RANDOM if method == "random" else gl.SPATIALLY_UNIFORM, sample_size=p, seed=5)
    r = ctx.run(img, prm)
    s = oc.random_sampling(W, H, p, 5) if method == "random" else oc.uniform_sampling(W, H, p)
    assert np.array_equal(ctx.get_samples(), s) and r["p"] == len(s) and r["m"] == len(s) - 1
    ref = o.run_pipeline(img, s)
    assert np.max(np.abs(r["mu"] - ref["mu"]) / ref["mu"]) <= TOL_MU
    assert _rel(r["z"], ref["z"]) <= TOL_Z and _rel(r["z"] - img, ref["z"] - img) <= TOL_DZ


def test_new_sample_draws_on_the_same_geometry(ctx):
    """The K_B layout is cached on (geometry, samples): a new random draw on the same image must rebuild it (with the strip
    count chosen for the geometry) and still match the oracle; going back to the first draw must reproduce it bit for bit."""
    from oracle import cpu_pipeline as cp
    W, H, p, h_loc = 1200, 300, 400, 12.0
    img = o.synthetic_image(W, H, 1, seed=8)
    lay = gl.kb_layout(W, 0, W * H, oc.random_sampling(W, H, p, 1), h_loc=h_loc)
    assert lay["strips"] > 1                                  # a wide image with a short reach: column strips are in play
    first = None
    for seed in (1, 2, 3, 1):
        prm = gl.default_params(sampling=gl.RANDOM, sample_size=p, seed=seed, h_loc=h_loc)
        r = ctx.run(img, prm)
        s = oc.random_sampling(W, H, p, seed)
        assert np.array_equal(ctx.get_samples(), s)
        ref = cp.run(img, s, h_loc=h_loc)
        assert np.max(np.abs(r["mu"] - ref["mu"]) / ref["mu"]) <= TOL_MU
        assert _rel(r["z"], ref["z"]) <= TOL_Z and _rel(r["z"] - img, ref["z"] - img) <= TOL_DZ, seed
        if first is None:
            first = r["z"].copy()
    assert np.array_equal(r["z"], first)


def test_large_sample_count(ctx):
    """The reference's default is p = 1 % of the pixels (hpc/image_processing.c:187), thousands of samples on its larger
    inputs: p > 3072 runs the Jacobi pairs in two shared-memory chunks, m_pad > 4096 the wide-Phi filter paths."""
    from oracle import cpu_pipeline as cp
    W, H, p_req = 350, 300, 4000
    img = o.synthetic_image(W, H, 1, seed=17)
    s = oc.uniform_sampling(W, H, p_req)
    assert len(s) > 3072
    prm = gl.default_params(sample_size=p_req)
    r = ctx.run(img, prm)
    assert np.array_equal(ctx.get_samples(), s) and r["m"] == len(s) - 1
    ref = cp.run(img, s)
    err_mu = float(np.max(np.abs(r["mu"] - ref["mu"]) / ref["mu"]))
    err_z, err_dz = _rel(r["z"], ref["z"]), _rel(r["z"] - img, ref["z"] - img)
    print(f"large p: {W}x{H} p={len(s)}: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_mu <= TOL_MU and err_z <= TOL_Z and err_dz <= TOL_DZ
    ctx.set_option("fuse_filter", 0)           # Nystroem, then the filter on the stored (wide) Phi
    try:
        r2 = ctx.run(img, prm)
    finally:
        ctx.set_option("fuse_filter", 1)
    assert _rel(r2["z"], ref["z"]) <= TOL_Z and _rel(r2["z"] - img, ref["z"] - img) <= TOL_DZ


def test_projection_and_apply_variants_agree(ctx, golden):
    """c = Phi^T y from the affinity sums (default) vs the stand-alone pass over Phi; warp-per-row vs generic apply."""
    for tag in ("barbara_uniform256", "lion_rgb_photometric500", "test_uniform100"):
        g = golden(tag)
        src, a = _run_case(ctx, g)
        z = g["z"].astype(np.float64)
        res = {}
        for proj, app in (("recompute", "warp"), ("sums", "generic"), ("recompute", "generic")):
            ctx.set_option("projection", proj)
            ctx.set_option("filter_apply", app)
            try:
                _, b = _run_case(ctx, g)
            finally:
                ctx.set_option("projection", "sums")
                ctx.set_option("filter_apply", "warp")
            res[(proj, app)] = (_rel(b["z"], z), _rel(b["z"] - src, z - src))
            assert res[(proj, app)][0] <= TOL_Z
            assert _rel(b["z"], a["z"].astype(np.float64)) < 2e-4
        print(f"variants {tag}: default dz={_rel(a['z'] - src, z - src):.2e} " + " ".join(f"{k}:dz={v[1]:.2e}" for k, v in res.items()))


@pytest.mark.parametrize("tag", ["barbara_uniform256", "lion_rgb_photometric500", "cat_small_uniform_m20"])
def test_fused_filter_matches_staged(ctx, golden, tag):
    """gl_run applies the filter inside the extrapolation GEMM's epilogue (gl_nystroem_filter); option fuse_filter=0
    runs Nystroem and the filter apart like the reference.  Same answer, and Phi is still produced."""
    g = golden(tag)
    src, a = _run_case(ctx, g)                         # fused (default)
    ctx.set_option("fuse_filter", 0)
    try:
        _, b = _run_case(ctx, g)
    finally:
        ctx.set_option("fuse_filter", 1)
    z = g["z"].astype(np.float64)
    for r in (a, b):
        assert _rel(r["z"], z) <= TOL_Z and _rel(r["z"] - src, z - src) <= TOL_DZ
    assert _rel(a["z"], b["z"].astype(np.float64)) < 2e-5
    # stage-level call: Phi comes back and equals the staged Phi
    img, s = src, g["sample_indices"]
    ctx.set_image(img)
    ctx.set_samples(s)
    K_A, K_B = ctx.affinity(str(g["kind"]), float(g["h_loc"]), float(g["h_val"]))
    L_A, L_B = ctx.laplacian(K_A, K_B)
    U, mu, mu_inv = ctx.eigensolve(L_A, int(g["m"]))
    phi_f, z_f = ctx.nystroem_filter(L_B, U, mu_inv, mu)
    none, z_n = ctx.nystroem_filter(L_B, U, mu_inv, mu, keep_phi=False)     # Phi never written (patch layout where it applies): same z
    assert none is None and _rel(z_n, z_f.astype(np.float64)) < 2e-5
    # the two reference calls, Nystroem then ComputeResultFromLaplacian: Phi is deferred and the filter call runs both as one pass
    phi_l = ctx.nystroem(L_B, U, mu_inv)
    z_l = ctx.filter(phi_l, mu)
    assert np.array_equal(z_l, z_n)
    assert np.array_equal(phi_l.download(), phi_f.download())
    # and really apart (the matrix computed by the Nystroem call, then read back by the filter)
    ctx.set_option("lazy_phi", 0)
    try:
        phi_s = ctx.nystroem(L_B, U, mu_inv)
        assert np.array_equal(phi_f.download(), phi_s.download())
        z_s = ctx.filter(phi_s, mu)
    finally:
        ctx.set_option("lazy_phi", 1)
    assert _rel(z_f, z_s.astype(np.float64)) < 2e-5
    assert _rel(z_f, z) <= TOL_Z


def test_kb_cutoff_blocks(ctx):
    """K_B's spatial cutoff (sample blocks whose entries fp16 flushes to zero are not stored, affinity.cu): the
    stored blocks shrink, the skipped entries really are < 2^-25 in the oracle, and the result equals the dense run."""
    W, H, p_req, h_loc = 200, 1500, 700, 12.0
    img = o.synthetic_image(W, H, 1, seed=3)
    s = oc.uniform_sampling(W, H, p_req)
    out = {}
    for cut in (1, 0):
        ctx.set_option("kb_cutoff", cut)
        try:
            ctx.set_image(img)
            ctx.set_samples(s)
            K_A, K_B = ctx.affinity(gl.BILATERAL, h_loc, 30.0)
            info = K_B.info
            dense_pairs = -(-img.size // 512) * 512 * (-(-len(s) // 64) * 64)     # every 512-pixel tile x all (padded) sample slots
            D = K_B.rowsums()
            L_A, L_B = ctx.laplacian(K_A, K_B)
            U, mu, mu_inv = ctx.eigensolve(L_A, -1)
            phi = ctx.nystroem(L_B, U, mu_inv)
            z = ctx.filter(phi, mu)
            KB = K_B.download()
            out[cut] = dict(blocks=info.stored_pairs, dense=dense_pairs, layout=info.layout, D=D, z=z.astype(np.float64), KB=KB,
                            mu=mu.download())
        finally:
            ctx.set_option("kb_cutoff", 1)
    assert out[0]["blocks"] == out[0]["dense"] and out[0]["layout"] == 0     # dense: the blocked layout, every pair stored
    assert out[1]["layout"] == 1                                              # cutoff: the patch layout
    assert out[1]["blocks"] < 0.25 * out[1]["dense"], (out[1]["blocks"], out[1]["dense"])
    # every entry the cutoff dropped is below fp16's flush-to-zero threshold in the fp64 oracle
    cols = np.arange(0, img.size, 97)
    ref = o.affinity_rows(img, s, cols, "bilateral", h_loc, 30.0).T
    dropped = (out[1]["KB"][cols] == 0) & (out[0]["KB"][cols] != 0)
    assert not dropped.any() or ref[dropped].max() < 2.0 ** -24
    assert np.max(np.abs(out[1]["KB"][cols] - ref)) < 6e-4
    assert np.max(np.abs(out[1]["D"] - out[0]["D"]) / out[0]["D"]) < 1e-6
    assert np.max(np.abs(out[1]["mu"] - out[0]["mu"]) / out[0]["mu"]) < 1e-6
    assert _rel(out[1]["z"], out[0]["z"]) < 1e-6
    refp = oc.run_pipeline(img, s, h_loc=h_loc)
    assert np.max(np.abs(out[1]["mu"] - refp["mu"]) / refp["mu"]) <= TOL_MU
    assert _rel(out[1]["z"], refp["z"]) <= TOL_Z and _rel(out[1]["z"] - img, refp["z"] - img) <= TOL_DZ


@pytest.mark.parametrize("W,H,ch,p,h_loc", [(640, 480, 1, 300, 12.0), (301, 203, 3, 130, 40.0), (1920, 1080, 1, 1000, 40.0)])
def test_kb_block_32_matches_64(ctx, W, H, ch, p, h_loc):
    """Option kb_block=32: K_B stored in 32-slot blocks (fewer padded slots; the GEMM runs with 32-deep K steps and
    64-byte swizzled tiles).  K_B, eigenvalues and the filtered image must equal the 64-slot run; Phi through the
    tcgen05 GEMM must equal the CUDA-core checker."""
    img = o.synthetic_image(W, H, ch, seed=11)
    s = oc.random_sampling(W, H, p, 5)
    small = W * H * p < 1e8            # matrices are downloaded as fp64: only where that is a few hundred MB
    cols = np.arange(0, W * H, 53)
    out = {}
    for blk in (64, 32):
        ctx.set_option("kb_block", blk)
        ctx.set_option("kb_layout", "blocked")       # (the block size is a knob of the blocked layout)
        try:
            ctx.set_image(img)
            ctx.set_samples(s)
            K_A, K_B = ctx.affinity(gl.BILATERAL, h_loc, 30.0)
            assert K_B.info.ld == blk
            L_A, L_B = ctx.laplacian(K_A, K_B)
            U, mu, mu_inv = ctx.eigensolve(L_A, -1)
            r = dict(D=K_B.rowsums(), mu=mu.download(), blocks=K_B.info.stored_blocks)
            if small:
                r["KB"] = K_B.download()[cols]
                ctx.set_option("lazy_phi", 0)
                r["phi"] = ctx.nystroem(L_B, U, mu_inv).download()
                ctx.set_option("gemm", "simple")
                r["chk"] = ctx.nystroem(L_B, U, mu_inv).download()
                ctx.set_option("gemm", "tcgen05")
                ctx.set_option("lazy_phi", 1)
            r["z"] = ctx.filter(ctx.nystroem(L_B, U, mu_inv), mu).astype(np.float64)     # deferred Phi: the fused pass
            out[blk] = r
        finally:
            ctx.set_option("gemm", "tcgen05")
            ctx.set_option("lazy_phi", 1)
            ctx.set_option("kb_block", 64)
            ctx.set_option("kb_layout", "patch")
    a, b = out[64], out[32]
    print(f"kb_block: stored slots {a['blocks'] * 64} (64) vs {b['blocks'] * 32} (32)")
    assert np.max(np.abs(a["D"] - b["D"]) / a["D"]) < 1e-6
    assert np.max(np.abs(a["mu"] - b["mu"]) / a["mu"]) < 1e-6
    if small:
        assert np.array_equal(a["KB"], b["KB"])
        assert _rel(b["phi"], b["chk"]) < 2e-3 and _rel(a["phi"], a["chk"]) < 2e-3
        assert _rel(b["phi"], a["phi"]) < 2e-3
    y = img.astype(np.float64).reshape(a["z"].shape)
    assert _rel(b["z"], a["z"]) < 1e-5
    assert _rel(b["z"] - y, a["z"] - y) < 2e-3


@pytest.mark.parametrize("W,H,ch,kind,h_loc,h_val", [
    (96, 64, 1, "bilateral", 6.0, 30.0),        # window (r_c = 28) smaller than the image: the cutoff is exercised
    (61, 47, 1, "bilateral", 40.0, 30.0),       # reference constants: window covers the whole image
    (50, 40, 3, "bilateral", 5.0, 25.0),
    (64, 48, 1, "photometric", 40.0, 10.0),     # no spatial term: O(n^2)
    (40, 33, 1, "spatial", 4.0, 30.0),
])
def test_full_no_approx_path(ctx, W, H, ch, kind, h_loc, h_val):
    """-no_approx (hpc/affinity.c:264-336, laplacian.c:44-65, display.c:128-149), matrix-free on the device, against
    the dense fp64 oracle."""
    img = o.synthetic_image(W, H, ch, seed=W * H)
    ctx.set_image(img)
    K = ctx.full_affinity(kind, h_loc, h_val)
    L = ctx.full_laplacian(K)
    alpha = L.info.scale
    z, z8 = ctx.full_result(L, want_u8=True)
    ref = o.run_full(img, kind, h_loc, h_val)
    assert abs(alpha - ref["alpha"]) <= 2e-6 * ref["alpha"]
    err_z, err_dz = _rel(z, ref["z"]), _rel(z - img, ref["z"] - img)
    print(f"full {W}x{H}x{ch} {kind}: err_alpha={abs(alpha - ref['alpha']) / ref['alpha']:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_z <= TOL_Z and err_dz <= TOL_DZ
    assert z.min() >= 0.0 and z.max() <= 255.0
    assert np.array_equal(z8, z.astype(np.uint8))


def test_filter_options(ctx, golden):
    g = golden("test_uniform100")
    img, s = g["image"], g["sample_indices"]
    for gain, power in ((-1.0, 1.0), (3.0, 6.0), (0.0, 1.0)):
        prm = gl.default_params(sample_size=100, gain=gain, power=power)
        r = ctx.run(img, prm)
        ref = o.run_pipeline(img, s, gain=gain, power=power)
        assert _rel(r["z"], ref["z"]) <= TOL_Z
        if gain == 0.0:
            assert np.array_equal(r["z"], img.astype(np.float32))
    # u8 output: clamp to [0,255] then truncate
    prm = gl.default_params(sample_size=100)
    z8 = np.zeros(img.shape, dtype=np.uint8)
    r = ctx.run(img, prm, z8_out=z8)
    assert np.array_equal(z8, o.quantise(r["z"]))


def test_python_interface_and_colour_branch(ctx):
    """image_processing(y, cr, cb, **kwargs) as the prototype exposes it (python/image_processing.py:244) and the colour
    branch of its main (:411-424): luma filtered, chroma untouched."""
    rgb = o.synthetic_image(160, 120, 3, seed=21)
    ycc = gl.rgb2ycc(rgb.astype(np.float64))
    assert np.allclose(gl.ycc2rgb(ycc), rgb, atol=1e-9)
    y8 = np.clip(np.rint(ycc[:, :, 0]), 0, 255).astype(np.uint8)
    ycr, ycb = ycc[:, :, 1], ycc[:, :, 2]
    z, cr, cb = gl.image_processing(y8, ycr, ycb, ctx=ctx, sampling="spatially_uniform", affinity="bilateral", sample_size=80)
    assert cr is ycr and cb is ycb                                                                  # passed through untouched
    ref = oc.run_pipeline(y8, oc.uniform_sampling(160, 120, 80))
    assert _rel(z, ref["z"]) <= TOL_Z and _rel(z - y8, ref["z"] - y8) <= TOL_DZ
    out = gl.image_processing_rgb(rgb, ctx=ctx, sample_size=80)
    back = gl.rgb2ycc(out)
    assert np.allclose(back[:, :, 1:], ycc[:, :, 1:], atol=1e-9)                                    # chroma unchanged
    assert _rel(back[:, :, 0] - ycc[:, :, 0], ref["z"] - y8) <= TOL_DZ                               # luma changed by the filter


def test_errors(ctx):
    with pytest.raises(gl.GLError):
        ctx.set_image(np.zeros((1, 1), dtype=np.uint8))
    ctx.set_synthetic_image(64, 64, 1, 1)
    with pytest.raises(gl.GLError):
        ctx.set_samples(np.array([5, 3, 9], dtype=np.uint32))      # not ascending
    with pytest.raises(gl.GLError):
        ctx.set_samples(np.array([5, 4096], dtype=np.uint32))      # out of range
    with pytest.raises(gl.GLError):
        ctx.sampling(gl.RANDOM, 64 * 64 + 1)


def test_eigenvalue_buffer_is_checked(ctx):
    """Uniform sampling returns up to ~4x the requested count (hpc/sampling.c:8-13): gl_run must not write more eigenvalues
    than the caller's buffer holds (100x100 with 1124 requested gives p = 2401)."""
    import ctypes as C
    img = o.synthetic_image(100, 100, 1, seed=3)
    prm = gl.default_params(sample_size=1124)
    r = ctx.run(img, prm)
    assert r["p"] == 2401 and r["m"] == 2400 and len(r["mu"]) == 2400 and np.all(np.diff(r["mu"]) >= 0)
    small = np.zeros(100, dtype=np.float64)
    guard = small.copy()
    p, m = C.c_uint(), C.c_int()
    rc = gl.lib().gl_run(ctx.h, img.ctypes.data, 100, 100, 1, C.byref(prm), None, None, C.byref(p), C.byref(m), small.ctypes.data, 100)
    assert rc == gl.ERR_ARG and np.array_equal(small, guard)


def test_indefinite_matrix_eigenvalues_ascending(ctx):
    """gl_eigensolve on an uploaded symmetric indefinite matrix: negative eigenvalues sort below the positive ones."""
    rng = np.random.default_rng(5)
    q, _ = np.linalg.qr(rng.standard_normal((96, 96)))
    lam = np.concatenate([-np.linspace(0.5, 3.0, 40), np.linspace(0.2, 5.0, 56)])
    A = (q * lam) @ q.T
    A = 0.5 * (A + A.T)
    U, mu, _ = ctx.eigensolve(ctx.upload(gl.MAT_KA, A), 96)
    got = mu.download()
    assert np.all(np.diff(got) >= 0)
    assert np.max(np.abs(got - np.sort(lam))) < 1e-5


def test_deferred_phi_rejects_changed_samples(ctx, golden):
    """A deferred Phi is computed by its first consumer from the samples of the context: resampling in between must fail
    loudly instead of mixing the old eigenvectors with new sample indices."""
    g = golden("cat_small_random50")
    ctx.set_image(g["image"])
    ctx.set_samples(g["sample_indices"])
    K_A, K_B = ctx.affinity()
    L_A, L_B = ctx.laplacian(K_A, K_B)
    U, mu, mu_inv = ctx.eigensolve(L_A, -1)
    phi = ctx.nystroem(L_B, U, mu_inv)
    ctx.sampling(gl.RANDOM, len(g["sample_indices"]), seed=11)
    with pytest.raises(gl.GLError):
        ctx.filter(phi, mu)


@pytest.mark.parametrize("m,opti_gs,eps", [(8, 1, 1e-7), (20, 1, 0.1), (12, 2, 1e-5), (49, 1, 1e-6)])
def test_inverse_iteration_matches_the_restatement(ctx, golden, m, opti_gs, eps):
    """gl_inverse_iteration (the reference's InversePowerIteration, hpc/inverse_power_it.c:86-252, on the device) against
    oracle_np.inverse_power_iteration from the same start: same number of outer iterations, same eigenvalues 1/norm, same subspace;
    with a tight epsilon the eigenvalues are the converged smallest ones."""
    g = golden("cat_small_random50")
    ctx.set_image(g["image"])
    ctx.set_samples(g["sample_indices"])
    K_A, K_B = ctx.affinity()
    L_A, L_B = ctx.laplacian(K_A, K_B)
    A = L_A.download()
    U, mu, mu_inv, it, res = ctx.inverse_iteration(L_A, m, opti_gs, eps)
    rmu, rV, rit, rres = o.inverse_power_iteration(A, m, opti_gs, eps)
    assert it == rit and res <= eps
    got_mu, got_U = mu.download(), U.download()
    assert np.max(np.abs(got_mu - rmu) / rmu) < 1e-6
    assert np.max(np.abs(mu_inv.download() * got_mu - 1.0)) < 1e-12
    assert np.max(np.abs(np.abs(got_U) - np.abs(rV))) < 1e-5                        # same iterates (fp32 storage), column by column
    if eps <= 1e-6:     # (the rule stops on the subspace: close eigenvalues inside it are only roughly separated, as in the reference)
        lam = np.linalg.eigvalsh(A)[:m]
        assert np.max(np.abs(np.sort(got_mu) - lam) / lam) < 5e-2
    # the pairs feed the rest of the path like those of gl_eigensolve
    z = ctx.filter(ctx.nystroem(L_B, U, mu_inv), mu)
    assert np.isfinite(z).all()


def test_repeatability_and_launch_count(ctx, golden):
    g = golden("cat_small_random50")
    n0 = ctx.kernel_launches()
    _, a = _run_case(ctx, g)
    n1 = ctx.kernel_launches()
    _, b = _run_case(ctx, g)
    assert n1 - n0 >= 15
    assert np.array_equal(a["z"], b["z"]) and np.array_equal(a["mu"], b["mu"])   # deterministic reductions
    t = ctx.stage_ms()
    assert t["affinity"] > 0 and t["eigen"] > 0 and t["nystroem"] > 0 and t["filter"] > 0
