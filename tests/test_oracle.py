"""CPU tests: the oracle (numpy + C) against the golden fixtures that were
produced by the reference's own Python modules (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as o

CASES = ["test_uniform100", "cat_small_random50", "barbara_uniform256", "lion_rgb_photometric500",
         "lion_photometric_h10", "test_spatial_h10", "cat_small_uniform_m20"]


def _src(g):
    img = g["image"]
    return np.repeat(img[:, :, None], 3, axis=2) if int(g["rgb"]) else img


def test_sampling_matches_reference_modules(golden):
    tab = golden("sampling")
    assert len(tab) == 28
    for key, ref in tab.items():
        parts = key.split("_")
        W, H = (int(v) for v in parts[1].split("x"))
        p = int(parts[2])
        if parts[0] == "uniform":
            got_np, got_c = o.uniform_sampling(W, H, p), oc.uniform_sampling(W, H, p)
        else:
            seed = int(parts[3][1:])
            got_c = oc.random_sampling(W, H, p, seed)
            # the pure-python generator is slow: spot-check it on the small shapes only
            got_np = o.random_sampling(W, H, p, seed) if p <= 500 else got_c
            assert len(ref) == p
        assert got_c.dtype == np.uint32
        assert np.array_equal(got_c, ref), key           # bit-exact
        assert np.array_equal(got_np, ref), key
        assert np.all(np.diff(ref.astype(np.int64)) > 0)  # strictly ascending


@pytest.mark.parametrize("tag", ["test_uniform100", "cat_small_random50", "lion_photometric_h10", "test_spatial_h10"])
def test_affinity_matches_reference_modules(golden, tag):
    g = golden(tag)
    img, s = g["image"], g["sample_indices"]
    kind, h_loc, h_val = str(g["kind"]), float(g["h_loc"]), float(g["h_val"])
    K_A = o.affinity_rows(img, s, s, kind, h_loc, h_val)
    assert np.allclose(K_A, g["ref_K_A"], rtol=1e-13, atol=1e-300)
    assert np.allclose(np.diag(K_A), 1.0) and np.allclose(K_A, K_A.T)
    n = img.size
    D = o.affinity_rows(img, s, np.arange(n), kind, h_loc, h_val).sum(axis=1)
    assert np.allclose(D, g["ref_D"], rtol=1e-12)
    assert np.allclose(g["D"], g["ref_D"], rtol=1e-12)     # the pipeline's D_A = rowsum(K_A)+rowsum(K_B)


def test_barbara_rowsums_match_reference_module(golden):
    g = golden("barbara_uniform256")
    assert np.allclose(g["D"], g["ref_D"], rtol=1e-12)
    assert len(g["sample_indices"]) == 256


@pytest.mark.parametrize("tag", CASES)
def test_c_oracle_matches_golden(golden, tag):
    g = golden(tag)
    src = _src(g)
    r = oc.run_pipeline(src, g["sample_indices"], m=int(g["m"]), kind=str(g["kind"]),
                        h_loc=float(g["h_loc"]), h_val=float(g["h_val"]))
    assert np.allclose(r["D"], g["D"], rtol=1e-11)
    assert abs(r["alpha"] - float(g["alpha"])) <= 1e-12 * float(g["alpha"])
    assert np.allclose(r["mu"], g["mu"], rtol=1e-9)
    z = g["z"].astype(np.float64)
    assert np.linalg.norm(r["z"] - z) <= 2e-6 * np.linalg.norm(z)      # fixture is stored as float32
    dz = np.linalg.norm(r["z"] - src)
    assert np.linalg.norm((r["z"] - src) - (z - src)) <= 1e-4 * dz


def test_numpy_oracle_invariants(golden):
    g = golden("cat_small_random50")
    r = o.run_pipeline(g["image"], g["sample_indices"], return_phi=True)
    L_A, mu, phi = r["L_A"], r["mu"], r["phi"]
    assert np.all(r["D"] > 0)
    assert np.all(np.abs(np.diag(L_A)) > np.sum(np.abs(L_A), axis=1) - np.abs(np.diag(L_A)))  # strictly diag. dominant
    assert np.all(mu > 0) and np.all(np.diff(mu) >= 0)
    G = phi.T @ phi
    assert np.linalg.norm(G - np.eye(G.shape[0])) < 2e-2      # nearly orthonormal before any GS
    assert np.array_equal(phi[g["sample_indices"].astype(np.int64)], r["U"])
    assert np.allclose(r["mu"], g["mu"], rtol=1e-10)
    assert np.linalg.norm(r["z"] - g["z"]) <= 2e-6 * np.linalg.norm(g["z"])


def test_gram_schmidt_oracles_agree(golden):
    g = golden("test_uniform100")
    rn = o.run_pipeline(g["image"], g["sample_indices"], orthonormalise=True)
    rc = oc.run_pipeline(g["image"], g["sample_indices"], orthonormalise=True)
    assert np.linalg.norm(rn["z"] - rc["z"]) <= 1e-9 * np.linalg.norm(rn["z"])
    assert np.linalg.norm(rn["z"] - g["z_gs"]) <= 2e-6 * np.linalg.norm(rn["z"])
    X = np.random.RandomState(0).randn(50, 7)
    Q, norms = o.gram_schmidt(X)
    assert np.allclose(Q.T @ Q, np.eye(7), atol=1e-12) and np.all(norms > 0)


def test_symeig_against_lapack():
    A = np.random.RandomState(1).randn(120, 120)
    A = A + A.T
    d, V = oc.symeig(A)
    assert np.allclose(d, np.linalg.eigvalsh(A), atol=1e-11)
    assert np.allclose(A @ V, V * d, atol=1e-10)


def test_synthetic_generators_agree():
    for ch in (1, 3):
        a, b = o.synthetic_image(257, 131, ch), oc.synthetic_image(257, 131, ch)
        assert a.dtype == np.uint8 and np.array_equal(a, b)
    img = o.synthetic_image(640, 360)
    assert 60 < img.mean() < 200 and img.std() > 20


def test_band_sample_is_a_restriction(golden):
    """rows=(r0,r1) of the C oracle (bench's bounded CPU sample) only touches band pixels."""
    g = golden("test_uniform100")
    img = g["image"]
    r = oc.run_pipeline(img, g["sample_indices"], rows=(20, 60))
    z = r["z"]
    assert np.array_equal(z[:20], img[:20].astype(np.float64)) and np.array_equal(z[60:], img[60:].astype(np.float64))
    assert not np.array_equal(z[20:60], img[20:60].astype(np.float64))


def test_full_path_oracle_identities():
    """run_full (the -no_approx restatement): rows of L = alpha (D - K) sum to zero, so a constant image is a fixed point;
    and the streamed form y - alpha (D y - K y) equals the explicit y - L y."""
    from oracle import oracle_np as o
    img = o.synthetic_image(24, 18, 1, seed=2)
    r = o.run_full(img, "bilateral", 6.0, 30.0)
    n = img.size
    K = o.affinity_rows(img, np.arange(n), np.arange(n), "bilateral", 6.0, 30.0)
    D = K.sum(axis=1)
    alpha = 1.0 / D.mean()
    L = alpha * (np.diag(D) - K)
    assert np.allclose(L.sum(axis=1), 0.0, atol=1e-12)
    y = img.reshape(-1).astype(np.float64)
    assert np.allclose(np.clip(y - L @ y, 0, 255).reshape(img.shape), r["z"], rtol=0, atol=1e-9)
    assert abs(r["alpha"] - alpha) < 1e-15
    flat = np.full((10, 12), 77, dtype=np.uint8)
    assert np.allclose(o.run_full(flat)["z"], 77.0)


PYREF = ["pyref_test100", "pyref_lion_crop", "pyref_lion_crop_photometric", "pyref_test100_spatial", "pyref_lion_crop_random7"]
# bandwidths hard-coded in the reference's plugins: bilateral 30 / 40 (bilateral.py:12-13), photometric 10, spatial 10
PYREF_H = {"bilateral": (40.0, 30.0), "photometric": (40.0, 10.0), "spatial": (10.0, 30.0)}


@pytest.mark.parametrize("tag", PYREF)
def test_oracle_reproduces_the_reference_python_pipeline(golden, tag):
    """tests/golden/pyref_*.npz hold the output of the reference's OWN image_processing(y, **kwargs)
    (python/image_processing.py:244-357, run by tests/golden/make_golden_pyref.py) for its three affinity plugins and both
    samplers.  With the prototype's constants -- all p eigenpairs, f(mu) = mu + 5, gain -1, no clipping -- the oracle must
    give the same image: this pins its Laplacian normalisation, eigenpair ordering, Nystroem extrapolation, permutation
    back to raster order and filter algebra to reference code."""
    from oracle import oracle_np as o
    g = golden(tag)
    img, s, kind, seed = g["image"], g["sample_indices"], str(g["kind"]), int(g["seed"])
    H, W = img.shape
    mine = o.uniform_sampling(W, H, int(W * H * 0.01)) if seed < 0 else o.random_sampling(W, H, int(W * H * 0.01), seed)
    assert np.array_equal(mine, s)
    h_loc, h_val = PYREF_H[kind]
    for streamed in (False, True):                      # both code paths of the oracle (Phi held / streamed)
        r = o.run_pipeline(img, s, kind=kind, h_loc=h_loc, h_val=h_val, all_pairs=True, f_of_mu=lambda mu: mu + 5.0, gain=-1.0,
                           clip=False, return_phi=not streamed)
        assert r["m"] == len(s)
        err = np.linalg.norm(r["z"] - g["z"]) / np.linalg.norm(g["z"])
        err_d = np.linalg.norm((r["z"] - img) - (g["z"] - img)) / np.linalg.norm(g["z"] - img)
        assert err < 1e-9 and err_d < 1e-8, (tag, streamed, err, err_d)


@pytest.mark.parametrize("name", ["sq24", "rect"])
def test_nlm_oracle_matches_the_reference_module(name):
    """python/affinity_methods/NLM.py run by tests/golden/make_golden_nlm.py.  The reference's rows are samples in raster
    order, its columns pixels in COLUMN-major order (im2col of the transposed image, NLM.py:21); the oracle uses raster
    order for both.  Through that index map the two agree to rounding; read as raster columns (what the reference's
    callers do, python/image_processing.py:60-64) they do not, not even on a square image."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"pyref_nlm_{name}.npz"))
    img, s, K = g["image"], g["sample_indices"], g["K_AB"]
    M, N = img.shape
    mine = o.nlm_affinity_rows(img, s, np.arange(M * N), 3.0)
    r, c = np.divmod(np.arange(M * N), N)
    assert np.max(np.abs(mine - K[:, c * M + r])) < 1e-13
    assert np.max(np.abs(mine - K)) > 0.1
    # properties of the intended matrix: symmetric sample block with a unit diagonal, values in (0, 1]
    K_A = mine[:, s]
    assert np.allclose(K_A, K_A.T, atol=1e-15) and np.allclose(np.diag(K_A), 1.0)
    assert mine.min() > 0 and mine.max() <= 1.0
    # and the whole oracle pipeline runs on it
    out = o.run_pipeline(img, s, kind=o.NLM, h_val=3.0)
    assert np.all(np.diff(out["mu"]) >= 0) and out["mu"][0] > 0 and np.isfinite(out["z"]).all()


def test_prototype_building_blocks_match_the_reference_functions():
    """oracle/proto_np.py against the reference's own affinity / nystroem / permutation / orthogonalisation / sinkhorn
    (python/image_processing.py:35-129, run by tests/golden/make_golden_proto.py).  Eigenvector signs are free, so the
    comparisons go through sign-invariant quantities."""
    from oracle import proto_np as pr
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pyref_proto.npz"))
    img, s = g["image"], g["sample_indices"]
    H, W = img.shape
    n, p = H * W, len(s)
    # affinity(): the bilateral plugin with its constants (h_spatial 40, h_photo 30), split into K_A | K_B
    K_AB = o.affinity_rows(img, s, np.arange(n), "bilateral", 40.0, 30.0)
    K_A, K_B = pr.split_affinity(K_AB, s)
    assert np.max(np.abs(K_A - g["K_A"])) < 1e-13 and np.max(np.abs(K_B - g["K_B"])) < 1e-13
    # nystroem(): same spectrum, same Phi diag(Pi) Phi^T
    phi, Pi = pr.nystroem(K_A, K_B)
    assert np.max(np.abs(Pi - g["Pi"]) / g["Pi"]) < 1e-10

    def recon(F, w):
        return (F * w) @ F.T
    assert np.max(np.abs(recon(phi, Pi) - recon(g["phi"], g["Pi"]))) < 1e-8
    # permutation(): row movement only -- apply it to the reference's own phi
    assert np.array_equal(pr.permutation(g["phi"], s), g["phi_perm"])
    # orthogonalisation(): orthonormal columns, capped eigenvalues, same projector-weighted matrix
    V, Pi_V = pr.orthogonalisation(K_A, K_B)
    assert np.max(np.abs(V.T @ V - np.eye(p))) < 1e-8
    assert np.max(np.abs(Pi_V - g["Pi_V"])) < 1e-10
    assert np.max(np.abs(np.abs(V) - np.abs(g["V"]))) < 1e-6 or np.max(np.abs(recon(V, Pi_V) - recon(g["V"], g["Pi_V"]))) < 1e-8
    # sinkhorn(): sign-free by construction
    W_A, W_B = pr.sinkhorn(g["phi"], g["Pi"])
    assert np.max(np.abs(W_A - g["W_A"])) < 1e-9 and np.max(np.abs(W_B - g["W_B"])) < 1e-9
    W_A2, _ = pr.sinkhorn(phi, Pi)                        # and from the oracle's own eigenvectors
    assert np.max(np.abs(W_A2 - g["W_A"])) < 1e-6
    # smoothing_matrix() / smoothing() / sharpening() (:151-241) from the reference's own (phi, Pi)
    V_s, L_s = pr.smoothing_matrix(s, g["phi"], g["Pi"])
    assert np.max(np.abs(L_s - g["L_s"])) < 1e-12 and np.max(np.abs(np.abs(V_s) - np.abs(g["V_s"]))) < 1e-9
    y = g["image"].astype(np.float64)
    assert np.max(np.abs(pr.smoothing(y, s, g["phi"], g["Pi"]) - g["z_smooth"])) < 1e-9
    assert np.max(np.abs(pr.sharpening(y, s, g["phi"], g["Pi"]) - g["z_sharp"])) < 1e-9


def test_inverse_power_iteration_restatement():
    """oracle_np.inverse_power_iteration (hpc/inverse_power_it.c:86-252 with exact solves): with a tight epsilon it returns the m
    smallest eigenpairs of an SPD matrix in ascending order (pinned to numpy's eigh); with the reference's default epsilon = 0.1
    it stops early with the residual it promises; opti_gs > 1 orthonormalises less often and ends with one more pass (:183-186)."""
    from oracle import oracle_np as o
    rng = np.random.default_rng(3)
    q, _ = np.linalg.qr(rng.standard_normal((60, 60)))
    lam = np.sort(rng.uniform(0.2, 2.0, 60))
    A = (q * lam) @ q.T
    A = 0.5 * (A + A.T)
    mu, V, it, r = o.inverse_power_iteration(A, 6, 1, 1e-10)
    assert r <= 1e-10 and it > 3
    # the stopping rule measures the invariance of the SUBSPACE: the subspace is converged, well separated eigenvalues come out
    # exactly as 1/norm, close ones (0.3032 / 0.3047 here) only as well as the single vectors have separated by then
    assert np.max(np.abs(mu[:3] - lam[:3]) / lam[:3]) < 1e-9
    assert np.max(np.abs(mu - lam[:6]) / lam[:6]) < 1e-2
    Qv, _ = np.linalg.qr(V)              # (the returned vectors are the normalised iterates BEFORE orthonormalisation, :230-235)
    assert np.linalg.norm(Qv @ Qv.T - q[:, :6] @ q[:, :6].T) < 1e-6
    mu2, V2, it2, r2 = o.inverse_power_iteration(A, 6, 1, 0.1)
    assert r2 <= 0.1 and it2 < it
    # opti_gs = 3: the norms the last orthonormalisation hands out belong to k un-normalised solves (k = 3, or it % 3 for the
    # closing pass), so the reference's 1 / norm is lambda^k -- a quirk of :172-175,204 that the restatement keeps
    mu3, V3, it3, r3 = o.inverse_power_iteration(A, 6, 3, 1e-8)
    k = it3 % 3 or 3
    assert r3 <= 1e-8 and np.max(np.abs(mu3 - lam[:6] ** k) / lam[:6] ** k) < 3e-2
    X0 = o.inverse_iteration_start(5, 3)
    assert X0.shape == (5, 3) and X0.min() > 0 and X0.max() < 1 and len(np.unique(X0)) == 15
