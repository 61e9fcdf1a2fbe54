"""The experimental blocks of the reference's Python prototype on the device (csrc/proto.cu) through the C ABI:
gl_sinkhorn, gl_orthogonalisation, gl_smoothing_matrix, gl_matrix_filter (smoothing / sharpening), and gl_nystroem on a bare K_B
(the prototype's nystroem(K_A, K_B)).

Checked (a) against tests/golden/pyref_proto.npz, the outputs of the reference's OWN functions (python/image_processing.py:69-241,
tests/golden/make_golden_proto.py) on a 24 x 30 crop of input/lion.png with 12 uniform samples, and (b) against the pinned
restatements oracle/proto_np.py on a larger synthetic image with better separated samples.

Tolerances.  Phi is stored in fp16 (relative rounding 4.9e-4) and the fixture's K_A has condition number 3.5e4, so quantities
that pass through Phi are compared at a few 1e-3 of their scale; eigenvalues at 1e-3 absolute (they are O(0.1 .. 1));
orthogonalisation runs in fp64 up to the fp16 store of V.  Eigenvector signs are free: columns are compared as |.| or through
V diag(L) V^T."""
import numpy as np
import pytest

import ipgl_b200 as gl
from oracle import oracle_np as o
from oracle import proto_np as pr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = gl.Context(0)
    yield c
    c.close()


def _setup(ctx, g):
    ctx.set_image(g["image"])
    ctx.set_samples(g["sample_indices"])
    phi = ctx.upload(gl.MAT_PHI, g["phi_perm"])       # the reference's own Phi, rows in raster order
    Pi = ctx.upload(gl.MAT_DIAG, g["Pi"])
    return phi, Pi


def _relmax(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_phi_upload_roundtrip(ctx, golden):
    g = golden("pyref_proto")
    phi, _ = _setup(ctx, g)
    got = phi.download()
    assert got.shape == g["phi_perm"].shape
    assert np.max(np.abs(got - g["phi_perm"]) / (np.abs(g["phi_perm"]) + 1e-3)) < 1e-3      # fp16 storage


def test_sinkhorn_matches_the_reference_function(ctx, golden):
    g = golden("pyref_proto")
    s = g["sample_indices"].astype(np.int64)
    phi, Pi = _setup(ctx, g)
    W_A, W_ABt = ctx.sinkhorn(phi, Pi, 100)
    wa = W_A.download()
    wab = W_ABt.download().T                                   # [p, n], columns in raster order
    assert wa.shape == g["W_A"].shape and wab.shape == (len(s), g["image"].size)
    wb = np.delete(wab, s, axis=1)
    e_a, e_b = _relmax(wa, g["W_A"]), _relmax(wb, g["W_B"])
    rows = wab.sum(axis=1)                                      # diag(r) K diag(c) has unit row sums
    print(f"sinkhorn: err W_A={e_a:.2e} W_B={e_b:.2e} row sums in [{rows.min():.5f}, {rows.max():.5f}]")
    assert e_a < 5e-3 and e_b < 5e-3
    assert np.max(np.abs(wab[:, s] - wa)) < 2e-3 * np.max(np.abs(wa))          # the sample columns of W_AB are W_A
    assert np.max(np.abs(rows - 1.0)) < 5e-3
    W_A0, _ = ctx.sinkhorn(phi, Pi, 0)                          # no iterations: r = c = 1, W = K
    K_AA = (g["phi"][:len(s)] * g["Pi"]) @ g["phi"][:len(s)].T
    assert _relmax(W_A0.download(), K_AA) < 2e-3


def test_smoothing_matrix_and_filters_match_the_reference_functions(ctx, golden):
    g = golden("pyref_proto")
    phi, Pi = _setup(ctx, g)
    V, L = ctx.smoothing_matrix(phi, Pi)
    Lg, Vg = L.download(), V.download()
    assert Vg.shape == g["V_s"].shape
    e_L = float(np.max(np.abs(Lg - g["L_s"])))
    e_V = float(np.max(np.abs(np.abs(Vg) - np.abs(g["V_s"]))))
    e_W = _relmax((Vg * Lg) @ Vg.T, (g["V_s"] * g["L_s"]) @ g["V_s"].T)
    print(f"smoothing_matrix: err L={e_L:.2e} |V|={e_V:.2e} V L V^T={e_W:.2e}")
    # (columns of close eigenvalues -- -0.1005 / -0.1039 here -- rotate into each other under the fp16 rounding of Phi: e_V is only
    # printed; what every consumer uses is V L V^T)
    assert e_L < 1e-3 and e_W < 1e-2
    gap = np.min(np.abs(np.subtract.outer(g["L_s"], g["L_s"])) + np.eye(len(Lg)), axis=1)
    iso = gap > 0.01                                                           # well separated eigenvalues: their vectors must agree
    assert iso.sum() >= 4 and np.max(np.abs(np.abs(Vg[:, iso]) - np.abs(g["V_s"][:, iso]))) < 1e-2
    s = g["sample_indices"].astype(np.int64)
    assert np.max(np.abs(Vg[s].T @ Vg[s] - np.eye(len(s)))) < 2e-3             # sample rows: the orthonormal eigenvectors of W_A themselves
    z_sm = ctx.smoothing(phi, Pi).astype(np.float64)
    z_sh = ctx.sharpening(phi, Pi).astype(np.float64)
    e_sm = float(np.linalg.norm(z_sm - g["z_smooth"]) / np.linalg.norm(g["z_smooth"]))
    e_sh = float(np.linalg.norm(z_sh - g["z_sharp"]) / np.linalg.norm(g["z_sharp"]))
    print(f"smoothing err={e_sm:.2e} sharpening err={e_sh:.2e}")
    assert e_sm < 1e-2 and e_sh < 2e-2
    # the polynomial filter itself, from the reference's own (V, L): only the fp16 store of V in between
    Vh, Lh = ctx.upload(gl.MAT_PHI, g["V_s"]), ctx.upload(gl.MAT_DIAG, g["L_s"])
    z1 = ctx.matrix_filter(Vh, Lh, [0.0, 1.0]).astype(np.float64)
    z2 = ctx.matrix_filter(Vh, Lh, [0.0, 0.0, 2.5, -1.5]).astype(np.float64)
    assert np.linalg.norm(z1 - g["z_smooth"]) / np.linalg.norm(g["z_smooth"]) < 2e-3
    assert np.linalg.norm(z2 - g["z_sharp"]) / np.linalg.norm(g["z_sharp"]) < 5e-3
    y = g["image"].astype(np.float64)
    assert np.array_equal(ctx.matrix_filter(Vh, Lh, [1.0]).astype(np.float64), y)       # W^0 = identity


def test_orthogonalisation_matches_the_reference_function(ctx, golden):
    g = golden("pyref_proto")
    ctx.set_image(g["image"])
    ctx.set_samples(g["sample_indices"])
    K_A, K_B = ctx.affinity(gl.BILATERAL)
    V, Pi = ctx.orthogonalisation(K_A, K_B)
    Vg, Pg = V.download(), Pi.download()
    assert Vg.shape == g["V"].shape
    e_pi = float(np.max(np.abs(Pg - g["Pi_V"])))
    e_orth = float(np.max(np.abs(Vg.T @ Vg - np.eye(Vg.shape[1]))))
    rec = lambda V_, P_: (V_ * P_) @ V_.T
    Vref = pr.permutation(g["V"], g["sample_indices"])          # the prototype returns V with the sample rows first (:122)
    e_rec = _relmax(rec(Vg, Pg), rec(Vref, g["Pi_V"]))
    print(f"orthogonalisation: err Pi={e_pi:.2e} V^T V - I={e_orth:.2e} V Pi V^T={e_rec:.2e}")
    assert e_pi < 1e-3 and e_orth < 5e-3 and e_rec < 5e-3


def test_prototype_nystroem_on_a_bare_K_B(ctx, golden):
    """nystroem(K_A, K_B) (:69-88) on the device: eigenpairs of K_A (descending), Phi = [Phi_A; K_B^T Phi_A / Pi] in raster order.
    Compared through K = Phi Pi Phi^T (signs cancel; the columns of the small eigenvalues are weighted down as in every use)."""
    g = golden("pyref_proto")
    ctx.set_image(g["image"])
    ctx.set_samples(g["sample_indices"])
    K_A, K_B = ctx.affinity(gl.BILATERAL)
    ctx.set_option("eig_largest", 1)
    try:
        U, Pi, Pi_inv = ctx.eigensolve(K_A, len(g["sample_indices"]))
    finally:
        ctx.set_option("eig_largest", 0)
    assert np.max(np.abs(Pi.download() - g["Pi"]) / g["Pi"]) < 1e-4
    phi = ctx.nystroem(K_B, U, Pi_inv)
    P = phi.download()
    rec = (P * Pi.download()) @ P.T
    ref = (g["phi_perm"] * g["Pi"]) @ g["phi_perm"].T
    e = _relmax(rec, ref)
    print(f"nystroem(K_A, K_B): err Phi Pi Phi^T={e:.2e}")
    assert e < 1e-2


@pytest.mark.parametrize("W,H,ch,p_req,h_loc,h_val", [(96, 64, 1, 24, 14.0, 30.0), (120, 80, 3, 40, 20.0, 60.0)])
def test_blocks_against_the_restatements_on_a_larger_image(ctx, W, H, ch, p_req, h_loc, h_val):
    """Same blocks against oracle/proto_np.py (pinned to the reference functions by tests/test_oracle.py) on a synthetic image whose
    samples are a few bandwidths apart (K_A well conditioned) and whose every pixel is within reach of some sample (a pixel whose row
    of K is ~1e-10 makes Sinkhorn's 1 / (K c) ill-posed in any 16-bit Phi), m_pad = 64, colour included."""
    img = o.synthetic_image(W, H, ch, seed=5)
    s = o.random_sampling(W, H, p_req, 2)
    n, p = W * H, len(s)
    K = o.affinity_rows(img, s, np.arange(n), kind=o.BILATERAL, h_loc=h_loc, h_val=h_val)          # [p, n]
    K_A, K_B = pr.split_affinity(K, s)
    phi, Pi = pr.nystroem(K_A, K_B)
    # the device holds Phi in fp16: give the restatements the same rounded matrix, so that the comparison measures the kernels and not
    # how sensitive a block is to its input (the colour case's sharpening moves by 12 % under that rounding alone)
    phi = phi.astype(np.float16).astype(np.float64)
    phi_perm = pr.permutation(phi, s)
    ctx.set_image(img)
    ctx.set_samples(s)
    P, D = ctx.upload(gl.MAT_PHI, phi_perm), ctx.upload(gl.MAT_DIAG, Pi)
    # sinkhorn
    rW_A, rW_B = pr.sinkhorn(phi, Pi)
    W_A, W_ABt = ctx.sinkhorn(P, D, 100)
    wab = W_ABt.download().T
    e_a, e_b = _relmax(W_A.download(), rW_A), _relmax(np.delete(wab, s, axis=1), rW_B)
    # smoothing_matrix, smoothing, sharpening (per channel)
    rV, rL = pr.smoothing_matrix(s, phi, Pi)
    V, L = ctx.smoothing_matrix(P, D)
    e_L = float(np.max(np.abs(L.download() - rL)) / np.max(np.abs(rL)))
    e_W = _relmax((V.download() * L.download()) @ V.download().T, (rV * rL) @ rV.T)
    imgc = img.reshape(H, W, ch).astype(np.float64)
    z_sh = ctx.sharpening(P, D).astype(np.float64).reshape(H, W, ch)
    r_sh = np.stack([pr.sharpening(imgc[:, :, c], s, phi, Pi) for c in range(ch)], axis=2)
    e_sh = float(np.linalg.norm(z_sh - r_sh) / np.linalg.norm(r_sh))
    # orthogonalisation
    dK_A, dK_B = ctx.affinity(gl.BILATERAL, h_loc=h_loc, h_val=h_val)
    Vo, Po = ctx.orthogonalisation(dK_A, dK_B)
    rVo, rPo = pr.orthogonalisation(K_A, K_B)
    rVo = pr.permutation(rVo, s)
    e_pi = float(np.max(np.abs(Po.download() - rPo)))
    e_rec = _relmax((Vo.download() * Po.download()) @ Vo.download().T, (rVo * rPo) @ rVo.T)
    print(f"{W}x{H}x{ch} p={p}: sinkhorn {e_a:.1e}/{e_b:.1e}  smoothing L {e_L:.1e} W {e_W:.1e}  sharpening {e_sh:.1e}  "
          f"orthogonalisation Pi {e_pi:.1e} rec {e_rec:.1e}")
    assert e_a < 5e-3 and e_b < 5e-3 and e_L < 1e-3 and e_W < 5e-3 and e_sh < 5e-3 and e_pi < 1e-3 and e_rec < 5e-3
