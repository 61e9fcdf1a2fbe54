"""TEST INFRASTRUCTURE ONLY (like the rest of oracle/): numpy restatements of the experimental building blocks of the
reference's Python prototype, pinned to the reference's own functions through tests/golden/pyref_proto.npz
(tests/golden/make_golden_proto.py): the checker of the device versions in csrc/proto.cu (tests/test_proto_gpu.py).

Each function cites the lines of /root/reference/python/image_processing.py it follows.  Eigenvector columns are only
defined up to sign (and up to rotations inside repeated eigenvalues), so callers compare invariants: Phi diag(Pi) Phi^T,
|columns|, or the downstream matrices."""
import numpy as np


def split_affinity(K_AB, sample_indices):
    """affinity(): K_A = K_AB[:, samples], K_B = K_AB with the sample columns deleted, the rest in raster order (:59-64)."""
    s = np.asarray(sample_indices, dtype=np.int64)
    rest = np.delete(np.arange(K_AB.shape[1]), s)
    return K_AB[:, s], K_AB[:, rest]


def nystroem(K_A, K_B):
    """nystroem(): eigenpairs of the symmetric PSD sample block (the reference takes its SVD, :71), descending, and the
    extension Phi = [Phi_A; K_B^T Phi_A diag(1/Pi)] (:78-84)."""
    w, U = np.linalg.eigh(K_A)
    w, U = w[::-1], U[:, ::-1]
    return np.concatenate((U, K_B.T @ (U / w))), w


def permutation(phi, sample_indices):
    """permutation() (:35-50) and hpc/utils.c:134-173: row i < p goes to raster position s_i, the remaining rows fill the
    non-sample positions in ascending order."""
    s = np.asarray(sample_indices, dtype=np.int64)
    n, p = phi.shape[0], len(s)
    out = np.empty_like(phi)
    out[s] = phi[:p]
    mask = np.ones(n, dtype=bool)
    mask[s] = False
    out[mask] = phi[p:]
    return out


def orthogonalisation(A, B):
    """orthogonalisation() (:112-129), the one-shot orthogonal Nystroem extension: with A^-1/2 from the eigenpairs of A,
    Q = A + A^-1/2 B B^T A^-1/2 = Phi_Q Pi_Q Phi_Q^T, V = [A; B^T] A^-1/2 Phi_Q Pi_Q^-1/2 (orthonormal columns), and the
    returned eigenvalues are Pi_Q capped at 1 (:125-126)."""
    w, U = np.linalg.eigh(A)
    A_isqrt = (U / np.sqrt(w)) @ U.T
    Q = A + A_isqrt @ B @ B.T @ A_isqrt
    wq, Uq = np.linalg.eigh(Q)
    wq, Uq = wq[::-1], Uq[:, ::-1]
    V = np.concatenate((A, B.T)) @ A_isqrt @ Uq @ np.diag(1.0 / np.sqrt(wq))
    return V, np.minimum(wq, 1.0)


def sinkhorn(phi, Pi, iterations=100):
    """sinkhorn() (:92-109): alternate row/column scalings r, c of K = Phi diag(Pi) Phi^T (never formed: two products with
    Phi per step), starting from r = 1; then the first p rows of diag(r) K diag(c), split into the sample block W_A and
    the rest W_B.  nan_to_num as in the reference (:99,101)."""
    M, N = phi.shape
    r = np.ones(M)
    c = r
    for _ in range(iterations):
        c = np.nan_to_num(1.0 / (phi @ (Pi * (phi.T @ r))))
        r = np.nan_to_num(1.0 / (phi @ (Pi * (phi.T @ c))))
    W_AB = (r[:N, None] * phi[:N] * Pi) @ (phi * c[:, None]).T
    return W_AB[:, :N], W_AB[:, N:]


def smoothing_matrix(sample_indices, phi, Pi):
    """smoothing_matrix() (:151-194).  phi has the sample rows first (nystroem()'s order).  K = Phi diag(Pi) Phi^T is only needed
    through its degrees D = K 1 = Phi (Pi o (Phi^T 1)), alpha = 1 / mean(D), and the first p rows of W = I + alpha (K - diag D):
    W_A = I + alpha (Phi_A Pi Phi_A^T - diag D_A), W_B = alpha Phi_A Pi Phi_B^T.  Then the eigenpairs (L, Phi_WA) of W_A, descending
    (:183-185), their extension V = [Phi_WA; W_B^T Phi_WA diag(1/L)] (:186-190), rows permuted back to raster order (:191)."""
    s = np.asarray(sample_indices, dtype=np.int64)
    p = len(s)
    D = phi @ (Pi * (phi.T @ np.ones(phi.shape[0])))
    alpha = 1.0 / np.mean(D)
    W_A = np.eye(p) + alpha * ((phi[:p] * Pi) @ phi[:p].T - np.diag(D[:p]))
    W_B = alpha * (phi[:p] * Pi) @ phi[p:].T
    L, U = np.linalg.eigh(W_A)
    L, U = L[::-1], U[:, ::-1]
    V = np.concatenate((U, W_B.T @ (U / L)))
    return permutation(V, s), L


def smoothing(y, sample_indices, phi, Pi):
    """smoothing() (:197-219): z = V diag(L) V^T y with (V, L) = smoothing_matrix()."""
    V, L = smoothing_matrix(sample_indices, phi, Pi)
    yv = np.asarray(y, dtype=np.float64).reshape(-1)
    return (V @ (L * (V.T @ yv))).reshape(np.shape(y))


def sharpening(y, sample_indices, phi, Pi, beta=1.5):
    """sharpening() (:222-241): with W = V diag(L) V^T, z = (1 + beta) W^2 y - beta W^3 y, beta = 1.5 (:231)."""
    V, L = smoothing_matrix(sample_indices, phi, Pi)
    yv = np.asarray(y, dtype=np.float64).reshape(-1)
    W = lambda x: V @ (L * (V.T @ x))
    w2 = W(W(yv))
    return ((1.0 + beta) * w2 - beta * W(w2)).reshape(np.shape(y))
