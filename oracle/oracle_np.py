"""CPU oracle (numpy, fp64) for the Nystroem graph-Laplacian filter path.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it, and only as the checker.  The product path
(image-processing-graph-laplacian_b200/) never imports or links this.

Parity status: the reference ships no tests or golden vectors for this path
 and its C pipeline cannot be built here (PETSc/SLEPc/MPI
absent).  The oracle is pinned instead against outputs of the reference's own
importable Python modules run in the build container -- python/sampling/* and
python/affinity_methods/{bilateral,photometric,spatial}.py -- via the fixtures
written by tests/golden/make_golden.py.  Stages past the eigensolve are commented
out in the C program at HEAD (hpc/image_processing.c:237-276) but live in the
Python prototype: tests/golden/make_golden_pyref.py runs the reference's own
python/image_processing.py:image_processing(y) and run_pipeline(), given the
prototype's constants, reproduces its output to 1e-10 (tests/test_oracle.py).
Only the C block's constants (gain +3, f(lambda) = lambda, m = p - 1, clip above
255) are "unpinned by the reference": they follow SURVEY.md section 8c.

Each function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# MT19937 -- the generator behind numpy's legacy np.random.seed / randint, which
# python/sampling/random.py:10-12 draws from.
# ----------------------------------------------------------------------------


class MT19937:
    """Plain restatement of the published MT19937 algorithm (Matsumoto &
    Nishimura 1998): init_genrand(seed) + genrand_int32()."""

    N, M = 624, 397

    def __init__(self, seed: int):
        mt = np.empty(self.N, dtype=np.uint64)
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, self.N):
            prev = int(mt[i - 1])
            mt[i] = (1812433253 * (prev ^ (prev >> 30)) + i) & 0xFFFFFFFF
        self.mt = mt.astype(np.uint32)
        self.pos = self.N

    def _twist(self):
        mt = self.mt
        N, M = self.N, self.M
        # three dependency-free phases, the usual way to vectorise the twist
        for lo, hi in ((0, N - M), (N - M, 2 * (N - M)), (2 * (N - M), N - 1)):
            hi = min(hi, N - 1)
            if lo >= hi:
                continue
            y = (mt[lo:hi] & np.uint32(0x80000000)) | (mt[lo + 1:hi + 1] & np.uint32(0x7FFFFFFF))
            src = (np.arange(lo, hi) + M) % N
            mag = np.where((y & np.uint32(1)) != 0, np.uint32(0x9908B0DF), np.uint32(0))
            mt[lo:hi] = mt[src] ^ (y >> np.uint32(1)) ^ mag
        y = (mt[N - 1] & np.uint32(0x80000000)) | (mt[0] & np.uint32(0x7FFFFFFF))
        mag = np.uint32(0x9908B0DF) if (int(y) & 1) else np.uint32(0)
        mt[N - 1] = mt[M - 1] ^ (y >> np.uint32(1)) ^ mag
        self.pos = 0

    def next_u32_block(self) -> np.ndarray:
        """Return the remaining tempered words of the current block."""
        if self.pos >= self.N:
            self._twist()
        y = self.mt[self.pos:].copy()
        self.pos = self.N
        y ^= y >> np.uint32(11)
        y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
        y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
        y ^= y >> np.uint32(18)
        return y

    def stream(self):
        while True:
            for v in self.next_u32_block():
                yield int(v)


def _mask_for(rng: int) -> int:
    mask = rng
    for sh in (1, 2, 4, 8, 16):
        mask |= mask >> sh
    return mask


# ----------------------------------------------------------------------------
# a-1  Sampling
# ----------------------------------------------------------------------------


def uniform_sampling(width: int, height: int, sample_size: int) -> np.ndarray:
    """hpc/sampling.c:6-23 (UniformSampling); same grid as
    python/sampling/spatially_uniform.py:9-24.  Returns ascending uint32 raster
    indices; len() is the rewritten *sample_size."""
    sample_dist = int(np.sqrt((width * height) // sample_size))  # integer division first (:8)
    xy0 = sample_dist // 2
    rows = range(xy0, height - 1, sample_dist)
    cols = range(xy0, width - 1, sample_dist)
    out = np.empty(len(rows) * len(cols), dtype=np.uint32)
    c = 0
    for i in rows:
        for j in cols:
            out[c] = width * i + j
            c += 1
    return out


def random_sampling(width: int, height: int, sample_size: int, seed: int) -> np.ndarray:
    """python/sampling/random.py:8-16 with np.random.seed(seed) in front.

    The legacy randint(0, n) draws one MT19937 word, masks it to the smallest
    2^k-1 >= n-1 and rejects values > n-1; the reference takes the set of the
    first `sample_size` draws and tops it up one draw at a time until it holds
    `sample_size` distinct values, then sorts.  That equals: the first
    `sample_size` distinct accepted values of the stream, sorted."""
    n = width * height
    rng = n - 1
    mask = _mask_for(rng)
    seen = set()
    gen = MT19937(seed).stream()
    # first batch of exactly sample_size accepted draws (duplicates collapse)
    taken = 0
    while taken < sample_size:
        v = next(gen) & mask
        if v <= rng:
            seen.add(v)
            taken += 1
    while len(seen) < sample_size:
        v = next(gen) & mask
        if v <= rng:
            seen.add(v)
    return np.sort(np.fromiter(seen, dtype=np.uint32, count=len(seen)))


# ----------------------------------------------------------------------------
# Synthetic images (SURVEY.md 8d asks for one bit-reproducible generator).
# Integer-only so numpy, C and CUDA agree bit for bit: smooth low-frequency
# product of triangle waves + 64-pixel checkerboard edges + hashed noise.
# ----------------------------------------------------------------------------

_SYN_PERIODS = ((97, 131), (113, 89), (71, 149))


def _hash32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x7FEB352D)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x846CA68B)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


def synthetic_image(width: int, height: int, channels: int = 1, seed: int = 1234) -> np.ndarray:
    """u8 image [H, W] (channels=1) or [H, W, 3]."""
    r = np.arange(height, dtype=np.int64)[:, None]
    c = np.arange(width, dtype=np.int64)[None, :]
    out = np.empty((height, width, channels), dtype=np.uint8)
    for ch in range(channels):
        pr, pc = _SYN_PERIODS[ch]
        tr = 64 - np.abs(((r % pr) * 256) // pr - 128)      # [-64, 64]
        tc = 64 - np.abs(((c % pc) * 256) // pc - 128)
        smooth = (80 * tr * tc + 4096 * 80) // 4096 - 80      # floor division on a non-negative number
        edges = 24 * (((r >> 6) + (c >> 6)) & 1)
        idx = (r * width + c) * channels + ch
        with np.errstate(over="ignore"):
            h = _hash32((idx & 0xFFFFFFFF).astype(np.uint32) + np.uint32((seed * 0x9E3779B1) & 0xFFFFFFFF))
        noise = (h % np.uint32(33)).astype(np.int64) - 16
        v = 128 + smooth + edges + noise
        out[:, :, ch] = np.clip(v, 0, 255).astype(np.uint8)
    return out[:, :, 0] if channels == 1 else out


# ----------------------------------------------------------------------------
# a-2  Affinity
# ----------------------------------------------------------------------------

BILATERAL, PHOTOMETRIC, SPATIAL, NLM = "bilateral", "photometric", "spatial", "nlm"
NLM_KSZ, NLM_SIGMA = 7, 1.2     # python/affinity_methods/NLM.py:13,17 (patch size, Gaussian weight of the patch elements)


def _as_hwc(img: np.ndarray) -> np.ndarray:
    img = np.asarray(img)
    return img[:, :, None] if img.ndim == 2 else img


def affinity_rows(img, sample_indices, cols, kind=BILATERAL, h_loc=40.0, h_val=30.0):
    """K(samples, cols) in fp64.

    hpc/affinity.c:59-113 (ComputeBilateralFilter), :8-17 (photometric), :19-57
    (spatial); Python twins python/affinity_methods/{bilateral,photometric,
    spatial}.py.  The reference multiplies two exponentials (affinity.c:99,107,
    110); colour (no counterpart in the reference) sums the
    squared channel differences in the photometric term."""
    if kind == NLM:
        return nlm_affinity_rows(img, sample_indices, cols, h_val)
    img = _as_hwc(img).astype(np.float64)
    H, W, C = img.shape
    flat = img.reshape(H * W, C)
    s = np.asarray(sample_indices, dtype=np.int64)
    q = np.asarray(cols, dtype=np.int64)
    sr, sc = s // W, s % W          # hpc/utils.c:11-19: num2x = row, num2y = col
    qr, qc = q // W, q % W
    K = np.ones((len(s), len(q)), dtype=np.float64)
    if kind in (BILATERAL, SPATIAL):
        d2 = (sr[:, None] - qr[None, :]) ** 2 + (sc[:, None] - qc[None, :]) ** 2
        K *= np.exp(-(d2.astype(np.float64)) / (h_loc * h_loc))
    if kind in (BILATERAL, PHOTOMETRIC):
        dv2 = np.zeros((len(s), len(q)), dtype=np.float64)
        for ch in range(C):
            dv2 += (flat[s, ch][:, None] - flat[q, ch][None, :]) ** 2
        K *= np.exp(-dv2 / (h_val * h_val))
    return K


def nlm_patch_weights() -> np.ndarray:
    """matlab_style_gauss2D((7,7), 1.2) normalised to sum 1 (python/utils.py:19-32, NLM.py:17-19), as [7][7]."""
    m = (NLM_KSZ - 1) / 2.0
    y, x = np.ogrid[-m:m + 1, -m:m + 1]
    h = np.exp(-(x * x + y * y) / (2.0 * NLM_SIGMA * NLM_SIGMA))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    h /= h.sum()
    return h / h.sum()          # NLM.py:19 normalises once more


def nlm_affinity_rows(img, sample_indices, cols, h=3.0):
    """Non-local-means patch affinity K(samples, cols), fp64 (python/affinity_methods/NLM.py:9-34):
    K = exp(-sum_k (G_k (P_s[k] - P_q[k]))^2 / h^2) over the 7x7 patches P around the two pixels of the symmetrically
    padded image (np.pad 'symmetric', NLM.py:16), G the normalised Gaussian patch weights, h = 3 (NLM.py:12).
    Rows AND columns are raster indices here; the reference's columns come out in column-major pixel order
    (im2col of the transposed image, NLM.py:21) although its callers read them as raster indices -- see
    tests/golden/make_golden_nlm.py and tests/test_oracle.py for the map that pins this function to the reference."""
    img = np.asarray(img)
    assert img.ndim == 2 or img.shape[2] == 1, "NLM affinity is defined on one channel (the reference filters luminance)"
    y = img.reshape(img.shape[0], img.shape[1]).astype(np.float64)
    H, W = y.shape
    rad = (NLM_KSZ - 1) // 2
    pad = np.pad(y, (rad, rad), "symmetric")
    G = nlm_patch_weights()
    s = np.asarray(sample_indices, dtype=np.int64)
    q = np.asarray(cols, dtype=np.int64)

    def patches(idx):
        r, c = idx // W, idx % W
        out = np.empty((len(idx), NLM_KSZ, NLM_KSZ))
        for a in range(NLM_KSZ):
            for b in range(NLM_KSZ):
                out[:, a, b] = pad[r + a, c + b] * G[a, b]
        return out.reshape(len(idx), -1)
    Ps, Pq = patches(s), patches(q)
    d2 = np.empty((len(s), len(q)))
    for i in range(len(s)):
        d2[i] = ((Pq - Ps[i][None, :]) ** 2).sum(axis=1)
    return np.exp(-d2 / (h * h))


def non_sample_indices(n: int, sample_indices) -> np.ndarray:
    """Ascending raster order of the pixels that are not samples
    (hpc/affinity.c:218-235)."""
    mask = np.ones(n, dtype=bool)
    mask[np.asarray(sample_indices, dtype=np.int64)] = False
    return np.nonzero(mask)[0]


# ----------------------------------------------------------------------------
# a-3 .. a-9  Laplacian, eigensolve, Nystroem, permutation, filter
# ----------------------------------------------------------------------------


def laplacian(K_A, rowsum_B):
    """hpc/laplacian.c:14-42: D_A = rowsum(K_A)+rowsum(K_B); alpha = 1/mean(D_A);
    L_A = alpha (D_A - K_A).  L_B = -alpha K_B is folded by the callers."""
    D = K_A.sum(axis=1) + rowsum_B
    alpha = 1.0 / D.mean()
    L_A = alpha * (np.diag(D) - K_A)
    return D, alpha, L_A


def smallest_eigenpairs(L_A, m):
    """Converged m smallest eigenpairs, ascending: what
    hpc/eigendecomposition.c:121-124 asks SLEPc for and what
    hpc/inverse_power_it.c:86-252 approximates."""
    mu, U = np.linalg.eigh(L_A)
    return mu[:m], U[:, :m]


def gram_schmidt(X):
    """hpc/gram_schmidt.c:29-64 (classical Gram-Schmidt, projections taken
    against the already-normalised u_j, then normalise).  Column-by-column,
    fp64.  Returns (Q, norms_before_normalisation)."""
    X = np.array(X, dtype=np.float64, copy=True)
    norms = np.empty(X.shape[1])
    for k in range(X.shape[1]):
        v = X[:, k].copy()
        if k:
            U = X[:, :k]
            coef = (U.T @ v) / np.einsum("ij,ij->j", U, U)   # <v,u>/<u,u> (:14-16)
            v = v - U @ coef
        norms[k] = np.linalg.norm(v)
        X[:, k] = v / norms[k]
    return X, norms


def inverse_iteration_start(p, m):
    """The fixed pseudo-random start X_0 (p x m, uniform in (0, 1)) shared with the device solver (csrc/dense_small.cu: k_ii_start).
    The reference fills X_0 from PETSc's generator seeded with the MPI rank (hpc/inverse_power_it.c:27-34), which nothing outside
    PETSc can reproduce and which makes its result depend on the process count."""
    idx = (np.arange(p * m, dtype=np.uint64) + np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF)
    h = _hash32(idx.astype(np.uint32))
    return (((h >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0).reshape(p, m)


def inverse_power_iteration(A, m, opti_gs=1, epsilon=0.1, X0=None, max_iterations=1000):
    """InversePowerIteration, hpc/inverse_power_it.c:86-252, with exact solves in place of its GMRES + additive-Schwarz ones:

      X_0 random, orthonormalised (:94-95); r = |(I - X X^T) A X|_F (ComputeResidualsNorm :49-80);
      while r > epsilon: X <- A^-1 X (:163-166); keep a copy (:169); every opti_gs-th step OrthonormaliseVecs with the norms out
      (:172-175); recompute r (:178); one more orthonormalisation if the last step skipped it (:183-186);
      lambda_i = 1 / norms[i] (:204); eigenvectors = the kept pre-orthonormalisation iterates, normalised (:230-235).

    Returns (lambda, vectors, outer iterations, last residual)."""
    A = np.asarray(A, dtype=np.float64)
    p = A.shape[0]
    X = inverse_iteration_start(p, m) if X0 is None else np.array(X0, dtype=np.float64, copy=True)
    opti_gs = max(1, int(opti_gs))
    X, norms = gram_schmidt(X)
    Xb = X.copy()

    def residual(Xk):
        AX = A @ Xk
        return float(np.linalg.norm(AX - Xk @ (Xk.T @ AX)))

    r, it = residual(X), 0
    while r > epsilon and it < max_iterations:
        it += 1
        X = np.linalg.solve(A, X)
        Xb = X.copy()
        if it % opti_gs == 0:
            X, norms = gram_schmidt(X)
        r = residual(X)
    if opti_gs != 1 and it % opti_gs != 0:
        X, norms = gram_schmidt(X)
    return 1.0 / norms, Xb / np.linalg.norm(Xb, axis=0), it, r


def run_pipeline(img, sample_indices, m=None, kind=BILATERAL, h_loc=40.0, h_val=30.0,
                 gain=3.0, power=1.0, orthonormalise=False, chunk=65536, return_phi=False,
                 f_of_mu=None, clip=True, all_pairs=False):
    """The restored block hpc/image_processing.c:183-275 in fp64:

      s -> K_A, K_B (affinity.c:129-262) -> D, alpha, L_A (laplacian.c:14-42) ->
      (mu, U) m smallest -> Phi[s] = U, Phi[rest] = (-alpha K_B)^T U diag(1/mu)
      (nystroem.c:25-57 + utils.c:134-173 permutation back to raster order) ->
      z = y + gain * Phi (mu^power o (Phi^T y)) (display.c:64-73; MatPow is a
      no-op, utils.c:721, hence power=1), z[z>255] = 255 (display.c:76).

    Returns a dict.  z is float64 [H, W] or [H, W, C]; negatives are NOT clipped
    (SURVEY 8c-iii).

    The Python prototype runs the same stages with other constants
    (python/image_processing.py:274-305): all p eigenpairs (`all_pairs`), the filter
    function f(mu) = mu + 5 (`f_of_mu`), gain -1 and no clipping (`clip=False`); with
    those arguments this function must reproduce the reference's own output, which is
    what tests/test_oracle.py checks against tests/golden/pyref_*.npz."""
    imgc = _as_hwc(img)
    H, W, C = imgc.shape
    n = H * W
    s = np.asarray(sample_indices, dtype=np.int64)
    p = len(s)
    if all_pairs:
        m = p                                       # python/image_processing.py:287-292 keeps every pair
    elif m is None or m < 0 or m >= p:
        m = p - 1                                   # image_processing.c:102-106
    rest = non_sample_indices(n, s)
    fmu = (lambda mu_: mu_[:, None] ** power) if f_of_mu is None else (lambda mu_: np.asarray(f_of_mu(mu_))[:, None])

    K_A = affinity_rows(imgc, s, s, kind, h_loc, h_val)
    rowsum_B = np.zeros(p)
    for a in range(0, len(rest), chunk):
        rowsum_B += affinity_rows(imgc, s, rest[a:a + chunk], kind, h_loc, h_val).sum(axis=1)
    D, alpha, L_A = laplacian(K_A, rowsum_B)
    mu, U = smallest_eigenpairs(L_A, m)

    Wm = (-alpha) * U / mu[None, :]                # p x m  (L_B^T . Phi_A . Lambda^-1, nystroem.c:41-42)
    y = imgc.reshape(n, C).astype(np.float64)

    phi = None
    if orthonormalise or return_phi:
        phi = np.empty((n, m))
        phi[s] = U
        for a in range(0, len(rest), chunk):
            idx = rest[a:a + chunk]
            phi[idx] = affinity_rows(imgc, s, idx, kind, h_loc, h_val).T @ Wm
        if orthonormalise:
            phi, _ = gram_schmidt(phi)
        c = phi.T @ y
        z = y + gain * (phi @ (fmu(mu) * c))
    else:
        # same arithmetic, streamed so Phi (n x m fp64) is never held
        c = U.T @ y[s]
        for a in range(0, len(rest), chunk):
            idx = rest[a:a + chunk]
            KB = affinity_rows(imgc, s, idx, kind, h_loc, h_val)
            c += Wm.T @ (KB @ y[idx])
        w = fmu(mu) * c                            # m x C
        z = y.copy()
        z[s] += gain * (U @ w)
        Ww = Wm @ w                                # p x C
        for a in range(0, len(rest), chunk):
            idx = rest[a:a + chunk]
            KB = affinity_rows(imgc, s, idx, kind, h_loc, h_val)
            z[idx] += gain * (KB.T @ Ww)
    if clip:
        z = np.minimum(z, 255.0)                   # AboveXSetY(z, 255, 255), display.c:76
    z = z.reshape(H, W, C)
    out = dict(sample_indices=s.astype(np.uint32), K_A=K_A, D=D, alpha=alpha, L_A=L_A, mu=mu,
               U=U, z=z[:, :, 0] if np.asarray(img).ndim == 2 else z, m=m, p=p)
    if return_phi:
        out["phi"] = phi
    return out


def run_full(img, kind=BILATERAL, h_loc=40.0, h_val=30.0, chunk=2048):
    """The reference's -no_approx mode in fp64: K is the n x n affinity of ALL pixels (hpc/affinity.c:264-336),
    D = K.1, alpha = 1/mean(D), L = alpha (diag D - K) (hpc/laplacian.c:44-65), z = y - L y clipped to [0, 255]
    (hpc/display.c:128-149).  Rows of K are formed in chunks; O(n^2) time, small images only."""
    imgc = _as_hwc(img)
    H, W, C = imgc.shape
    n = H * W
    y = imgc.reshape(n, C).astype(np.float64)
    allpx = np.arange(n)
    D = np.empty(n)
    Ky = np.empty((n, C))
    for a in range(0, n, chunk):
        K = affinity_rows(imgc, allpx[a:a + chunk], allpx, kind, h_loc, h_val)
        D[a:a + chunk] = K.sum(axis=1)
        Ky[a:a + chunk] = K @ y
    alpha = 1.0 / D.mean()
    z = y - alpha * (D[:, None] * y - Ky)
    z = np.clip(z, 0.0, 255.0).reshape(H, W, C)
    return dict(D=D, alpha=alpha, z=z[:, :, 0] if np.asarray(img).ndim == 2 else z)


def quantise(z):
    """OneColMat2pngbytes (hpc/utils.c:492-534) casts double -> png_byte; negative
    inputs are UB there, so the build clamps to [0,255] then truncates
    (SURVEY 8c-iii)."""
    return np.clip(z, 0.0, 255.0).astype(np.uint8)
