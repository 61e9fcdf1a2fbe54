"""Full-size golden for BASELINE.json config 4 (synthetic 3840x2160 grey, p=1000 random samples seed 0, m=999).

Run once in the build container (~10 minutes on 8 cores).  Same arithmetic as oracle_np.run_pipeline's streamed branch
(the restored block hpc/image_processing.c:183-275, fp64), with the kernel rows evaluated by the C/OpenMP oracle
(oracle.c: orc_affinity_rows) in chunks of 65 536 pixels so that neither K_B (66 GB) nor Phi (66 GB) is ever held.
Kept compact: eigenvalues, D, z on a fixed lattice of pixels, the sum of z and of (z - y)^2.  The image and the samples
are regenerated on the test side from their seeds (bit-exact generators)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_c as oc  # noqa: E402
from oracle import oracle_np as o  # noqa: E402

W, H, P, SEED_IMG, SEED_S = 3840, 2160, 1000, 1234, 0
STRIDE = 97   # lattice of pixels kept: every 97th raster index (coprime with the width)
CHUNK = 1 << 16

if __name__ == "__main__":
    img = oc.synthetic_image(W, H, 1, SEED_IMG)
    s = oc.random_sampling(W, H, P, SEED_S).astype(np.int64)
    n, p, m = W * H, P, P - 1
    y = img.reshape(-1).astype(np.float64)
    t0 = time.time()
    # pass 1: D = rowsum(K_A) + rowsum(K_B) and T = [K_A K_B] y over ALL pixels (hpc/laplacian.c:18-20)
    D = np.zeros(p)
    T = np.zeros(p)
    for a in range(0, n, CHUNK):
        q = np.arange(a, min(n, a + CHUNK))
        K = oc.affinity_rows(img, s, q)
        D += K.sum(axis=1)
        T += K @ y[q]
    print("pass 1: %.1f s" % (time.time() - t0), flush=True)
    K_A = oc.affinity_rows(img, s, s)
    alpha = 1.0 / D.mean()
    L_A = alpha * (np.diag(D) - K_A)
    mu, U = o.smallest_eigenpairs(L_A, m)
    Wm = (-alpha) * U / mu[None, :]                           # nystroem.c:41-42
    c = U.T @ y[s] + Wm.T @ (T - K_A @ y[s])                  # Phi^T y: sample rows are Phi_A, the rest K_B^T Wm
    w = mu * c                                                # MatPow no-op (utils.c:721): f(lambda) = lambda
    Ww = Wm @ w
    z = y.copy()
    for a in range(0, n, CHUNK):
        q = np.arange(a, min(n, a + CHUNK))
        K = oc.affinity_rows(img, s, q)
        z[q] += 3.0 * (K.T @ Ww)                              # display.c:73
    z[s] = y[s] + 3.0 * (U @ w)                               # sample rows of Phi are Phi_A (nystroem.c:25-34)
    z = np.minimum(z, 255.0)                                  # display.c:76
    print("total: %.1f s" % (time.time() - t0), flush=True)
    idx = np.arange(0, n, STRIDE)
    out = os.path.join(HERE, "c4_full.npz")
    np.savez_compressed(out, width=W, height=H, p=P, seed_img=SEED_IMG, seed_samples=SEED_S, stride=STRIDE, mu=mu, D=D,
                        z_lattice=z[idx].astype(np.float32), sum_z=z.sum(), sum_dz2=((z - y) ** 2).sum(),
                        norm_z=np.linalg.norm(z), sample_indices=s.astype(np.uint32))
    print("wrote", out, os.path.getsize(out))
