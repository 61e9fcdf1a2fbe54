"""Full-size golden for BASELINE.json config 5 (synthetic 8192x8192 colour, p=2000 random samples seed 0, m=1999, bilateral affinity
on position + RGB -- the colour affinity does not exist in the reference, SURVEY 8c-vii: parity unpinned, oracle = same formula).

Run once in the build container (~30-40 minutes on 8 cores).  Same arithmetic as make_golden_c4.py, per channel: the restored block
hpc/image_processing.c:183-275 in fp64 with the kernel rows evaluated by the C/OpenMP oracle in chunks, so that neither K_B (1 TB)
nor Phi is ever held.  Kept compact: eigenvalues, D, z on a lattice of pixels (all channels), the sums of z and of (z - y)^2."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_c as oc  # noqa: E402
from oracle import oracle_np as o  # noqa: E402

W, H, C, P, SEED_IMG, SEED_S = 8192, 8192, 3, 2000, 1234, 0
STRIDE = 769   # lattice of pixels kept: every 769th raster index (prime, coprime with the width): 87 267 pixels
CHUNK = 1 << 16

if __name__ == "__main__":
    img = oc.synthetic_image(W, H, C, SEED_IMG)
    s = oc.random_sampling(W, H, P, SEED_S).astype(np.int64)
    n, p, m = W * H, P, P - 1
    y = img.reshape(n, C).astype(np.float64)
    t0 = time.time()
    D = np.zeros(p)
    T = np.zeros((p, C))
    for a in range(0, n, CHUNK):
        q = np.arange(a, min(n, a + CHUNK), dtype=np.uint32)
        K = oc.affinity_rows(img, s, q)
        D += K.sum(axis=1)
        T += K @ y[a:a + len(q)]
        if (a // CHUNK) % 128 == 0:
            print("pass 1: chunk %d / %d, %.0f s" % (a // CHUNK, n // CHUNK, time.time() - t0), flush=True)
    K_A = oc.affinity_rows(img, s, s.astype(np.uint32))
    alpha = 1.0 / D.mean()
    L_A = alpha * (np.diag(D) - K_A)
    mu, U = o.smallest_eigenpairs(L_A, m)
    Wm = (-alpha) * U / mu[None, :]                           # nystroem.c:41-42
    c = U.T @ y[s] + Wm.T @ (T - K_A @ y[s])                  # Phi^T y per channel
    w = mu[:, None] * c                                       # MatPow no-op (utils.c:721): f(lambda) = lambda
    Ww = Wm @ w                                               # p x C
    z = y.copy()
    for a in range(0, n, CHUNK):
        q = np.arange(a, min(n, a + CHUNK), dtype=np.uint32)
        K = oc.affinity_rows(img, s, q)
        z[a:a + len(q)] += 3.0 * (K.T @ Ww)                   # display.c:73
        if (a // CHUNK) % 128 == 0:
            print("pass 2: chunk %d / %d, %.0f s" % (a // CHUNK, n // CHUNK, time.time() - t0), flush=True)
    z[s] = y[s] + 3.0 * (U @ w)                               # sample rows of Phi are Phi_A (nystroem.c:25-34)
    z = np.minimum(z, 255.0)                                  # display.c:76
    idx = np.arange(0, n, STRIDE)
    out = os.path.join(HERE, "c5_full.npz")
    np.savez_compressed(out, width=W, height=H, channels=C, p=P, seed_img=SEED_IMG, seed_samples=SEED_S, stride=STRIDE, mu=mu, D=D,
                        z_lattice=z[idx].astype(np.float32), sum_z=z.sum(), sum_dz2=((z - y) ** 2).sum(),
                        norm_z=np.linalg.norm(z), sample_indices=s.astype(np.uint32))
    print("total: %.1f s; wrote %s (%d bytes)" % (time.time() - t0, out, os.path.getsize(out)), flush=True)
