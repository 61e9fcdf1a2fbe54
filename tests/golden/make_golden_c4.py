"""Full-size golden for BASELINE.json config 4 (synthetic 3840x2160 grey, p=1000 random samples seed 0, m=999):
the C/OpenMP fp64 oracle (oracle/oracle.c) run once in the build container (~minutes on 8 cores); kept compact:
the eigenvalues, D, and z on a fixed lattice of pixels plus the exact sum of z and of (z - y)^2.
The image and the samples are regenerated on the test side from their seeds (bit-exact generators)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_c as oc  # noqa: E402

W, H, P, SEED_IMG, SEED_S = 3840, 2160, 1000, 1234, 0
STRIDE = 97   # lattice of pixels kept: every 97th raster index (coprime with the width)

if __name__ == "__main__":
    img = oc.synthetic_image(W, H, 1, SEED_IMG)
    s = oc.random_sampling(W, H, P, SEED_S)
    t = time.time()
    r = oc.run_pipeline(img, s)
    print("oracle C4: %.1f s" % (time.time() - t))
    z = np.asarray(r["z"], dtype=np.float64).reshape(-1)
    y = img.reshape(-1).astype(np.float64)
    idx = np.arange(0, W * H, STRIDE)
    np.savez_compressed(os.path.join(HERE, "c4_full.npz"), width=W, height=H, p=P, seed_img=SEED_IMG, seed_samples=SEED_S,
                        stride=STRIDE, mu=r["mu"], D=r["D"], z_lattice=z[idx].astype(np.float32), sum_z=z.sum(),
                        sum_dz2=((z - y) ** 2).sum(), norm_z=np.linalg.norm(z), sample_indices=s)
    print("wrote c4_full.npz", os.path.getsize(os.path.join(HERE, "c4_full.npz")))
