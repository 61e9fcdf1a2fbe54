"""Golden vectors for the NLM patch affinity from the reference's OWN module (python/affinity_methods/NLM.py:9-34), run in
the build container: K_AB = NLM_affinity(y, sample_indices) for two small grey images (one of them not square).

What the reference's function returns, established by reading it and checked by tests/test_oracle.py: row i is sample i
(raster index -> (row, col) with num2xy, NLM.py:26), but the COLUMNS follow im2col(img.T) (NLM.py:21), i.e. column j is
the pixel with COLUMN-major index j = col * M + row, while every caller treats j as a raster index
(python/image_processing.py:60-64).  The oracle and the CUDA kernel use raster order for both (the evidently intended
matrix); the fixture keeps the reference's output untouched and the test applies the index map.

Writes tests/golden/pyref_nlm_<name>.npz: image (uint8), sample_indices, K_AB (float64 [p][M*N]) as returned."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "python"))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import affinity_methods  # noqa: E402  (the reference package)
import sampling  # noqa: E402
from oracle import oracle_np as o  # noqa: E402  (only for the synthetic test image)


def run(name, img, p_req):
    M, N = img.shape
    s = sampling.methods[sampling.SPATIALLY_UNIFORM](M, N, p_req)
    K = affinity_methods.methods[affinity_methods.NLM](img.astype(np.float64), s)
    out = os.path.join(HERE, f"pyref_nlm_{name}.npz")
    np.savez_compressed(out, image=img, sample_indices=np.asarray(s, dtype=np.uint32), K_AB=np.asarray(K, dtype=np.float64))
    print(out, img.shape, "p =", len(s), "K range", K.min(), K.max())


if __name__ == "__main__":
    # smooth synthetic images with +-2 grey levels of noise: with h = 3 the kernel is then neither all ones nor all zeros
    rng = np.random.RandomState(5)
    for name, (H, W) in (("sq24", (24, 24)), ("rect", (20, 31))):
        r, c = np.mgrid[0:H, 0:W]
        img = np.clip(120 + 6 * np.sin(r / 5.0) + 5 * np.cos(c / 7.0) + rng.randint(-2, 3, (H, W)), 0, 255).astype(np.uint8)
        run(name, img, 12)
