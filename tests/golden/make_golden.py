"""Regenerates the fixtures in tests/golden/ (run in the build container only;
/root/reference does not exist on the GPU box).

What comes from the REFERENCE ITSELF (imported from /root/reference/python):
  * sample indices from python/sampling/spatially_uniform.py and random.py
    (random: np.random.seed(seed) first, the reference is unseeded);
  * affinity rows from python/affinity_methods/{bilateral,photometric,spatial}.py,
    stored as K_A = K[:, samples] and the row sums D = K.sum(axis=1)
    (= rowsum(K_A) + rowsum(K_B), hpc/laplacian.c:18-20).
What comes from oracle/oracle_np.py (stages the reference cannot run here,
SURVEY.md 8c): alpha, mu, z.  Those are "unpinned by the reference".

Inputs are the decoded grey pixels of the reference's input/*.png (decoded with
PIL), stored as arrays so the GPU box needs neither the PNGs nor a PNG decoder.
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "python"))
sys.path.insert(0, ROOT)

import affinity_methods  # noqa: E402  (reference module)
import sampling  # noqa: E402  (reference module)

from oracle import oracle_np as o  # noqa: E402


def load(name):
    return np.asarray(Image.open(os.path.join(REF, "input", name + ".png")).convert("L"))


def ref_sampling(H, W, p, method, seed):
    if method == "random":
        np.random.seed(seed)
        return np.asarray(sampling.methods["random"](H, W, p), dtype=np.uint32)
    return np.asarray(sampling.methods["spatially_uniform"](H, W, p), dtype=np.uint32)


def case(tag, img, p_req, method, seed, kind, ref_kind=None, h_loc=40.0, h_val=30.0, rgb=False, m=None):
    H, W = img.shape
    s = ref_sampling(H, W, p_req, method, seed)
    out = dict(image=img, sample_indices=s, p_req=p_req, method=method, seed=seed, kind=kind,
               h_loc=h_loc, h_val=h_val, rgb=int(rgb))
    if ref_kind is not None:
        K = affinity_methods.methods[ref_kind](img, s)       # p x n, reference code
        out["ref_D"] = K.sum(axis=1)
        out["ref_K_A"] = K[:, s.astype(np.int64)]
        del K
    src = np.repeat(img[:, :, None], 3, axis=2) if rgb else img
    r = o.run_pipeline(src, s, m=m, kind=kind, h_loc=h_loc, h_val=h_val)
    out.update(D=r["D"], alpha=r["alpha"], mu=r["mu"], z=r["z"].astype(np.float32), m=r["m"])
    r2 = o.run_pipeline(src, s, m=m, kind=kind, h_loc=h_loc, h_val=h_val, orthonormalise=True)
    out["z_gs"] = r2["z"].astype(np.float32)
    path = os.path.join(HERE, tag + ".npz")
    np.savez_compressed(path, **out)
    print(tag, "p=%d m=%d" % (len(s), r["m"]), "z-y rel %.3e" % (np.linalg.norm(r["z"] - src) / np.linalg.norm(src)),
          "%.1f KB" % (os.path.getsize(path) / 1024))


def sampling_table():
    """Sample index sets for the BASELINE.json config sizes, from the reference
    modules (bit-exactness target for the device samplers)."""
    out = {}
    for (W, H, p) in ((450, 300, 50), (512, 512, 256), (350, 350, 500), (3840, 2160, 1000), (100, 100, 100),
                      (8192, 8192, 2000), (721, 558, 4023)):
        out["uniform_%dx%d_%d" % (W, H, p)] = ref_sampling(H, W, p, "uniform", 0)
        for seed in (0, 1, 1234):
            out["random_%dx%d_%d_s%d" % (W, H, p, seed)] = ref_sampling(H, W, p, "random", seed)
    np.savez_compressed(os.path.join(HERE, "sampling.npz"), **out)
    print("sampling table:", len(out), "sets")


if __name__ == "__main__":
    sampling_table()
    # C0: smoke-sized input, 1 % uniform as hpc/image_processing.c:187
    case("test_uniform100", load("test"), 100, "uniform", 0, "bilateral", ref_kind="bilateral")
    # BASELINE.json config 1: cat_small grey, p=50 random (seed 0)
    case("cat_small_random50", load("cat_small"), 50, "random", 0, "bilateral", ref_kind="bilateral")
    # config 2: barbara grey, p=256 uniform
    case("barbara_uniform256", load("barbara"), 256, "uniform", 0, "bilateral", ref_kind="bilateral")
    # config 3: lion expanded to RGB, photometric-RGB affinity (h_val 30), p=500 random seed 0
    # (no counterpart in the reference: oracle only)
    case("lion_rgb_photometric500", load("lion"), 500, "random", 0, "photometric", rgb=True)
    # reference photometric / spatial plugins (grey, their own h=10) pin the two other kernels
    case("lion_photometric_h10", load("lion"), 100, "uniform", 0, "photometric", ref_kind="photometric", h_val=10.0)
    case("test_spatial_h10", load("test"), 100, "uniform", 0, "spatial", ref_kind="spatial", h_loc=10.0)
    # partial spectrum (-num_eigvals)
    case("cat_small_uniform_m20", load("cat_small"), 200, "uniform", 0, "bilateral", m=20)
