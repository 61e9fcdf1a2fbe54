"""Golden vectors from the reference's OWN Python pipeline (python/image_processing.py:244-357), run in the build
container.  The module does not import as it stands (matplotlib is absent and scipy.misc.imread no longer exists), so
the two plotting/IO dependencies are stubbed out -- nothing numerical is touched -- and image_processing(y) is called
exactly as the reference's __main__ does for a grey image (:405-407): 1 % spatially-uniform samples, bilateral affinity,
ALL p eigenpairs of L_A, Phi = [Phi_A; L_B^T Phi_A mu^-1] permuted back to raster order, z = y - Phi (mu + 5) Phi^T y.

Writes tests/golden/pyref_<name>.npz: the input image, the reference's z (float64) and its sample indices.
This is what pins the oracle's stages past the eigensolve (extrapolation, permutation, filter algebra) to reference code."""
import os
import sys
import types

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

# ---- stubs for the plotting / image-reading imports of python/image_processing.py:4-8 ----
mpl = types.ModuleType("matplotlib")
mpl.use = lambda *a, **k: None
plt = types.ModuleType("matplotlib.pyplot")
for name in ("figure", "plot", "savefig", "show", "imshow", "title", "close"):
    setattr(plt, name, lambda *a, **k: None)
mpl.pyplot = plt
sys.modules["matplotlib"] = mpl
sys.modules["matplotlib.pyplot"] = plt
import scipy  # noqa: E402
misc = types.ModuleType("scipy.misc")
misc.imread = lambda path: np.asarray(Image.open(path))
sys.modules["scipy.misc"] = misc
scipy.misc = misc

sys.path.insert(0, os.path.join(REF, "python"))
import image_processing as ref  # noqa: E402  (the reference module itself)
import sampling  # noqa: E402


def run(name, img, affinity=None, random_seed=None):
    """affinity: one of the reference's plugin codes (python/affinity_methods/__init__.py:8-13), default bilateral;
    random_seed: use the reference's random sampler (python/sampling/random.py) after np.random.seed(random_seed)."""
    import affinity_methods
    os.makedirs("results", exist_ok=True)            # the reference saves its eigenvalue plots there (stubbed)
    M, N = img.shape
    kw = {}
    if affinity is not None:
        kw["affinity"] = affinity
    if random_seed is not None:
        kw["sampling"] = sampling.RANDOM
        np.random.seed(random_seed)
    z, _, _ = ref.image_processing(img, **kw)
    if random_seed is not None:
        np.random.seed(random_seed)                   # the same draw again, to record it
        s = sampling.methods[sampling.RANDOM](M, N, int(M * N * 0.01))
    else:
        s = sampling.methods[sampling.SPATIALLY_UNIFORM](M, N, int(M * N * 0.01))
    out = os.path.join(HERE, f"pyref_{name}.npz")
    np.savez_compressed(out, image=img, z=np.asarray(z, dtype=np.float64), sample_indices=np.asarray(s, dtype=np.uint32),
                        kind=str(affinity if affinity is not None else affinity_methods.BILATERAL),
                        seed=-1 if random_seed is None else random_seed)
    print("wrote", out, img.shape, "p =", len(s), "z range", float(z.min()), float(z.max()))


if __name__ == "__main__":
    test = np.asarray(Image.open(os.path.join(REF, "input", "test.png")).convert("L"))
    run("test100", test)                                                   # the reference's 100 x 100 smoke input
    lion = np.asarray(Image.open(os.path.join(REF, "input", "lion.png")).convert("L"))
    crop = np.ascontiguousarray(lion[100:228, 60:220])                     # 128 x 160 crop (the n x n matrices stay small)
    run("lion_crop", crop)
    import affinity_methods
    run("lion_crop_photometric", crop, affinity=affinity_methods.PHOTOMETRIC)   # h = 10 (python/affinity_methods/photometric.py:12)
    run("test100_spatial", test, affinity=affinity_methods.SPATIAL)             # h = 10 (python/affinity_methods/spatial.py:11)
    run("lion_crop_random7", crop, random_seed=7)                               # python/sampling/random.py after np.random.seed(7)
