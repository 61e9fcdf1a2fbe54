"""Golden vectors for the experimental building blocks of the reference's Python prototype, produced by the
reference's OWN functions in python/image_processing.py, run in the build container on a small image:

    affinity(y, s, bilateral)      :52-65    K_A, K_B (sample rows; K_B over the non-sample pixels in raster order)
    nystroem(K_A, K_B)             :68-89    Phi = [Phi_A; K_B^T Phi_A Pi^-1], Pi (SVD of K_A)
    permutation(phi, s)            :35-50    rows back to raster order
    orthogonalisation(K_A, K_B)    :112-129  one-shot orthogonal Nystroem eigenvectors V, eigenvalues min(Pi_Q, 1)
    sinkhorn(phi, Pi)              :92-109   100 Sinkhorn iterations on Phi Pi Phi^T, then W_A, W_B
    smoothing_matrix(s, phi, Pi)   :151-194  W = I + alpha (K - D) of K = Phi Pi Phi^T, eigenpairs of its sample block, extended (V, L)
    smoothing(y, s, phi, Pi)       :197-219  z = V L V^T y
    sharpening(y, s, phi, Pi)      :222-241  z = (1 + beta) W^2 y - beta W^3 y, W = V L V^T, beta = 1.5

The module's matplotlib / scipy.misc imports are stubbed exactly as in make_golden_pyref.py; nothing numerical is touched.
Writes tests/golden/pyref_proto.npz.  The oracle restatements (oracle/proto_np.py) are pinned to it by tests/test_oracle.py;
the device versions (csrc/proto.cu) are compared with the same file by tests/test_proto_gpu.py."""
import os
import sys
import types

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

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
import affinity_methods  # noqa: E402
import image_processing as ref  # noqa: E402
import sampling  # noqa: E402

if __name__ == "__main__":
    lion = np.asarray(Image.open(os.path.join(REF, "input", "lion.png")).convert("L"))
    y = np.ascontiguousarray(lion[120:144, 80:110]).astype(np.float64)           # 24 x 30 crop
    M, N = y.shape
    s = sampling.methods[sampling.SPATIALLY_UNIFORM](M, N, 12)
    K_A, K_B = ref.affinity(y, s, affinity_methods.methods[affinity_methods.BILATERAL])
    phi, Pi = ref.nystroem(K_A, K_B)
    phi_perm = ref.permutation(phi, s)
    V, Pi_V = ref.orthogonalisation(K_A.copy(), K_B.copy())
    W_A, W_B = ref.sinkhorn(phi, Pi)
    ref.display_or_save = lambda *a, **k: None                                   # smoothing() saves three eigenvector pictures
    V_s, L_s = ref.smoothing_matrix(s, phi.copy(), Pi.copy())
    z_smooth = ref.smoothing(y, s, phi.copy(), Pi.copy())
    z_sharp = ref.sharpening(y, s, phi.copy(), Pi.copy())
    out = os.path.join(HERE, "pyref_proto.npz")
    np.savez_compressed(out, image=y.astype(np.uint8), sample_indices=np.asarray(s, dtype=np.uint32), K_A=K_A, K_B=K_B, phi=phi, Pi=Pi,
                        phi_perm=phi_perm, V=V, Pi_V=Pi_V, W_A=W_A, W_B=W_B, V_s=V_s, L_s=L_s, z_smooth=z_smooth, z_sharp=z_sharp)
    print("wrote", out, y.shape, "p =", len(s), "Pi", Pi[:3], "Pi_V", Pi_V[:3])
