"""GPU diagnostic: runs the extrapolation GEMM with the tcgen05 kernel and with the CUDA-core checker on
the same inputs and reports where they differ (rows mod 128 / cols mod 64 patterns point at descriptor,
swizzle or TMEM-lane mistakes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipgl_b200 as gl
from oracle import oracle_np as o


def phi_for(ctx, img, p, impl, kind="bilateral"):
    ctx.set_option("gemm", impl)
    ctx.set_image(img)
    ctx.sampling(gl.SPATIALLY_UNIFORM, p)
    K_A, K_B = ctx.affinity(kind)
    L_A, L_B = ctx.laplacian(K_A, K_B)
    U, mu, mu_inv = ctx.eigensolve(L_A, -1)
    phi = ctx.nystroem(L_B, U, mu_inv)
    ctx.sync()
    return phi.download(), ctx.stage_ms()["nystroem"]


def main():
    with gl.Context(0) as ctx:
        for (W, H, p) in ((192, 128, 60), (320, 200, 300), (640, 360, 1000)):
            img = o.synthetic_image(W, H, 1, seed=3)
            a, ta = phi_for(ctx, img, p, "simple")
            try:
                b, tb = phi_for(ctx, img, p, "tcgen05")
            except gl.GLError as e:
                print(f"{W}x{H} p={p}: tcgen05 FAILED: {e}")
                return 1
            d = np.abs(a - b)
            scale = np.abs(a).max()
            print(f"{W}x{H} p={p}: phi {a.shape} max|simple|={scale:.3e} max diff={d.max():.3e} "
                  f"(rel {d.max() / scale:.2e}) simple {ta:.3f} ms tcgen05 {tb:.3f} ms")
            if d.max() > 1e-2 * scale:
                bad = d > 1e-2 * scale
                rows, cols = np.nonzero(bad)
                print("  bad entries:", bad.sum(), "of", bad.size)
                print("  rows mod 128 histogram (nonzero bins):", np.nonzero(np.bincount(rows % 128, minlength=128))[0][:40])
                print("  cols mod 64 histogram (nonzero bins):", np.nonzero(np.bincount(cols % 64, minlength=64))[0][:64])
                print("  first bad:", rows[:8], cols[:8])
                print("  simple :", a[rows[0], max(0, cols[0] - 2):cols[0] + 6])
                print("  tcgen05:", b[rows[0], max(0, cols[0] - 2):cols[0] + 6])
    return 0


if __name__ == "__main__":
    sys.exit(main())
