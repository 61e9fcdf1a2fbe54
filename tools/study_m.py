"""Config 4: how the filtered image depends on the number of eigenpairs m (-num_eigvals), and what the two eigensolvers cost for a
partial spectrum.  Prints one line per m: rel. L2 of z - y against the m = p - 1 result, eigen stage ms with
the block Jacobi (all pairs, first m kept) and with the reference's inverse subspace iteration (gl_inverse_iteration, epsilon 1e-3)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipgl_b200 as gl

ctx = gl.Context(0)
W, H = 3840, 2160
ctx.set_synthetic_image(W, H, 1, 1234)
y = ctx.get_image().astype(np.float64)
ctx.sampling(gl.RANDOM, 1000, seed=0)
K_A, K_B = ctx.affinity()
L_A, L_B = ctx.laplacian(K_A, K_B)
p = 1000


def run(m, solver):
    if solver == "jacobi":
        U, mu, mu_inv = ctx.eigensolve(L_A, m)
        it, res = 0, 0.0
    else:
        U, mu, mu_inv, it, res = ctx.inverse_iteration(L_A, m, 1, 1e-3, 500)
    ms = ctx.stage_ms()["eigen"]
    z = ctx.filter(ctx.nystroem(L_B, U, mu_inv), mu).astype(np.float64)
    return z, ms, it, res, mu.download()


z_full, ms_full, _, _, mu_full = run(p - 1, "jacobi")
den = np.linalg.norm(z_full - y)
print(f"m={p - 1}: jacobi eigen {ms_full:.3f} ms; |z - y| / |y| = {den / np.linalg.norm(y):.3e}")
for m in (10, 20, 50, 100, 200, 500, 900):
    zj, msj, _, _, _ = run(m, "jacobi")
    line = f"m={m}: err_dz(jacobi, first m)={np.linalg.norm(zj - z_full) / den:.3e} eigen {msj:.3f} ms"
    if m <= 200:
        try:
            zi, msi, it, res, mui = run(m, "inverse")
            emu = float(np.max(np.abs(np.sort(mui) - mu_full[:m]) / mu_full[:m]))
            line += f" | inverse iteration: {it} steps, residual {res:.2e}, eigen {msi:.3f} ms, err_mu {emu:.2e}, err_dz {np.linalg.norm(zi - z_full) / den:.3e}"
        except gl.GLError as e:
            line += f" | inverse iteration: {e}"
    print(line, flush=True)
ctx.close()
