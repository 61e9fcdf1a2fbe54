"""GPU tuning harness for the eigensolver on the L_A of the C4 workload (built on the device, downloaded, solved with numpy
for reference)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ipgl_b200 as gl

os.environ["GLB200_VERBOSE"] = "1"
with gl.Context(0) as ctx:
    ctx.set_option("verbose", 1)
    ctx.set_synthetic_image(3840, 2160, 1, 1234)
    ctx.sampling(gl.RANDOM, 1000, 0)
    K_A, K_B = ctx.affinity()
    L_A, L_B = ctx.laplacian(K_A, K_B)
    A = L_A.download()
    w, V = np.linalg.eigh(A)
    p = A.shape[0]
    L = ctx.upload(gl.MAT_KA, A)
    for tol, inner in ((2e-6, 0), (2e-6, 1), (2e-6, 2), (1e-5, 0), (1e-5, 1), (5e-5, 1), (5e-5, 0)):
        ctx.set_option("jacobi_tol", tol)
        ctx.set_option("jacobi_inner", inner)
        ts, tk = [], []
        for _ in range(4):
            U, mu, mui = ctx.eigensolve(L, p - 1)
            s = ctx.stage_ms()
            ts.append(s["eigen"]); tk.append(s["k_jacobi"])
            if _ < 3:
                U.destroy(); mu.destroy(); mui.destroy()
        got = mu.download(); Ug = U.download()
        err_mu = np.max(np.abs(got - w[:p - 1]) / w[:p - 1])
        orth = np.max(np.abs(Ug.T @ Ug - np.eye(p - 1)))
        res = np.max(np.abs(A @ Ug - Ug * got)) / np.max(np.abs(w))
        # what the filter sees: the operator U diag(mu) U^T against the exact one
        Fg = (Ug * got) @ Ug.T
        Fe = (V[:, :p - 1] * w[:p - 1]) @ V[:, :p - 1].T
        err_op = np.linalg.norm(Fg - Fe) / np.linalg.norm(Fe)
        print(f"tol={tol:g} inner={inner}: eigen {np.median(ts):.3f} ms (k_jacobi {np.median(tk):.3f}) err_mu={err_mu:.2e} orth={orth:.2e} res={res:.2e} op={err_op:.2e}", flush=True)
        U.destroy(); mu.destroy(); mui.destroy()
