"""GPU tuning harness: times the extrapolation GEMM (k_gemm_tcgen05) of the C4 workload under different options."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ipgl_b200 as gl

W, H, p = 3840, 2160, 1000
with gl.Context(0) as ctx:
    ctx.set_synthetic_image(W, H, 1, 1234)
    ctx.sampling(gl.RANDOM, p, 0)
    for cut in (1, 0):
        ctx.set_option("kb_cutoff", cut)
        K_A, K_B = ctx.affinity()
        L_A, L_B = ctx.laplacian(K_A, K_B)
        U, mu, mu_inv = ctx.eigensolve(L_A)
        for stages, pf in ((0, 2), (4, 0), (4, 2), (3, 0), (3, 2)):
            ctx.set_option("gemm_stages", stages)
            ctx.set_option("gemm_prefetch", pf)
            ts = []
            for _ in range(5):
                phi = ctx.nystroem(L_B, U, mu_inv)
                ts.append(ctx.stage_ms()["k_gemm"])
                phi.destroy()
            print(f"cutoff={cut} stages={stages} prefetch={pf}: k_gemm {np.median(ts):.3f} ms (min {min(ts):.3f})", flush=True)
        for m in (K_A, K_B, L_A, L_B, U, mu, mu_inv):
            m.destroy()
