"""GPU check of the tcgen05 Gram kernel (MN-major operands) against the CUDA-core tiles, and timing at C4."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ipgl_b200 as gl
from oracle import oracle_np as o

def phi_for(ctx, W, H, p, seed=0):
    ctx.set_synthetic_image(W, H, 1, 1234)
    ctx.sampling(gl.RANDOM, p, seed)
    K_A, K_B = ctx.affinity()
    L_A, L_B = ctx.laplacian(K_A, K_B)
    U, mu, mu_inv = ctx.eigensolve(L_A)
    return L_B, U, mu, mu_inv

with gl.Context(0) as ctx:
    for (W, H, p) in ((320, 200, 200), (400, 300, 500)):
        L_B, U, mu, mu_inv = phi_for(ctx, W, H, p)
        res = {}
        for name, opts in (("simple", {"gram": "simple"}), ("tc", {"gram": "tcgen05", "gram_lbo": 8192}), ("tc_swapped", {"gram": "tcgen05", "gram_lbo": 1024})):
            for k, v in opts.items():
                ctx.set_option(k, v)
            phi = ctx.nystroem(L_B, U, mu_inv)
            try:
                norms = ctx.orthonormalise(phi)
                Q = phi.download()
                res[name] = (norms, Q)
                print(f"{W}x{H} p={p} {name}: orth err {np.max(np.abs(Q.T @ Q - np.eye(Q.shape[1]))):.2e} gs ms {ctx.stage_ms()['gram_schmidt']:.3f}", flush=True)
            except gl.GLError as e:
                print(f"{W}x{H} p={p} {name}: FAILED {e}", flush=True)
            phi.destroy()
        if "tc" in res:
            print("   tc vs simple: norms", np.max(np.abs(res["tc"][0] - res["simple"][0]) / res["simple"][0]), "Q", np.max(np.abs(res["tc"][1] - res["simple"][1])))
    ctx.set_option("gram", "tcgen05"); ctx.set_option("gram_lbo", 8192)
    L_B, U, mu, mu_inv = phi_for(ctx, 3840, 2160, 1000)
    for name in ("tcgen05", "simple"):
        ctx.set_option("gram", name)
        phi = ctx.nystroem(L_B, U, mu_inv)
        ctx.orthonormalise(phi)
        print(f"C4 gram={name}: gram_schmidt stage {ctx.stage_ms()['gram_schmidt']:.2f} ms", flush=True)
        phi.destroy()
