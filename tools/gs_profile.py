import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipgl_b200 as gl
W, H, p = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (400, 300, 500)
with gl.Context(0) as ctx:
    ctx.set_synthetic_image(W, H, 1, 1234)
    ctx.sampling(gl.RANDOM, p, 0)
    K_A, K_B = ctx.affinity()
    L_A, L_B = ctx.laplacian(K_A, K_B)
    U, mu, mu_inv = ctx.eigensolve(L_A)
    for _ in range(2):
        phi = ctx.nystroem(L_B, U, mu_inv)
        ctx.orthonormalise(phi)
        print("gs ms", ctx.stage_ms()["gram_schmidt"])
        phi.destroy()
