"""Config 4 through gl_run_resident with option keep_phi=0 a few times (the kernel of interest for ncu: k_gemm_tcgen05)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipgl_b200 as gl

ctx = gl.Context(0)
ctx.set_synthetic_image(3840, 2160, 1, 1234)
ctx.set_option("keep_phi", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
prm = gl.default_params(sampling=gl.RANDOM, sample_size=1000, seed=0)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    ctx.run_resident(prm)
ctx.sync()
print({k: round(v, 4) for k, v in ctx.stage_ms().items() if v})
ctx.close()
