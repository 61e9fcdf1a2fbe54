"""Config 4 through gl_run_resident N times; prints the median of the per-kernel timers (for A/B runs, ncu, GLB200_PT_PROF=1)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipgl_b200 as gl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = gl.Context(0)
ctx.set_synthetic_image(3840, 2160, 1, 1234)
prm = gl.default_params(sampling=gl.RANDOM, sample_size=1000, seed=0)
acc = []
for _ in range(n):
    ctx.run_resident(prm)
    acc.append(ctx.stage_ms())
ctx.sync()
keys = acc[0].keys()
tail = acc[len(acc) // 2:] if len(acc) > 3 else acc
print({k: round(float(np.median([a[k] for a in tail])), 4) for k in keys})
ctx.close()
