"""One rank of the multi-GPU parity test (launched by tests/test_multi_gpu.py under torchrun, one process per GPU):
runs the whole path through the C ABI on this rank's band of image rows and saves the band's z, the eigenvalues
and the sample indices for the parent test to compare with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_fixture(out_dir, tag):
    """One of the committed golden cases (tests/golden/<tag>.npz, e.g. BASELINE.json config 3 = the lion image as RGB with the
    photometric affinity) through gl_run on this rank's band."""
    import torch
    import torch.distributed as dist

    import ipgl_b200 as gl
    from ipgl_b200 import dist as gd

    g = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"))
    img = g["image"]
    src = np.ascontiguousarray(np.repeat(img[:, :, None], 3, axis=2)) if int(g["rgb"]) else img
    rank, world, local = gd.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gl.Context(local, rank, world)
    gd.init_comm(ctx, dist, device="cuda")
    prm = gl.default_params(affinity=str(g["kind"]), sampling=gl.RANDOM if str(g["method"]) == "random" else gl.SPATIALLY_UNIFORM,
                            h_loc=float(g["h_loc"]), h_val=float(g["h_val"]), sample_size=int(g["p_req"]), seed=int(g["seed"]),
                            num_eigvals=int(g["m"]))
    z = np.zeros(src.shape, dtype=np.float32)
    r = ctx.run(src, prm, z_out=z)
    r0, r1 = ctx.band()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), z=z[r0:r1], mu=r["mu"], s=ctx.get_samples(), band=np.array([r0, r1]),
             outside=np.array([float(np.abs(z[:r0]).sum() + np.abs(z[r1:]).sum())]))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    if sys.argv[2].startswith("golden:"):
        return run_fixture(sys.argv[1], sys.argv[2][7:])
    out_dir, W, H, ch, p_req, sampling, gs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6], int(sys.argv[7])
    affinity = sys.argv[8] if len(sys.argv) > 8 else "bilateral"
    import torch
    import torch.distributed as dist

    import ipgl_b200 as gl
    from ipgl_b200 import dist as gd

    rank, world, local = gd.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gl.Context(local, rank, world)
    gd.init_comm(ctx, dist, device="cuda")
    ctx.set_synthetic_image(W, H, ch, 77)
    img = ctx.get_image()
    if affinity == "NLM":
        img = (img // 8 + 100).astype(np.uint8)              # low contrast: the h = 3 patch kernel stays alive
    # gl_run must bring everything it reads to the device itself: leave a DIFFERENT image resident, so that a rank reading
    # pixels it never uploaded (outside its band, other than the samples) would compute from the wrong data
    ctx.set_synthetic_image(W, H, ch, 999)
    prm = gl.default_params(sampling=sampling, sample_size=p_req, seed=3, gram_schmidt=gs, affinity=affinity)
    z = np.zeros(img.shape, dtype=np.float32)
    r = ctx.run(img, prm, z_out=z)
    r2 = ctx.run(img, prm, z_out=np.zeros_like(z))           # run twice: the comm must survive reuse
    assert np.array_equal(r["mu"], r2["mu"])
    r0, r1 = ctx.band()
    assert (r0, r1) == gd.band(H, rank, world)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), z=z[r0:r1], mu=r["mu"], s=ctx.get_samples(), band=np.array([r0, r1]),
             outside=np.array([float(np.abs(z[:r0]).sum() + np.abs(z[r1:]).sum())]))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
