"""world_size-2 `gloo` tests (CPU) of the N>1 path: the host plumbing in image-processing-graph-laplacian_b200/dist.py
and the sharding scheme itself -- pixels split into contiguous bands of image rows, the sampled block
replicated, and exactly two reductions crossing ranks (row sums D after the affinity stage, c = Phi^T y in the
filter) -- restated with the numpy oracle and checked against the unsharded oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _FakeCtx:
    """Stands in for ipgl_b200.Context (which needs a GPU): records what init_comm hands the library."""
    made = b"".join(bytes([i, 255 - i]) for i in range(64))

    def __init__(self, rank, world):
        self.rank, self.world, self.uid = rank, world, None

    def unique_id(self):
        assert self.rank == 0, "only rank 0 may create the NCCL id"
        return self.made

    def init_comm(self, uid):
        self.uid = uid


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    import ipgl_b200 as gl
    from ipgl_b200 import dist as gd
    from oracle import oracle_np as o

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert gd.env_rank_world() == (rank, world, rank)
        # --- plumbing -------------------------------------------------------------------------------------------
        ctx = _FakeCtx(rank, world)
        gd.init_comm(ctx, dist)
        assert ctx.uid == _FakeCtx.made                       # the id made on rank 0 reached every rank intact
        t = gd.max_over_ranks([1.0 + rank, 5.0 - rank], dist)
        assert t == [float(world), 5.0]                       # element-wise max over ranks
        # --- the sharded algorithm ----------------------------------------------------------------------------------
        W, H, p_req = 97, 61, 40
        img = o.synthetic_image(W, H, 1, seed=5)
        s = o.uniform_sampling(W, H, p_req).astype(np.int64)
        p, m = len(s), len(s) - 1
        r0, r1 = gd.band(H, rank, world)
        q = np.arange(r0 * W, r1 * W)                         # this rank's pixels, raster order
        y = img.reshape(-1).astype(np.float64)
        K_band = o.affinity_rows(img, s, q)                   # p x band pixels, sample pixels included
        D = torch.from_numpy(K_band.sum(axis=1))              # band partial of rowsum(K_A) + rowsum(K_B)
        dist.all_reduce(D)                                    # reduction 1 (p doubles)
        D = D.numpy()
        alpha = 1.0 / D.mean()
        L_A = alpha * (np.diag(D) - o.affinity_rows(img, s, s))      # replicated p x p block
        mu, U = o.smallest_eigenpairs(L_A, m)                 # replicated eigensolve
        phi = K_band.T @ ((-alpha) * U / mu[None, :])         # extrapolation: no communication
        mine = (s >= q[0]) & (s <= q[-1])
        phi[s[mine] - q[0]] = U[mine]                         # sample rows = Phi_A (nystroem.c:25-34)
        c = torch.from_numpy(phi.T @ y[q])
        dist.all_reduce(c)                                    # reduction 2 (m doubles)
        z_band = np.minimum(y[q] + 3.0 * (phi @ (mu * c.numpy())), 255.0)
        np.save(os.path.join(out_dir, f"z{rank}.npy"), z_band)
        np.save(os.path.join(out_dir, f"mu{rank}.npy"), mu)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_band_partition():
    from ipgl_b200 import dist as gd
    for H in (1, 2, 7, 61, 2160, 8192):
        for world in (1, 2, 3, 4, 8):
            bs = gd.bands(H, world)
            assert bs[0][0] == 0 and bs[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(bs, bs[1:]))          # contiguous, no gap, no overlap
            sizes = [b - a for a, b in bs]
            assert max(sizes) - min(sizes) <= 1                            # balanced to one row
    with pytest.raises(ValueError):
        gd.band(10, 2, 2)


def test_world2_gloo_sharded_path_matches_unsharded_oracle(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle_np as o
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    W, H, p_req = 97, 61, 40
    img = o.synthetic_image(W, H, 1, seed=5)
    ref = o.run_pipeline(img, o.uniform_sampling(W, H, p_req))
    z = np.concatenate([np.load(tmp_path / f"z{r}.npy") for r in range(world)]).reshape(H, W)
    mu0, mu1 = np.load(tmp_path / "mu0.npy"), np.load(tmp_path / "mu1.npy")
    assert np.array_equal(mu0, mu1)                                        # replicated solve, identical bits
    assert np.max(np.abs(mu0 - ref["mu"]) / ref["mu"]) < 1e-10
    assert np.linalg.norm(z - ref["z"]) / np.linalg.norm(ref["z"]) < 1e-10
    assert np.linalg.norm((z - img) - (ref["z"] - img)) / np.linalg.norm(ref["z"] - img) < 1e-8
