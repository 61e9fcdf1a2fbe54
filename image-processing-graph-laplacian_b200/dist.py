"""One-process-per-GPU plumbing around the C ABI: which image rows a rank owns, how the NCCL id of
libglcuda's own communicator reaches every rank, and how per-rank device times become one number.

The reference is SPMD over MPI ranks (hpc/image_processing.c:30-38); here the ranks are `torch.distributed`
processes (torchrun), torch is used ONLY as the rendezvous/broadcast channel -- the three small reductions of the
path (row sums D, optional Gram block, c = Phi^T y) run inside libglcuda.so on its own NCCL communicator.
Everything in this file also works on the `gloo` backend with CPU tensors, which is how it is tested without GPUs.
"""
from __future__ import annotations

import os


def band(height: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous band [row0, row1) of whole image rows owned by `rank` -- the same arithmetic as
    set_image_geometry() in csrc/api.cu, so host code can size per-rank buffers without asking the device."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    return height * rank // world, height * (rank + 1) // world


def bands(height: int, world: int) -> list[tuple[int, int]]:
    return [band(height, r, world) for r in range(world)]


def env_rank_world() -> tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def broadcast_bytes(payload: bytes | None, nbytes: int, dist, device="cpu", src: int = 0) -> bytes:
    """Rank `src` passes `payload` (exactly nbytes); every rank returns it."""
    import torch
    if dist.get_rank() == src:
        if payload is None or len(payload) != nbytes:
            raise ValueError("source rank must supply exactly nbytes")
        t = torch.tensor(list(payload), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def init_comm(ctx, dist, device="cpu") -> None:
    """Join this rank's Context to the library's NCCL communicator: rank 0 makes the 128-byte id
    (gl_comm_unique_id), torch.distributed carries it, every rank calls gl_comm_init."""
    if ctx.world == 1:
        return
    uid = ctx.unique_id() if ctx.rank == 0 else None
    ctx.init_comm(broadcast_bytes(uid, 128, dist, device))


def max_over_ranks(values, dist=None, device="cpu") -> list[float]:
    """Element-wise maximum over ranks of a list of per-rank device times (the multi-GPU timing rule)."""
    vals = [float(v) for v in values]
    if dist is None:
        return vals
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu().tolist()]


def sum_over_ranks(values, dist=None, device="cpu") -> list[float]:
    """Element-wise sum over ranks (fp64): per-band error sums of a parity check become whole-image sums."""
    vals = [float(v) for v in values]
    if dist is None:
        return vals
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.cpu().tolist()]
