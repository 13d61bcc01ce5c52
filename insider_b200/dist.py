"""One process per GPU. torch.distributed is used only as plumbing (rendezvous + broadcasting the 128-byte NCCL id);
the per-iteration exchange itself (all-reduce of the row-side sufficient statistics) is issued by libinsider_b200 on its
own NCCL communicator and CUDA stream.

Sharding: genes (columns) are split into contiguous blocks whose boundaries are multiples of 32 genes, rank r owns
gene_block(P, world, r). Everything indexed by rows (U, A_c, W, level ids, X) is replicated.
"""
from __future__ import annotations

import os

import numpy as np


def gene_block(P: int, world: int, rank: int):
    """(first gene, number of genes) owned by `rank` — must match split_genes() in csrc/lib.cu."""
    words = (P + 31) // 32
    q, r = divmod(words, world)
    w0 = rank * q + min(rank, r)
    w1 = w0 + q + (1 if rank < r else 0)
    j0 = min(P, w0 * 32)
    return j0, min(P, w1 * 32) - j0


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a fixed-size byte string from `src` over the default torch.distributed group (gloo or nccl)."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def make_context(device: int | None = None):
    """Creates the (possibly gene-sharded) library context for this process from the torchrun environment."""
    from . import _cabi

    rank, world, local = env_rank()
    device = local if device is None else device
    if world == 1:
        return _cabi.Context(device)
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("torch.distributed must be initialised before make_context() when WORLD_SIZE > 1")
    uid = _cabi.Context.nccl_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128, 0)
    return _cabi.Context(device, rank, world, uid)


def allreduce_stats_numpy(B_local: np.ndarray) -> np.ndarray:
    """CPU stand-in of the exchange step for tests: sums an array over ranks with the default process group."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(B_local))
    dist.all_reduce(t)
    return t.numpy()
