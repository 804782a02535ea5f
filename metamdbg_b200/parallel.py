"""Multi-GPU host plumbing (one process per GPU): record sharding and the key-owner
function of the count-table merge.  The exchange itself is inside libmdbg_b200
(`mdbg_count_merge`, NCCL all-to-all); torch.distributed is only used to hand the
128-byte NCCL unique id to every rank.

Reference analogue: the `hash128 % P` partitioning of KminmerCounter::partitionKminmer
(src/graph/CreateMdbg.hpp:3714-3724) -- P is not observable in the outputs, so the
GPU form uses the high word of Murmur h1 instead."""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous record range of one rank."""
    return n_items * rank // world, n_items * (rank + 1) // world


def owner_of(h1: np.ndarray, world: int) -> np.ndarray:
    """Owner rank of a k-min-mer from the high 64 bits (Murmur h1) of its hash128:
    ((h1 >> 32) * world) >> 32  -- same arithmetic as `owner_of` in csrc/engine.cuh."""
    hi = (np.asarray(h1, dtype=np.uint64) >> np.uint64(32))
    return ((hi * np.uint64(world)) >> np.uint64(32)).astype(np.int64)


def init_engine_comm(engine, rank: int, world: int, device=None) -> None:
    """Create the engine's NCCL communicator; the unique id travels through torch.distributed."""
    import torch
    import torch.distributed as dist
    from .engine import Engine

    backend = dist.get_backend()
    dev = device if (device is not None and backend == "nccl") else "cpu"
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(Engine.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    engine.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
