"""Multi-GPU batch split / join: one process per GPU, clips sharded contiguously, no
collective inside the compute.  NCCL (through the C ABI, ``b2l_comm_*``) is used only to scatter a
device-resident batch from a root GPU and to gather the results back over NVLink; host-resident batches
are simply sliced per rank and uploaded directly (no funnel through one GPU).

The rendezvous (rank, world size, exchanging the 128-byte NCCL id) is the launcher's job: under
``torchrun`` use ``init_from_torch()``; any other mechanism can pass its own ``bcast_bytes`` callable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional, Tuple

import numpy as np

from . import _native as nat


def shard_range(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of clips for ``rank``: ceil(n/world) per rank, tail ranks may be short."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    per = -(-int(n_clips) // world)
    lo = min(n_clips, rank * per)
    return lo, min(n_clips, lo + per)


def split_batch(y: np.ndarray, rank: int, world: int) -> np.ndarray:
    """This rank's clips of a host batch ``(clips, ..., n)`` (sharded over the first axis)."""
    lo, hi = shard_range(y.shape[0], rank, world)
    return y[lo:hi]


def join_batches(parts) -> np.ndarray:
    """Inverse of split_batch for per-rank results gathered in rank order."""
    parts = [p for p in parts if p is not None and p.shape[0] > 0]
    return np.concatenate(parts, axis=0)


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


class Communicator:
    """NCCL communicator bound to a Context (one per process / GPU)."""

    def __init__(self, ctx: nat.Context, rank: int, world: int, bcast_bytes: Callable[[Optional[bytes]], bytes]):
        self.ctx, self.rank, self.world = ctx, rank, world
        L = nat.lib()
        buf = (C.c_char * 128)()
        if rank == 0:
            nat.check(L.b2l_comm_unique_id(C.cast(buf, C.c_void_p)))
            uid = bcast_bytes(bytes(buf))
        else:
            uid = bcast_bytes(None)
        C.memmove(buf, uid, 128)
        nat.check(L.b2l_comm_init(ctx.handle, C.cast(buf, C.c_void_p), rank, world))

    def broadcast(self, arr: nat.DeviceArray, root: int = 0):
        nat.check(nat.lib().b2l_comm_broadcast(self.ctx.handle, C.c_void_p(arr.ptr), arr.nbytes, root))

    def scatter(self, full: Optional[nat.DeviceArray], shard: nat.DeviceArray, root: int = 0):
        """Root's ``full`` (world equal shards, rank order) -> every rank's ``shard``."""
        ptr = full.ptr if (full is not None and self.rank == root) else 0
        nat.check(nat.lib().b2l_comm_scatter(self.ctx.handle, C.c_void_p(ptr), C.c_void_p(shard.ptr), shard.nbytes, root))

    def gather(self, shard: nat.DeviceArray, full: Optional[nat.DeviceArray], root: int = 0):
        ptr = full.ptr if (full is not None and self.rank == root) else 0
        nat.check(nat.lib().b2l_comm_gather(self.ctx.handle, C.c_void_p(shard.ptr), C.c_void_p(ptr), shard.nbytes, root))

    def barrier(self):
        nat.check(nat.lib().b2l_comm_barrier(self.ctx.handle))

    def close(self):
        nat.check(nat.lib().b2l_comm_destroy(self.ctx.handle))


def torch_bcast_bytes(payload: Optional[bytes]) -> bytes:
    """Broadcast helper for processes launched by torchrun (torch.distributed must be initialised)."""
    import torch.distributed as dist

    box = [payload]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def init_from_torch(ctx: Optional[nat.Context] = None) -> Communicator:
    import torch.distributed as dist

    rank, world, local = env_rank_world()
    if ctx is None:
        ctx = nat.default_context(local)
    if not dist.is_initialized():
        dist.init_process_group(backend="gloo")
    return Communicator(ctx, rank, world, torch_bcast_bytes)
