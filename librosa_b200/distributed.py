"""Multi-GPU batch split / join: one process per GPU, clips sharded contiguously, no
collective inside the compute.  NCCL (through the C ABI, ``b2l_comm_*``) is used only to scatter a
device-resident batch from a root GPU and to gather the results back over NVLink; host-resident batches
are simply sliced per rank and uploaded directly (no funnel through one GPU).

The rendezvous (rank, world size, exchanging the 128-byte NCCL id) needs no framework: ``init_from_env()``
reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT (what ``torchrun`` or any launcher exports)
and rank 0 hands the id to the other ranks over a plain TCP socket (``tcp_bcast_bytes``).  Any other
mechanism can pass its own ``bcast_bytes`` callable to ``Communicator``.  No PyTorch import anywhere.
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


def rendezvous_endpoint() -> Tuple[str, int]:
    """(host, port) of the id exchange: B2L_RDZV_PORT, else MASTER_PORT + 23 (the launcher's own store keeps
    MASTER_PORT itself), on MASTER_ADDR (default 127.0.0.1)."""
    host = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = os.environ.get("B2L_RDZV_PORT")
    if port is None:
        port = int(os.environ.get("MASTER_PORT", "29500")) + 23
    return host, int(port)


def tcp_bcast_bytes(payload: Optional[bytes], rank: int, world: int, *, host: Optional[str] = None,
                    port: Optional[int] = None, timeout: float = 120.0) -> bytes:
    """Rank 0 serves ``payload`` to the world - 1 other ranks over TCP; everyone returns it.

    The message is length-prefixed; clients retry the connection until the server is up (ranks start in any
    order) and fail after ``timeout`` seconds."""
    import socket
    import struct
    import time

    if world <= 1:
        return payload
    ep_host, ep_port = rendezvous_endpoint()
    host = ep_host if host is None else host
    port = ep_port if port is None else port
    if rank == 0:
        if payload is None:
            raise ValueError("rank 0 must supply the payload")
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind(("" if host not in ("127.0.0.1", "localhost") else "127.0.0.1", port))
        srv.listen(world)
        srv.settimeout(timeout)
        try:
            for _ in range(world - 1):
                conn, _addr = srv.accept()
                with conn:
                    conn.sendall(struct.pack("<I", len(payload)) + payload)
        finally:
            srv.close()
        return payload
    deadline = time.monotonic() + timeout
    while True:
        try:
            with socket.create_connection((host, port), timeout=5.0) as conn:
                conn.settimeout(timeout)
                head = _recv_exact(conn, 4)
                (size,) = struct.unpack("<I", head)
                return _recv_exact(conn, size)
        except (ConnectionRefusedError, ConnectionResetError, socket.timeout, OSError):
            if time.monotonic() > deadline:
                raise TimeoutError(f"rank {rank}: no NCCL id from {host}:{port} within {timeout} s")
            time.sleep(0.05)


def _recv_exact(conn, size: int) -> bytes:
    buf = bytearray()
    while len(buf) < size:
        chunk = conn.recv(size - len(buf))
        if not chunk:
            raise ConnectionResetError("peer closed during the id exchange")
        buf += chunk
    return bytes(buf)


def init_from_env(ctx: Optional[nat.Context] = None) -> Communicator:
    """One process per GPU under any launcher that exports RANK / WORLD_SIZE / LOCAL_RANK (torchrun, mpirun
    wrappers, a shell loop): create the NCCL communicator of this rank's context."""
    rank, world, local = env_rank_world()
    if ctx is None:
        ctx = nat.default_context(local)
    return Communicator(ctx, rank, world, lambda payload: tcp_bcast_bytes(payload, rank, world))
