"""CPU: the N>1 host logic with world_size=2 over gloo — contiguous clip sharding, the per-rank
pipeline, the rank-ordered join, and the max-over-ranks timing reduction bench.py uses.  The per-rank
compute is the oracle here (no GPU in this container); on the GPU box the same split / join code drives
the CUDA path (tests/test_gpu_parity.py::test_two_rank_split_join_on_gpu)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    import signals
    from librosa_b200 import distributed as D
    from oracle import ref_np as O

    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    assert D.env_rank_world() == (rank, world, rank)
    batch = signals.make("A", (5, 6000), seed=100)          # 5 clips over 2 ranks: 3 + 2
    mine = D.split_batch(batch, rank, world)
    lo, hi = D.shard_range(batch.shape[0], rank, world)
    assert mine.shape[0] == hi - lo
    part = O.melspectrogram(y=mine, sr=16000, n_fft=1024, hop_length=256)
    gathered = [None] * world
    dist.all_gather_object(gathered, part)
    full = D.join_batches(gathered)
    # broadcast of an opaque id (what Communicator does with the NCCL unique id): the product's own TCP
    # rendezvous, no torch involved
    uid = D.tcp_bcast_bytes(bytes(range(128)) if rank == 0 else None, rank, world)
    assert uid == bytes(range(128))
    assert D.rendezvous_endpoint() == ("127.0.0.1", port + 23)
    # max-over-ranks of a per-rank time, as bench.py reports it
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == 10.0 + world - 1
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_split_join(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import signals
    from oracle import ref_np as O

    batch = signals.make("A", (5, 6000), seed=100)
    want = O.melspectrogram(y=batch, sr=16000, n_fft=1024, hop_length=256)
    got = np.load(os.path.join(str(tmp_path), "full.npy"))
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9)


def _tcp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from librosa_b200 import distributed as D

    got = D.tcp_bcast_bytes(os.urandom(128) if rank == 0 else None, rank, world)
    q.put((rank, got, "torch" in sys.modules))


def test_tcp_rendezvous_three_ranks_without_torch():
    """The id exchange of librosa_b200.distributed works for any launcher and never imports torch."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tcp_worker, args=(r, 3, port, q)) for r in (2, 1, 0)]   # server starts last
    for p in procs:
        p.start()
    got = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    payloads = {g[1] for g in got}
    assert len(payloads) == 1 and len(next(iter(payloads))) == 128
    assert not any(g[2] for g in got), "librosa_b200.distributed pulled in torch"


def test_shard_ranges_cover_the_batch():
    sys.path.insert(0, ROOT)
    from librosa_b200 import distributed as D

    for n in (0, 1, 5, 8, 1024, 8191):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
