"""The N>1 host logic on CPU: world_size-2 gloo groups exercise the batch split and the output gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torchaudio_contrib_b200.distributed import all_gather_output, shard_batch, shard_range
        full_in = torch.arange(n_items * 6, dtype=torch.float32).reshape(n_items, 2, 3)
        mine = shard_batch(full_in)
        lo, hi = shard_range(n_items, rank, world)
        assert torch.equal(mine, full_in[lo:hi])
        local_out = mine * 2.0 + 1.0                       # stand-in for the per-rank kernel work
        gathered = all_gather_output(local_out, n_items)
        ok = torch.equal(gathered, full_in * 2.0 + 1.0)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, bool(ok), float(t[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 7])
def test_two_rank_split_and_gather(n_items):
    import build_native
    build_native.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(0, True, 2.0), (1, True, 2.0)]


def test_shard_range_partition():
    import build_native
    build_native.build()
    from torchaudio_contrib_b200.distributed import shard_range
    for n in (0, 1, 7, 64, 8192):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fd_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torchaudio_contrib_b200.distributed import MulticastGatheredOutput, _share_fd
        fd = -1
        if rank == 0:                                       # stands in for the exported multicast handle
            fd = os.memfd_create("tac_share_fd_test")
            os.write(fd, b"multicast-handle")
        got = _share_fd(fd, rank, world, None)
        os.lseek(got, 0, os.SEEK_SET) if rank == 0 else None
        data = os.pread(got, 64, 0)
        os.close(got)
        refused = False
        try:                                                # no CUDA device here: the collective constructor must not hang
            MulticastGatheredOutput((4, 1, 8, 16), "cuda:0")
        except Exception:
            refused = True
        q.put((rank, data == b"multicast-handle", refused))
    finally:
        dist.destroy_process_group()


def test_two_rank_fd_passing_for_the_multicast_object():
    """The host plumbing of the NVSwitch-multicast gather on CPU: rank 0's file descriptor (the exported multicast handle
    on a GPU box) reaches the other process over a unix socket (SCM_RIGHTS) and refers to the same open file."""
    import build_native
    build_native.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fd_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(0, True, True), (1, True, True)]
