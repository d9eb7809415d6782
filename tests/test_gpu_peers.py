"""Two ranks on two GPUs of one box: the mel kernel's epilogue stores every frame into both ranks' full
output (tac_melspec_banded_peers_f32) and the result equals the oracle on the whole batch.
Skipped on a single-GPU box; the handle exchange runs over gloo (no NCCL needed)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, layout, to_db, q, kind="peer"):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torchaudio_contrib_b200 as tac
        from torchaudio_contrib_b200.distributed import MulticastGatheredOutput, PeerGatheredOutput, shard_batch, multicast_supported
        from oracle import ref_chain
        torch.cuda.set_device(rank)
        if kind == "multicast" and not multicast_supported(rank):
            q.put((rank, "unsupported"))
            return
        dev = torch.device("cuda", rank)
        torch.manual_seed(7)
        x_all = torch.randn(2 * world, 2, 24000)                     # the whole batch, known to every rank
        fb = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=16000).get_filterbank()
        mine = shard_batch(x_all, rank, world).contiguous().to(dev)
        prep = tac.PreparedMelspectrogram(mine.shape, dev, fb, 2048, 512, to_db=to_db, layout=layout)
        cls = MulticastGatheredOutput if kind == "multicast" else PeerGatheredOutput
        buf = cls((x_all.shape[0],) + prep.out_shape[1:], dev)
        ok = True
        for rounds in range(3):                                      # epochs advance, buffers are re-used
            buf.tensor.fill_(float("nan"))
            torch.cuda.synchronize(dev)
            dist.barrier()
            got = prep.gather_into(mine, buf)
            buf.barrier()
            buf.check()
            want = ref_chain.melspectrogram(x_all, 128, 16000, to_db=to_db, fft_length=2048, hop_length=512)
            g = got.cpu()
            ok = ok and g.shape == want.shape and bool(torch.isfinite(g).all())
            if to_db:
                ok = ok and float((g - want).abs().max()) < 1e-3
            else:
                ok = ok and float(((g - want).abs() / want.abs().clamp_min(1e-30)).max()) < 1e-4     # north_star: 1e-4 rel
            local = prep(mine, prep.empty_output())                  # same bits as the single-GPU call
            lo = rank * mine.shape[0]
            ok = ok and bool(torch.equal(got[lo:lo + mine.shape[0]], local))
            buf.barrier()                                            # everyone has read before the next round overwrites
        buf.check()
        buf.close()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("layout,to_db", [("reference", False), ("contiguous", True)])
def test_peer_gather_two_gpus(layout, to_db):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs of one box")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, layout, to_db, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    alive = [p for p in procs if p.is_alive()]
    for p in alive:
        p.kill()
    assert not alive, "peer-gather worker hung"
    results = sorted(q.get(timeout=5) for _ in range(2))
    assert results == [(0, True), (1, True)]


@pytest.mark.parametrize("layout,to_db", [("reference", False), ("contiguous", True)])
def test_multicast_gather_two_gpus(layout, to_db):
    """The same gather with ONE store per value to a CUDA multicast address (csrc/multicast.cu): both ranks' replicas hold
    the whole batch, equal to the oracle and bit-identical to the single-GPU call."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs of one box")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, layout, to_db, q, "multicast")) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    alive = [p for p in procs if p.is_alive()]
    for p in alive:
        p.kill()
    assert not alive, "multicast-gather worker hung"
    results = sorted(q.get(timeout=5) for _ in range(2))
    if results == [(0, "unsupported"), (1, "unsupported")]:
        pytest.skip("devices cannot join a multicast object")
    assert results == [(0, True), (1, True)]


def test_multicast_gather_single_rank_is_refused():
    """The driver refuses a multicast object with one device (cuMulticastCreate: invalid argument): the class says so
    up front instead of surfacing a driver error."""
    import torch.distributed as dist
    from torchaudio_contrib_b200.distributed import MulticastGatheredOutput
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        with pytest.raises(NotImplementedError):
            MulticastGatheredOutput((2, 1, 10, 128), torch.device("cuda", 0))
    finally:
        dist.destroy_process_group()


def test_peer_gather_single_rank():
    """world_size 1 on one GPU: the peer allocation, the PEERS variant of the mel kernel and the flag barrier run
    (this is what a single-GPU box can exercise of the N > 1 path); result identical to the plain call."""
    import torch.distributed as dist
    import torchaudio_contrib_b200 as tac
    from torchaudio_contrib_b200.distributed import PeerGatheredOutput
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        torch.manual_seed(11)
        dev = torch.device("cuda", 0)
        x = torch.randn(3, 2, 20000, device=dev)
        fb = tac.MelFilterbank(num_freqs=1025, num_mels=128, sample_rate=16000).get_filterbank()
        for layout in ("reference", "contiguous"):
            prep = tac.PreparedMelspectrogram(x.shape, dev, fb, 2048, 512, to_db=True, layout=layout)
            buf = PeerGatheredOutput(prep.out_shape, dev)
            lib = tac._cabi.lib()
            n0 = lib.tac_launch_count()
            got = prep.gather_into(x, buf)
            buf.barrier()
            buf.check()
            assert lib.tac_launch_count() - n0 == 2                 # the mel kernel + the barrier kernel
            want = prep(x, prep.empty_output())
            assert got.shape == want.shape and got.stride() == want.stride()
            assert torch.equal(got, want)
            with pytest.raises(RuntimeError):
                prep.gather_into(x, buf, item_offset=1)              # would run past the gathered batch
            buf.close()
    finally:
        dist.destroy_process_group()
