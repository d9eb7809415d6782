"""Multi-GPU plumbing for the batch-sharded path (BASELINE config 4): one process per GPU, the batch
split contiguously across ranks, no data-path collective; an optional all-gather of the output.

Every (batch, channel) sequence is independent (reference functional.py:89-91 flattens all leading
dims into one batch), so rank r of W simply owns a contiguous range of batch items.  The only
collective is the optional `all_gather_output`, NCCL over NVLink on GPUs (gloo in the CPU tests).

`PeerGatheredOutput` is the same gather done by the mel kernel itself: every rank's
full output buffer is mapped into every other rank (CUDA IPC), `PreparedMelspectrogram.gather_into`
stores each frame's bands into all of them over NVLink while it computes, and one flag barrier
(`tac_peer_barrier`) replaces the collective.  torch.distributed only carries the 64-byte handles.

`MulticastGatheredOutput` is the same with ONE store per value: the ranks' buffers are the replicas
of a CUDA multicast object and the kernel stores to the multicast address; the NVSwitch delivers
the store to every GPU (1x NVLink egress instead of world-1 x).
"""
import ctypes
import os
import socket
import time

import torch
import torch.distributed as dist

from . import _cabi

__all__ = ["shard_range", "shard_batch", "all_gather_output", "PeerGatheredOutput", "MulticastGatheredOutput",
           "multicast_supported"]


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) of `n_items` owned by `rank`; the first `n_items % world` ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x, rank=None, world=None):
    """This rank's slice of a batch-first tensor (a view, no copy)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(x.size(0), rank, world)
    return x[lo:hi]


def all_gather_output(local_out, n_items, group=None):
    """Gather per-rank outputs (batch-first, sharded by `shard_range(n_items, ...)`) into the full
    `(n_items, ...)` tensor on every rank.  Equal shards use one `all_gather_into_tensor`; ragged
    shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [hi - lo for lo, hi in (shard_range(n_items, r, world) for r in range(world))]
    if local_out.size(0) != sizes[rank]:
        raise ValueError("rank %d holds %d items, expected %d" % (rank, local_out.size(0), sizes[rank]))
    local_out = local_out.contiguous()
    tail = tuple(local_out.shape[1:])
    if len(set(sizes)) == 1:
        full = torch.empty((n_items,) + tail, dtype=local_out.dtype, device=local_out.device)
        dist.all_gather_into_tensor(full, local_out, group=group)
        return full
    biggest = max(sizes)
    padded = torch.zeros((biggest,) + tail, dtype=local_out.dtype, device=local_out.device)
    padded[:sizes[rank]] = local_out
    buf = torch.empty((world * biggest,) + tail, dtype=local_out.dtype, device=local_out.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * biggest:r * biggest + sizes[r]] for r in range(world)], dim=0)


class _PeerBlock(object):
    """Owner of one tac_peer_alloc allocation, exposed to torch through __cuda_array_interface__."""

    def __init__(self, shape, device):
        self.shape = tuple(int(d) for d in shape)
        n = 1
        for d in self.shape:
            n *= d
        self.device = device
        self.base = ctypes.c_void_p()
        self.handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(device):
            _cabi.check(_cabi.lib().tac_peer_alloc(4 * n, ctypes.byref(self.base), ctypes.cast(self.handle, ctypes.c_void_p)))
        self.payload = self.base.value + 128                 # TAC_PEER_HEADER_BYTES
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4", "data": (self.payload, False),
                                         "version": 2, "strides": None}

    def __del__(self):
        base, self.base = getattr(self, "base", None), None
        if base is not None and base.value:
            try:
                with torch.cuda.device(self.device):
                    _cabi.lib().tac_peer_free(base)
            except Exception:                                 # interpreter shutdown
                pass


class PeerGatheredOutput(object):
    """The full `(n_items, ...)` output of a batch-sharded call, resident on every rank and written by
    every rank's kernel directly (no collective on the data path).

        buf = PeerGatheredOutput((B, C, frames, num_bands), device)        # collective: all ranks construct it
        prep.gather_into(x_local, buf)                                     # kernel stores to all ranks
        buf.barrier()                                                      # every rank's frames have landed
        full = buf.tensor                                                  # (B, C, frames, num_bands) on this rank

    `barrier()` is stream-ordered (a kernel that publishes a flag to every peer and waits for theirs).
    Call it a second time after consuming `tensor` and before the next `gather_into` if ranks run at
    different paces (a fast rank would otherwise overwrite what a slow rank is still reading), or
    alternate between two buffers.  All ranks must be processes on GPUs of one box."""

    def __init__(self, shape, device, group=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerGatheredOutput needs an initialised torch.distributed process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise NotImplementedError("PeerGatheredOutput: %d ranks; peer stores cover the 8 GPUs of one box" % self.world)
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._own = _PeerBlock(shape, self.device)
        self.tensor = torch.as_tensor(self._own, device=self.device)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(self._own.handle), group=group)
        self._bases = []
        lib = _cabi.lib()
        failure = None
        try:
            with torch.cuda.device(self.device):
                for r, h in enumerate(handles):
                    if r == self.rank:
                        self._bases.append(self._own.base.value)
                        continue
                    mapped = ctypes.c_void_p()
                    buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                    _cabi.check(lib.tac_peer_open(ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(mapped)))
                    self._bases.append(mapped.value)
        except Exception as exc:                               # tell the others instead of leaving them in a collective
            failure = "rank %d: %s" % (self.rank, exc)
        failures = [None] * self.world
        dist.all_gather_object(failures, failure, group=group)   # also: every rank has mapped every buffer
        failures = [f for f in failures if f]
        if failures:
            self._unmap()                                      # whatever this rank did map before a peer failed
            self._own = None
            raise RuntimeError("PeerGatheredOutput: mapping peer memory failed (%s)" % "; ".join(failures))
        self.base_array = (ctypes.c_void_p * self.world)(*self._bases)
        self.payload_array = (ctypes.c_void_p * self.world)(*[b + 128 for b in self._bases])
        self.epoch = 0

    def barrier(self, timeout_s=20.0):
        """Enqueue the flag barrier on the current stream: after it, every rank's stores issued before its own
        `barrier()` call are visible in this rank's `tensor`."""
        self.epoch += 1
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().tac_peer_barrier(ctypes.cast(self.base_array, ctypes.c_void_p), self.world, self.rank,
                                                     self.epoch, float(timeout_s), _cabi.stream_ptr(self.device)))

    def wait(self, timeout_s=20.0):
        """`barrier()`, then block the host until it has run and raise if it gave up on a peer: the call to use when the
        gathered tensor is about to be consumed on the host or handed on (a timed-out barrier otherwise leaves a
        partially filled `tensor` with nothing but `check()` to say so)."""
        self.barrier(timeout_s)
        self.check()
        return self.tensor

    def _unmap(self):
        """Unmap the peers' buffers this rank has opened (no collective)."""
        bases, self._bases = self._bases, None
        if not bases:
            return
        lib = _cabi.lib()
        with torch.cuda.device(self.device):
            for r, b in enumerate(bases):
                if r != self.rank and b:
                    lib.tac_peer_close(ctypes.c_void_p(b))

    def __del__(self):
        try:                                                   # a buffer dropped without close(): at least unmap the peers
            if getattr(self, "_bases", None):
                torch.cuda.synchronize(self.device)
                self._unmap()
        except Exception:                                      # interpreter shutdown
            pass

    def check(self):
        """Synchronise and raise if a barrier gave up waiting for a peer (mandatory after `barrier()` before the result
        is trusted; `wait()` does both)."""
        flag = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().tac_peer_timed_out(ctypes.c_void_p(self._own.base.value), ctypes.byref(flag)))
        if flag.value:
            raise RuntimeError("PeerGatheredOutput: a barrier timed out waiting for a peer rank")

    def close(self):
        """Collective: unmap the peers' buffers, then release this rank's (after every peer has unmapped it)."""
        if self._bases is None:
            return
        torch.cuda.synchronize(self.device)
        lib = _cabi.lib()
        with torch.cuda.device(self.device):
            for r, b in enumerate(self._bases):
                if r != self.rank:
                    _cabi.check(lib.tac_peer_close(ctypes.c_void_p(b)))
        self._bases = None
        dist.barrier(group=self.group)
        self.tensor = None
        self._own = None


def multicast_supported(device=None):
    """True when the device can join a CUDA multicast object (NVSwitch box, driver support)."""
    flag = ctypes.c_int(0)
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        _cabi.check(_cabi.lib().tac_mc_supported(ctypes.byref(flag)))
    return bool(flag.value)


def _share_fd(fd, rank, world, group):
    """Rank 0 hands a file descriptor to every other rank (processes of one box): unix socket + SCM_RIGHTS; the socket
    path travels through torch.distributed.  Returns this rank's descriptor (rank 0: the one it passed in)."""
    if world == 1:
        return fd
    path = [None]
    server = None
    if rank == 0:
        path[0] = "/tmp/tac_mc_%d_%d.sock" % (os.getpid(), int(time.time() * 1e6) & 0xffffff)
        if os.path.exists(path[0]):
            os.unlink(path[0])
        server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        server.bind(path[0])
        server.listen(world)
    dist.broadcast_object_list(path, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    try:
        if rank == 0:
            server.settimeout(60.0)
            for _ in range(world - 1):
                conn, _addr = server.accept()
                with conn:
                    socket.send_fds(conn, [b"m"], [fd])
            return fd
        client = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        with client:
            client.settimeout(60.0)
            client.connect(path[0])
            _msg, fds, _flags, _addr = socket.recv_fds(client, 16, 1)
        if not fds:
            raise RuntimeError("no file descriptor received from rank 0")
        return fds[0]
    finally:
        if server is not None:
            server.close()
            try:
                os.unlink(path[0])
            except OSError:
                pass


class MulticastGatheredOutput(object):
    """As `PeerGatheredOutput`, on NVSwitch multicast memory (csrc/multicast.cu): `tensor` is this rank's replica of the
    full `(n_items, ...)` output, `gather_into` stores every value ONCE to the multicast address and the switch delivers
    it to all ranks.  Collective constructor / `close()`; `barrier()`, `wait()`, `check()` as above.  Raises
    NotImplementedError where the devices cannot join a multicast object (no NVSwitch / driver support)."""

    def __init__(self, shape, device, group=None):
        if not dist.is_initialized():
            raise RuntimeError("MulticastGatheredOutput needs an initialised torch.distributed process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.shape = tuple(int(d) for d in shape)
        n = 1
        for d in self.shape:
            n *= d
        lib = _cabi.lib()
        self._obj = None
        if self.world < 2:
            raise NotImplementedError("MulticastGatheredOutput: a multicast object needs at least two devices "
                                      "(use PeerGatheredOutput, or no gather, on one GPU)")
        oks = [None] * self.world
        dist.all_gather_object(oks, multicast_supported(self.device), group=group)
        if not all(oks):
            raise NotImplementedError("MulticastGatheredOutput: multicast is not supported on rank(s) %s"
                                      % [r for r, ok in enumerate(oks) if not ok])
        failure, fd, obj = None, -1, ctypes.c_void_p()
        with torch.cuda.device(self.device):
            try:
                if self.rank == 0:
                    cfd = ctypes.c_int(-1)
                    _cabi.check(lib.tac_mc_create(4 * n, self.world, ctypes.byref(obj), ctypes.byref(cfd)))
                    fd = cfd.value
            except Exception as exc:
                failure = "rank 0: %s" % exc
            state = [failure]
            dist.broadcast_object_list(state, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            if state[0]:
                raise RuntimeError("MulticastGatheredOutput: creating the multicast object failed (%s)" % state[0])
            try:
                got = _share_fd(fd, self.rank, self.world, group)
                if self.rank != 0:
                    _cabi.check(lib.tac_mc_import(int(got), 4 * n, self.world, ctypes.byref(obj)))
                os.close(got)
                self._obj = obj
                _cabi.check(lib.tac_mc_add_device(obj))
            except Exception as exc:
                failure = "rank %d: %s" % (self.rank, exc)
            failure = self._agree(failure)                       # also the barrier: every device has been added
            local, mc = ctypes.c_void_p(), ctypes.c_void_p()
            if not failure:
                try:
                    _cabi.check(lib.tac_mc_bind(obj, ctypes.byref(local), ctypes.byref(mc)))
                except Exception as exc:
                    failure = "rank %d: %s" % (self.rank, exc)
            failure = self._agree(failure)                       # every replica is bound before anyone stores
            if failure:
                self._release()
                raise RuntimeError("MulticastGatheredOutput: set-up failed (%s)" % failure)
        self._local, self._mc = local.value, mc.value
        self.mc_payload = self._mc + 128                           # TAC_PEER_HEADER_BYTES
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4", "data": (self._local + 128, False),
                                         "version": 2, "strides": None}
        self.tensor = torch.as_tensor(self, device=self.device)
        self.epoch = 0

    def _agree(self, failure):
        failures = [None] * self.world
        dist.all_gather_object(failures, failure, group=self.group)
        failures = [f for f in failures if f]
        return "; ".join(failures) if failures else None

    def _release(self):
        obj, self._obj = self._obj, None
        if obj is not None and obj.value:
            with torch.cuda.device(self.device):
                _cabi.lib().tac_mc_free(obj)

    def barrier(self, timeout_s=20.0):
        self.epoch += 1
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().tac_mc_barrier(self._obj, self.rank, self.epoch, float(timeout_s), _cabi.stream_ptr(self.device)))

    def check(self):
        flag = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().tac_mc_timed_out(self._obj, ctypes.byref(flag)))
        if flag.value:
            raise RuntimeError("MulticastGatheredOutput: a barrier timed out waiting for a peer rank")

    def wait(self, timeout_s=20.0):
        self.barrier(timeout_s)
        self.check()
        return self.tensor

    def close(self):
        """Collective: every rank stops using the buffer, then releases its replica and its handle of the object."""
        if self._obj is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self.tensor = None
        self._release()

    def __del__(self):
        try:
            if getattr(self, "_obj", None) is not None:
                self.tensor = None
                self._release()
        except Exception:                                      # interpreter shutdown
            pass
