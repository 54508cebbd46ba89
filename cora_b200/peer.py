"""Peer memory between the GPUs of one box (one process per GPU).

``PeerGroup`` allocates buffers through the library (``cora_b200_peer_alloc``: cudaMalloc + CUDA IPC
handle), swaps the handles over ``torch.distributed`` and maps every other rank's buffer, so
that kernels can store straight into the GPU that will consume the data (NVLink / NVSwitch
stores from the kernel epilogue -- the exchange of ``cora/core/skysim.py:128`` without a separate
collective).  ``LocalPeers`` is the same interface for several *virtual* ranks living in one
process on one GPU; the single-GPU tests drive the exact kernels and pointer tables of the
multi-GPU path through it.
"""

import ctypes

import numpy as np

from . import _dev, _lib

_TYPESTR = {"float64": "<f8", "complex128": "<c16", "int32": "<i4", "int64": "<i8", "uint8": "|u1", "uint64": "<u8"}


class RawBuf(object):
    """Device memory that torch does not own (a library allocation or a mapped peer buffer)."""

    def __init__(self, ptr, nbytes, keep=None):
        self.ptr, self.nbytes, self._keep = int(ptr), int(nbytes), keep

    def data_ptr(self):
        return self.ptr

    def tensor(self, shape, dtype):
        """Zero-copy torch view (``__cuda_array_interface__``)."""
        t = _dev.torch()
        name = str(dtype).replace("torch.", "")
        shape = tuple(int(s) for s in shape)
        need = int(np.prod(shape)) * np.dtype(_TYPESTR[name]).itemsize
        if need > self.nbytes:
            raise ValueError("RawBuf view of %d bytes exceeds the %d-byte buffer" % (need, self.nbytes))

        class _View(object):
            pass

        v = _View()
        v.__cuda_array_interface__ = {"shape": shape, "typestr": _TYPESTR[name], "data": (self.ptr, False), "version": 2,
                                      "strides": None}
        v._owner = self
        out = t.as_tensor(v, device=_dev.device())
        out._cora_owner = self   # keep the allocation alive as long as the view
        return out


class PeerSetupError(_lib.CoraB200Error):
    """Peer memory could not be set up on every rank (raised on ALL ranks of the group together)."""


def _env_float(name, default):
    import os

    try:
        return float(os.environ.get(name, default))
    except ValueError:
        return float(default)


class PeerGroup(object):
    """Real multi-process group.  Every method that allocates is collective: all ranks call it
    in the same order.  A failure on any rank (no peer access between some pair of GPUs, out of
    memory) is agreed on inside the call: every rank takes part in the same collectives, releases
    what it had allocated or mapped for the failed request, and raises ``PeerSetupError``.

    ``timeout_s`` (default: ``CORA_B200_PEER_TIMEOUT_S`` or 60 s): how long a barrier waits for a
    peer.  ``fatal`` (default on; ``CORA_B200_PEER_FATAL=0`` turns it off): a timed-out barrier traps
    the kernel, so that nothing queued behind it runs on half-written buffers -- every later CUDA
    call of the process raises.  With ``fatal`` off the caller must call ``check()`` before trusting
    results produced after a barrier."""

    def __init__(self, rank, size, group=None, timeout_s=None, fatal=None):
        import torch.distributed as dist

        self.rank, self.size, self.group, self._dist = int(rank), int(size), group, dist
        self._own, self._opened = [], []
        self._epoch = 0
        self.timeout_s = _env_float("CORA_B200_PEER_TIMEOUT_S", 60.0) if timeout_s is None else float(timeout_s)
        self.fatal = bool(int(_env_float("CORA_B200_PEER_FATAL", 1))) if fatal is None else bool(fatal)
        t = _dev.torch()
        self.status = _dev.zeros((1,), t.int32)
        try:
            self._flags, flag_ptrs = self.alloc(8 * self.size)
        except PeerSetupError:
            self._release()
            raise
        self._flag_ptrs = _dev.to_device(np.array(flag_ptrs, dtype=np.uint64).view(np.int64), t.int64)
        t.cuda.synchronize()
        dist.barrier(group=group)   # every rank's zeroed flags exist before anybody signals

    def alloc(self, nbytes):
        """-> (own RawBuf, [pointer of every rank's buffer as mapped in this process])."""
        nbytes = max(256, int(nbytes))
        p = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        err = None
        try:
            _lib.call("cora_b200_peer_alloc", nbytes, ctypes.byref(p), handle)
        except _lib.CoraB200Error as exc:
            err, p = "rank %d: %s" % (self.rank, exc), ctypes.c_void_p()
        handles = [None] * self.size
        self._dist.all_gather_object(handles, None if err else bytes(handle.raw), group=self.group)
        ptrs, opened = [], []
        if err is None and all(h is not None for h in handles):
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(p.value)
                    continue
                q = ctypes.c_void_p()
                try:
                    _lib.call("cora_b200_peer_open", ctypes.create_string_buffer(h, 64), ctypes.byref(q))
                except _lib.CoraB200Error as exc:
                    err = "rank %d mapping rank %d: %s" % (self.rank, r, exc)
                    break
                opened.append(q.value)
                ptrs.append(q.value)
        elif err is None:
            err = "rank %d: a peer could not allocate" % self.rank
        # agree on the outcome: every rank has run the same collectives up to here whatever happened locally
        errs = [None] * self.size
        self._dist.all_gather_object(errs, err, group=self.group)
        if any(e is not None for e in errs):
            lib = _lib.load()
            for q in opened:
                lib.cora_b200_peer_close(ctypes.c_void_p(q))
            self._dist.barrier(group=self.group)        # nobody still maps a buffer about to be freed
            if p.value:
                lib.cora_b200_peer_free(p)
            raise PeerSetupError("peer memory setup failed: " + "; ".join(e for e in errs if e))
        self._own.append(p.value)
        self._opened.extend(opened)
        return RawBuf(p.value, nbytes, keep=self), ptrs

    def barrier(self, stream=None):
        """Stream-ordered flag barrier over peer memory (all earlier stores of every rank are
        visible to every rank's later kernels)."""
        self._epoch += 1
        _lib.call("cora_b200_peer_barrier", _lib.ptr(self._flag_ptrs), self.rank, self.size,
                  ctypes.c_ulonglong(self._epoch), self.timeout_s, _lib.ptr(self.status), int(self.fatal),
                  _lib.stream_ptr(stream))

    def check(self):
        """Raise if a barrier timed out (synchronises)."""
        s = int(self.status.item())
        if s:
            raise _lib.CoraB200Error("peer barrier timed out waiting for rank %d" % (s - 1))

    def _release(self):
        """Unmap and free everything (local; the caller orders it against the other ranks)."""
        lib = _lib.load()
        for q in self._opened:
            lib.cora_b200_peer_close(ctypes.c_void_p(q))
        self._opened = []
        for p in self._own:
            lib.cora_b200_peer_free(ctypes.c_void_p(p))
        self._own = []

    def close(self):
        t = _dev.torch()
        t.cuda.synchronize()
        self._dist.barrier(group=self.group)    # nobody is still writing into a buffer about to go
        lib = _lib.load()
        for q in self._opened:
            lib.cora_b200_peer_close(ctypes.c_void_p(q))
        self._opened = []
        self._dist.barrier(group=self.group)
        for p in self._own:
            lib.cora_b200_peer_free(ctypes.c_void_p(p))
        self._own = []


class LocalPeers(object):
    """G virtual ranks in one process (tests): buffers are ordinary torch allocations, the
    pointer lists are resolved once every virtual rank has allocated its share."""

    def __init__(self, size):
        self.size = int(size)
        self._slots = []       # per alloc call index: [tensor or None] * size
        self._calls = [0] * self.size

    def view(self, rank):
        return _LocalView(self, rank)


class _LocalView(object):
    def __init__(self, owner, rank):
        self.owner, self.rank, self.size = owner, int(rank), owner.size
        self.status = _dev.zeros((1,), _dev.torch().int32)

    def alloc(self, nbytes):
        o = self.owner
        k = o._calls[self.rank]
        o._calls[self.rank] += 1
        while len(o._slots) <= k:
            o._slots.append([None] * o.size)
        buf = _dev.zeros((max(256, int(nbytes)),), _dev.torch().uint8)
        o._slots[k][self.rank] = buf
        slot = o._slots[k]

        class _Lazy(list):
            """Pointer list that fills itself in on first use (all virtual ranks allocated by then)."""

            def resolve(self_inner):
                if any(b is None for b in slot):
                    raise RuntimeError("LocalPeers: not every virtual rank has allocated this buffer yet")
                return [b.data_ptr() for b in slot]

        return RawBuf(buf.data_ptr(), buf.numel(), keep=buf), _Lazy()

    def barrier(self, stream=None):
        pass    # virtual ranks are stepped phase by phase by the caller

    def check(self):
        pass

    def close(self):
        pass


def resolve(ptrs):
    """Pointer list of ``alloc`` -> plain list of ints."""
    return ptrs.resolve() if hasattr(ptrs, "resolve") else list(ptrs)
