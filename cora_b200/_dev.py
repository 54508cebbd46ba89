"""Device-side plumbing shared by the host modules: buffers (torch owns the memory), SHT plan
cache, workspace sizing."""

import ctypes

import numpy as np

from . import _lib

_plans = {}

# bytes moved across PCIe by this package (bench.py reports them as h2d/d2h_bytes_per_step)
traffic = {"h2d": 0, "d2h": 0}


def torch():
    return _lib.require_cuda()


def device():
    t = torch()
    return t.device("cuda", t.cuda.current_device())


def to_device(a, dtype=None):
    """numpy array / torch tensor -> contiguous CUDA tensor (a copy for host input)."""
    t = torch()
    if isinstance(a, t.Tensor):
        x = a.to(device())
        if dtype is not None and x.dtype != dtype:
            x = x.to(dtype)
        return x.contiguous()
    arr = np.ascontiguousarray(a)
    traffic["h2d"] += arr.nbytes
    x = t.from_numpy(arr).to(device(), non_blocking=False)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x


def to_host(x):
    """CUDA tensor -> numpy array through a pinned host buffer (torch's pinned-memory cache owns
    the block; the returned array keeps it alive and hands it back when it is dropped)."""
    t = torch()
    x = x.contiguous()
    h = t.empty(x.shape, dtype=x.dtype, pin_memory=True)
    h.copy_(x, non_blocking=True)
    t.cuda.current_stream().synchronize()
    traffic["d2h"] += h.numel() * h.element_size()
    return h.numpy()


_stage = {}


def stream_to_host(x, chunk_bytes=1 << 30):
    """Copy a CUDA tensor to the host through a bounded two-slot pinned staging ring (for results
    larger than what can be held pinned, e.g. 26 GB of maps per GPU): every byte crosses PCIe,
    nothing but the ring is retained.  Returns the number of bytes copied."""
    t = torch()
    flat = x.contiguous().view(-1)
    esz = flat.element_size()
    n = max(1, int(chunk_bytes) // esz)
    key = (t.cuda.current_device(), flat.dtype, n)
    if key not in _stage:
        _stage[key] = [t.empty((n,), dtype=flat.dtype, pin_memory=True) for _ in range(2)]
    ring = _stage[key]
    evs = [None, None]
    st = t.cuda.current_stream()
    for i, o in enumerate(range(0, flat.numel(), n)):
        k = i & 1
        if evs[k] is not None:
            evs[k].synchronize()
        m = min(n, flat.numel() - o)
        ring[k][:m].copy_(flat[o : o + m], non_blocking=True)
        evs[k] = t.cuda.Event()
        evs[k].record(st)
    st.synchronize()
    traffic["d2h"] += flat.numel() * esz
    return flat.numel() * esz


def bind_host_to_gpu(index=None):
    """Pin the calling process to the CPU cores NVML reports as local to the GPU, so that pinned
    host buffers are first-touched on the GPU's NUMA node (the map D2H copy is the slowest leg of
    the end-to-end path and halves its speed across a socket link).  Returns True if applied."""
    import os

    try:
        import pynvml

        t = torch()
        idx = t.cuda.current_device() if index is None else int(index)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            idx = int(vis.split(",")[idx])
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx))
        return True
    except Exception:
        return False


_copy_streams = {}


def copy_stream():
    """Side stream for device->host copies overlapped with compute (one per device)."""
    t = torch()
    d = t.cuda.current_device()
    if d not in _copy_streams:
        _copy_streams[d] = t.cuda.Stream()
    return _copy_streams[d]


def empty(shape, dtype):
    return torch().empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype):
    return torch().zeros(shape, dtype=dtype, device=device())


def workspace(nbytes):
    return torch().empty(int(nbytes), dtype=torch().uint8, device=device())


def free_bytes():
    free, _ = torch().cuda.mem_get_info()
    return int(free)


def sht_plan(nside, lmax):
    """Cached opaque SHT plan for (device, nside, lmax)."""
    t = torch()
    key = (t.cuda.current_device(), int(nside), int(lmax))
    if key not in _plans:
        h = ctypes.c_void_p()
        _lib.call("cora_b200_sht_plan_create", int(nside), int(lmax), ctypes.byref(h))
        _plans[key] = h
    return _plans[key]


def sht_workspace(plan, layout, nchan, mult=1, reserve=2 << 30):
    """Workspace for alm2map: whole-problem size if it fits, else what is free minus a reserve."""
    lib = _lib.load()
    per16 = lib.cora_b200_alm2map_workspace_bytes(plan, layout, 16) * mult
    need = lib.cora_b200_alm2map_workspace_bytes(plan, layout, int(nchan)) * mult
    avail = free_bytes() - reserve
    nbytes = min(need, max(per16, avail))
    return workspace(nbytes), nbytes
