"""Frequency-sharded map writer: the step after the hot path (``cora/scripts/makesky.py:412-450`` ``write_map``,
container layout ``cora/core/containers.py:90-128``).

The reference writes one HDF5 file with the dataset ``map[freq, pol, pixel]`` (float64; distributed over ``freq`` in
caput's memh5 convention) and ``index_map/{freq (centre, width), pol, pixel}``.  h5py / libhdf5 are not available
in this environment, and a 206 GB result (nside 1024 x 2048 channels) lives in eight GPUs' memory, so the writer
here streams every rank's channel block ``[freq_local, pol, pixel]`` from its GPU into its own ``.npy`` file
(through a bounded pinned staging ring, one PCIe pass, nothing else held on the host) and rank 0 writes a JSON
header with the index maps, the attributes and the shard table:

    outdir/
      index_map.json        {"axis": ["freq","pol","pixel"], "freq": {"centre": [...], "width": [...]},
                             "pol": ["I","Q","U","V"], "npix": N, "dtype": "<f8",
                             "attrs": {"__memh5_distributed_file": true}, "distributed_axis": "freq",
                             "shards": [{"file": "map.0000.npy", "freq_start": 0, "freq_end": 256}, ...]}
      map.0000.npy ...      float64 [freq_end - freq_start, npol, npix], C order (numpy format 1.0/2.0)

``tools/map_to_hdf5.py`` (h5py, ~30 lines) converts the directory into the reference's exact HDF5 layout, shard by
shard; ``read_map`` loads it back as one array (small maps, tests).
"""

import json
import os

import numpy as np

POL_FULL = ["I", "Q", "U", "V"]


def _as_3d_shape(shape, include_pol):
    """Shape and polarisation labels of ``write_map``'s 3-D array for a local block (``makesky.py:418-429``)."""
    if len(shape) == 3:
        return (shape[0], shape[1], shape[2]), POL_FULL[: shape[1]] if shape[1] != 4 else POL_FULL, None
    if include_pol:
        return (shape[0], 4, shape[1]), POL_FULL, 0          # Stokes I in plane 0, Q = U = V = 0
    return (shape[0], 1, shape[1]), ["I"], 0


def write_map(outdir, data, freq, fwidth=None, include_pol=True, freq_start=0, rank=0, size=1, chunk_bytes=256 << 20):
    """Write this rank's block of the map.

    ``data``: ``[freq_local, npix]`` or ``[freq_local, npol, npix]`` float64, a CUDA tensor (streamed) or a numpy
    array; ``freq``: the FULL frequency axis (every rank passes the same); ``freq_start``: global index of this
    rank's first channel.  Every rank writes ``map.<rank>.npy``; rank 0 also writes ``index_map.json`` -- for
    ``size > 1`` it needs the other ranks' channel ranges, which follow caput's block split of ``len(freq)``
    (``dist.block_partition``), the distribution the sharded generator uses.
    """
    from .dist import block_partition

    os.makedirs(outdir, exist_ok=True)
    freq = np.asarray(freq, dtype=np.float64)
    shape3, polmap, plane = _as_3d_shape(tuple(int(s) for s in data.shape), include_pol)
    fname = "map.%04d.npy" % rank
    out = np.lib.format.open_memmap(os.path.join(outdir, fname), mode="w+", dtype=np.float64, shape=shape3)
    is_torch = hasattr(data, "is_cuda")
    nfl = shape3[0]
    npix = shape3[2]
    per_chan = (1 if plane is not None else shape3[1]) * npix * 8
    step = max(1, int(chunk_bytes) // per_chan)
    if is_torch and data.is_cuda:
        import torch

        ring = [torch.empty((step,) + tuple(data.shape[1:]), dtype=torch.float64, pin_memory=True) for _ in range(2)]
        evs = [None, None]
        st = torch.cuda.current_stream()
        pending = []
        for k, c0 in enumerate(range(0, nfl, step)):
            n = min(step, nfl - c0)
            slot = k & 1
            if evs[slot] is not None:
                evs[slot].synchronize()
                _flush(out, pending.pop(0), plane)
            ring[slot][:n].copy_(data[c0:c0 + n], non_blocking=True)
            evs[slot] = torch.cuda.Event()
            evs[slot].record(st)
            pending.append((c0, n, ring[slot]))
        st.synchronize()
        while pending:
            _flush(out, pending.pop(0), plane)
    else:
        arr = data.cpu().numpy() if is_torch else np.asarray(data, dtype=np.float64)
        if plane is not None:
            out[:, plane, :] = arr
        else:
            out[...] = arr
    out.flush()
    del out
    if rank == 0:
        width = float(fwidth) if fwidth is not None else float(np.abs(np.diff(freq)[0]))
        shards = []
        for r in range(size):
            lo, hi = block_partition(len(freq), size, r) if size > 1 else (int(freq_start), int(freq_start) + nfl)
            shards.append({"file": "map.%04d.npy" % r, "freq_start": int(lo), "freq_end": int(hi)})
        hdr = {"axis": ["freq", "pol", "pixel"], "freq": {"centre": freq.tolist(), "width": [width] * len(freq)},
               "pol": polmap, "npix": int(npix), "dtype": "<f8", "distributed_axis": "freq",
               "attrs": {"__memh5_distributed_file": True}, "shards": shards}
        with open(os.path.join(outdir, "index_map.json"), "w") as f:
            json.dump(hdr, f)
    return os.path.join(outdir, fname)


def _flush(out, item, plane):
    c0, n, buf = item
    a = buf[:n].numpy()
    if plane is not None:
        out[c0:c0 + n, plane, :] = a
    else:
        out[c0:c0 + n] = a


def read_header(outdir):
    with open(os.path.join(outdir, "index_map.json")) as f:
        return json.load(f)


def read_map(outdir):
    """-> (map float64[nfreq, npol, npix], header dict).  Loads every shard (use the shards' memmaps for big maps)."""
    hdr = read_header(outdir)
    nfreq = len(hdr["freq"]["centre"])
    full = np.empty((nfreq, len(hdr["pol"]), hdr["npix"]), dtype=np.float64)
    for sh in hdr["shards"]:
        full[sh["freq_start"]:sh["freq_end"]] = np.load(os.path.join(outdir, sh["file"]), mmap_mode="r")
    return full, hdr
