"""A minimal distributed array over ``torch.distributed`` with the part of ``caput.mpiarray.MPIArray``'s
interface that the hot path uses (``cora/core/skysim.py:97-134``, ``cora/signal/lss.py:436-470``):
``zeros``, ``wrap``, ``local_array``, ``global_shape``, ``local_shape``, ``local_offset``, ``axis``,
``enumerate``, ``allgather`` and ``redistribute``.

One axis is split over the ranks of a process group in caput's contiguous blocks (the first
``n % size`` ranks hold one extra item).  The local part is a numpy array on the host; the collectives
run on whatever backend the group has (gloo on CPU tensors, NCCL through a device staging copy).
``skysim.mkfullsky`` accepts these (and anything else that exposes ``local_array`` / ``global_shape`` /
``axis``) as its distributed ``corr`` input and returns its result wrapped the same way.
"""

import numpy as np


def _dist():
    import torch.distributed as dist

    return dist


def _size_rank(group):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def split_block(n, size, rank):
    """(start, stop) of rank's block of ``n`` items: caput's ``mpiutil.split_local`` rule."""
    base, rem = divmod(int(n), int(size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _staging_device(group):
    """Device that collective buffers must live on for this group's backend (None = host)."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl":
        import torch

        return torch.device("cuda", torch.cuda.current_device())
    return None


class MPIArray(object):
    """``local_array``: this rank's block (numpy); the global array is the blocks concatenated along ``axis``."""

    def __init__(self, local_array, axis, global_shape, group=None):
        self.local_array = local_array
        self.axis = int(axis)
        self.global_shape = tuple(int(x) for x in global_shape)
        self.comm = group
        size, rank = _size_rank(group)
        lo, hi = split_block(self.global_shape[self.axis], size, rank)
        if local_array.shape[self.axis] != hi - lo:
            raise ValueError("local block has %d items along axis %d, the block split gives rank %d %d"
                             % (local_array.shape[self.axis], self.axis, rank, hi - lo))
        off = [0] * local_array.ndim
        off[self.axis] = lo
        self.local_offset = tuple(off)

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def zeros(cls, global_shape, dtype=np.float64, axis=0, comm=None):
        size, rank = _size_rank(comm)
        lo, hi = split_block(global_shape[axis], size, rank)
        shp = list(global_shape)
        shp[axis] = hi - lo
        return cls(np.zeros(shp, dtype=dtype), axis, global_shape, comm)

    @classmethod
    def wrap(cls, array, axis, comm=None):
        """Turn every rank's local block into a distributed array (collective: the block lengths are summed and
        checked against the block split, like ``MPIArray.wrap``)."""
        size, rank = _size_rank(comm)
        array = np.asarray(array)
        n = array.shape[axis]
        if size > 1:
            import torch

            dist = _dist()
            dev = _staging_device(comm)
            t = torch.tensor([n], dtype=torch.int64, device=dev)
            dist.all_reduce(t, group=comm)
            n = int(t.item())
        gshape = list(array.shape)
        gshape[axis] = n
        return cls(array, axis, gshape, comm)

    # ---- views ------------------------------------------------------------------------------
    @property
    def local_shape(self):
        return tuple(self.local_array.shape)

    @property
    def shape(self):
        return tuple(self.local_array.shape)

    @property
    def dtype(self):
        return self.local_array.dtype

    def __getitem__(self, key):
        return self.local_array[key]

    def __setitem__(self, key, value):
        self.local_array[key] = value

    def enumerate(self, axis):
        """(local index, global index) pairs along ``axis``."""
        start = self.local_offset[axis]
        return [(i, start + i) for i in range(self.local_array.shape[axis])]

    # ---- collectives ------------------------------------------------------------------------
    def allgather(self):
        """The whole array on every rank (numpy)."""
        size, rank = _size_rank(self.comm)
        if size == 1:
            return np.array(self.local_array)
        import torch

        dist = _dist()
        dev = _staging_device(self.comm)
        n = self.global_shape[self.axis]
        nmax = -(-n // size)
        loc = np.moveaxis(self.local_array, self.axis, 0)
        rest = loc.shape[1:]
        buf = np.zeros((nmax,) + rest, dtype=self.local_array.dtype)
        buf[: loc.shape[0]] = loc
        tin = torch.from_numpy(np.ascontiguousarray(buf).view(np.uint8).reshape(-1)).to(dev)
        touts = [torch.empty_like(tin) for _ in range(size)]
        dist.all_gather(touts, tin, group=self.comm)
        allb = torch.stack(touts).cpu().numpy().view(self.local_array.dtype).reshape((size, nmax) + rest)
        parts = []
        for r in range(size):
            lo, hi = split_block(n, size, r)
            parts.append(allb[r, : hi - lo])
        return np.moveaxis(np.concatenate(parts, axis=0), 0, self.axis)

    def redistribute(self, axis):
        """The same global array split along another axis (one all-to-all)."""
        axis = int(axis)
        if axis == self.axis:
            return self
        size, rank = _size_rank(self.comm)
        gshape = self.global_shape
        if size == 1:
            return type(self)(self.local_array, axis, gshape, self.comm)
        import torch

        dist = _dist()
        dev = _staging_device(self.comm)
        dt = self.local_array.dtype
        mylo, myhi = split_block(gshape[axis], size, rank)
        pieces, in_split, out_split, out_shapes = [], [], [], []
        for r in range(size):
            lo, hi = split_block(gshape[axis], size, r)
            sl = [slice(None)] * self.local_array.ndim
            sl[axis] = slice(lo, hi)
            piece = np.ascontiguousarray(self.local_array[tuple(sl)]).view(np.uint8).reshape(-1)
            pieces.append(piece)
            in_split.append(piece.size)
            olo, ohi = split_block(gshape[self.axis], size, r)      # what rank r holds along the old axis
            shp = list(gshape)
            shp[self.axis] = ohi - olo
            shp[axis] = myhi - mylo
            out_shapes.append(shp)
            out_split.append(int(np.prod(shp)) * dt.itemsize)
        tin = torch.from_numpy(np.concatenate(pieces) if pieces else np.zeros(0, np.uint8)).to(dev)
        tout = torch.empty(sum(out_split), dtype=torch.uint8, device=dev)
        dist.all_to_all_single(tout, tin, output_split_sizes=out_split, input_split_sizes=in_split, group=self.comm)
        flat = tout.cpu().numpy()
        parts, o = [], 0
        for shp, nb in zip(out_shapes, out_split):
            parts.append(flat[o:o + nb].view(dt).reshape(shp))
            o += nb
        return type(self)(np.concatenate(parts, axis=self.axis), axis, gshape, self.comm)


def zeros(global_shape, dtype=np.float64, axis=0, comm=None):
    """``caput.mpiarray.zeros``."""
    return MPIArray.zeros(global_shape, dtype=dtype, axis=axis, comm=comm)


def is_distributed(x):
    """Duck test used by ``skysim.mkfullsky``: caput's ``MPIArray`` and this module's both pass."""
    return hasattr(x, "local_array") and hasattr(x, "global_shape")
