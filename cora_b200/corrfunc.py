"""xi(r) -> C_l(chi, chi'): the LSS front end that produces ``corr`` for ``mkfullsky``.

Mirrors ``cora/signal/corrfunc.py:265-400`` (``legendre_array``, ``corr_to_clarray``: same names, arguments and
return layout).  The correlation function is the caller's host callable, evaluated chunk by chunk of the
Gauss-Legendre nodes exactly as the reference does; everything that scales -- the radial bin quadrature, the
weighted Legendre table and the ``(L x M) . (M x nx^2)`` contraction -- runs in the CUDA kernels
(``csrc/corrfunc.cu``: FP64 tensor-core GEMM accumulated chunk after chunk, so the ``M x nx^2`` integrand is never
held at once; ``csrc/cl.cu``: the two-sided weighted bin average)."""

import numpy as np

from . import _dev, _lib


def cosine_rule(mu, x1, x2):
    """Separation of points at radii ``x1[a]``, ``x2[b]`` whose directions make an angle with cosine ``mu[i]``:
    ``float64[len(mu), len(x1), len(x2)]`` (``caput.astro.coordinates.spherical.cosine_rule`` as called at
    ``corrfunc.py:369``), in the cancellation-free form ``sqrt((x1 - x2)^2 + 2 x1 x2 (1 - mu))``."""
    mu = np.asarray(mu)[:, np.newaxis, np.newaxis]
    a = np.asarray(x1)[np.newaxis, :, np.newaxis]
    b = np.asarray(x2)[np.newaxis, np.newaxis, :]
    return np.sqrt((a - b) ** 2 + 2.0 * a * b * (1.0 - mu))


class TabulatedCorrelation(object):
    """A correlation function given by samples: ``xi(r)`` piecewise linear in ``r`` (``kind="linear"``) or in
    ``ln r`` (``kind="log"``), constant beyond both ends -- ``numpy.interp`` semantics.  Callable on the host like
    any ``corr`` the reference takes; ``corr_to_clarray`` recognises it and evaluates cosine rule, interpolation and
    the radial-bin quadrature in one CUDA kernel (``cora_b200_corr_bins``) instead of calling back to the host."""

    def __init__(self, r, xi, kind="linear"):
        if kind not in ("linear", "log"):
            raise ValueError("kind must be 'linear' or 'log'")
        self.kind = kind
        r = np.asarray(r, dtype=np.float64)
        self.knots = np.ascontiguousarray(np.log(r) if kind == "log" else r)
        self.values = np.ascontiguousarray(xi, dtype=np.float64)
        if self.knots.ndim != 1 or self.knots.shape != self.values.shape or self.knots.size < 2 or np.any(np.diff(self.knots) <= 0):
            raise ValueError("r must be 1-D, increasing, and match xi")

    def __call__(self, r):
        r = np.asarray(r, dtype=np.float64)
        with np.errstate(divide="ignore"):
            x = np.log(r) if self.kind == "log" else r
        return np.interp(x, self.knots, self.values)


def legendre_array(lmax, mu, scale=None, device_out=False):
    """Legendre polynomials ``P_l(mu_i)`` up to ``lmax``: ``float64[lmax + 1, len(mu)]`` (``corrfunc.py:265-287``),
    optionally times ``scale[i]``; three-term recurrence on the GPU (``cora_b200_legendre_table``)."""
    t = _dev.torch()
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    n = mu.size
    out = _dev.empty((lmax + 1, n), t.float64)
    mud = _dev.to_device(mu, t.float64)
    sd = None if scale is None else _dev.to_device(np.ascontiguousarray(scale, dtype=np.float64), t.float64)
    _lib.call("cora_b200_legendre_table", _lib.ptr(mud), _lib.ptr(sd), n, int(lmax), _lib.ptr(out), n, _lib.stream_ptr())
    return out if device_out else _dev.to_host(out)


def _radial_nodes(xarray, xromb, xwidth):
    import scipy.special as ss

    if xromb <= 0:
        return xarray, np.ones(1), 1
    if xwidth is None:
        xhalf = np.empty(xarray.shape)
        xhalf[0] = np.abs(xarray[1] - xarray[0]) / 2.0       # first and second bin share a width (corrfunc.py:343-346)
        xhalf[1:] = np.abs(xarray[1:] - xarray[:-1]) / 2.0
    else:
        xhalf = np.ones(xarray.shape) * xwidth / 2.0
    xint = 2**xromb + 1
    x_r, x_w, x_wsum = ss.roots_legendre(xint, mu=True)
    return (xarray[:, np.newaxis] + xhalf[:, np.newaxis] * x_r).flatten(), x_w / x_wsum, xint


def corr_to_clarray(corr, lmax, xarray, xromb=3, xwidth=None, q=2, chunksize=50, device_out=False, group=None):
    """Calculate an array of ``C_l(chi_1, chi_2)`` from a real-space correlation function.

    Same contract as ``cora/signal/corrfunc.py:290-400``: ``corr(r)`` is evaluated at the separations of every pair
    of radial samples for each of the ``M = q lmax`` Gauss-Legendre nodes in ``mu`` (in chunks of ``chunksize``
    nodes), averaged over the radial bins with a ``2**xromb + 1``-point Gauss-Legendre rule, and contracted with
    ``P_l(mu_i) w_i 4 pi / sum(w)``.  Returns ``float64[lmax + 1, nx, nx]``.

    Under an initialised ``torch.distributed`` group of more than one rank the nodes are split over the ranks in
    caput's contiguous blocks (the reference's ``mpiarray.zeros((M, nx, nx), axis=0)``, ``corrfunc.py:362-364``),
    every rank contracts its own nodes, the partial sums are added over ranks, and the result comes back
    distributed over l as a ``cora_b200.mpiarray.MPIArray`` (the reference's ``redistribute(axis=0)``,
    ``:397-399``) -- the form ``mkfullsky`` takes."""
    import scipy.special as ss

    from . import mpiarray

    t = _dev.torch()
    xarray = np.asarray(xarray, dtype=np.float64)
    M = int(q) * int(lmax)
    mu, w, wsum = ss.roots_legendre(M, mu=True)
    xa, x_w, xint = _radial_nodes(xarray, xromb, xwidth)
    xlen = xarray.size
    L = int(lmax) + 1
    size, rank = mpiarray._size_rank(group)
    clo, chi_ = mpiarray.split_block(M, size, rank)
    nloc = chi_ - clo

    lm = legendre_array(lmax, mu[clo:chi_], scale=(w * 4.0 * np.pi / wsum)[clo:chi_], device_out=True)
    out = _dev.zeros((L, xlen * xlen), t.float64)
    wd = _dev.to_device(np.ascontiguousarray(x_w), t.float64)
    fused = type(corr) is TabulatedCorrelation       # (a subclass may override __call__: take the generic path)
    if fused:
        tr, tv = _dev.to_device(corr.knots, t.float64), _dev.to_device(corr.values, t.float64)
        xad = _dev.to_device(np.ascontiguousarray(xa), t.float64)
        mud = _dev.to_device(np.ascontiguousarray(mu[clo:chi_]), t.float64)
    first = True
    # (fewer local nodes than one chunk: np.array_split(..., 0) raises ValueError, as in the reference, corrfunc.py:367)
    for msec in np.array_split(np.arange(nloc), nloc // int(chunksize)):
        nm = len(msec)
        blk = None
        if fused:
            red = _dev.empty((nm, xlen * xlen), t.float64)
            _lib.call("cora_b200_corr_bins", _lib.ptr_off(mud, 8 * int(msec[0])), nm, _lib.ptr(xad), _lib.ptr(wd), xlen, xint,
                      _lib.ptr(tr), _lib.ptr(tv), int(corr.knots.size), 1 if corr.kind == "log" else 0, _lib.ptr(red),
                      _lib.stream_ptr())
        else:
            rc = cosine_rule(mu[clo + msec], xa, xa)
            corr1 = np.ascontiguousarray(np.broadcast_to(corr(rc), rc.shape), dtype=np.float64)
            blk = _dev.to_device(corr1, t.float64)
            if xromb > 0:
                red = _dev.empty((nm, xlen * xlen), t.float64)
                _lib.call("cora_b200_cl_romberg_reduce", _lib.ptr(blk), _lib.ptr(wd), nm, xlen, xint, _lib.ptr(red), _lib.stream_ptr())
            else:
                red = blk.reshape(nm, xlen * xlen)
        # out += lm[:, msec] @ red     (A = the chunk's columns of the weighted Legendre table, row pitch nloc)
        _lib.call("cora_b200_dgemm", _lib.ptr_off(lm, 8 * int(msec[0])), _lib.ptr(red), _lib.ptr(out), L, xlen * xlen, nm,
                  nloc, xlen * xlen, xlen * xlen, 0 if first else 1, _lib.stream_ptr())
        first = False
        t.cuda.current_stream().synchronize()      # blk / red are released before the next chunk is staged
    if size > 1:
        import torch.distributed as dist

        dist.all_reduce(out, group=group)
        lo, hi = mpiarray.split_block(L, size, rank)
        loc = _dev.to_host(out[lo:hi]).reshape(hi - lo, xlen, xlen)
        return mpiarray.MPIArray(loc, 0, (L, xlen, xlen), group)
    out = out.reshape(L, xlen, xlen)
    return out if device_out else _dev.to_host(out)
