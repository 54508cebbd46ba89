"""Full-sky correlated Gaussian fields: ``clarray`` and ``mkfullsky``.

Mirrors ``cora/core/skysim.py:10-136`` (same names, argument meaning, return layouts and
error behaviour); the arithmetic runs in the CUDA kernels of ``csrc/`` through the C ABI.
"""

import ctypes

import numpy as np

from . import _dev, _lib, hputil, nputil


def romberg_weights(zromb):
    """Normalised Romberg weights ``w`` (sum 1) for ``2**zromb + 1`` equally spaced samples:
    ``romb(y, dx) / (2 h) == w . y``.  Built from the Richardson tableau of the trapezoid rule
    (what ``scipy.integrate.romb`` evaluates, ``skysim.py:64-65``)."""
    n = 2**zromb + 1
    if zromb == 0:
        return np.ones(1)
    rows = []
    for k in range(zromb + 1):  # trapezoid with 2^k intervals, as weights on the n samples
        step = 2 ** (zromb - k)
        w = np.zeros(n)
        w[::step] = 1.0
        w[0] = w[-1] = 0.5
        rows.append(w * step)
    R = [rows]
    for j in range(1, zromb + 1):
        prev = R[-1]
        R.append([(4**j * prev[i + 1] - prev[i]) / (4**j - 1) for i in range(len(prev) - 1)])
    return R[-1][0] / (n - 1)


def _sample_frequencies(zarray, zromb, zwidth):
    zarray = np.asarray(zarray, dtype=np.float64)
    if zromb == 0:
        return zarray.copy(), 1
    zsort = np.sort(zarray)
    zhalf = np.abs(zsort[1] - zsort[0]) / 2.0 if zwidth is None else zwidth / 2.0
    zint = 2**zromb + 1
    za = (zarray[:, np.newaxis] + np.linspace(-zhalf, zhalf, zint)[np.newaxis, :]).flatten()
    return za, zint


def _fused_fill(aps, owner):
    """The owner's fused fill kernel, but only if ``aps`` is this package's own implementation: a subclass
    that overrides ``angular_powerspectrum`` (or, for the separable foregrounds, ``angular_ps`` /
    ``frequency_covariance``; for 21cm any method the spectrum is built from) must be evaluated through the
    callable, as the reference always does (``cora/core/skysim.py:57``)."""
    if owner is None or getattr(aps, "__name__", "") != "angular_powerspectrum":
        return None
    fill = getattr(owner, "_b200_fill", None)
    origin = getattr(type(owner), "_b200_fill_origin", None)
    if fill is None or origin is None:
        return None
    for name in origin._b200_fill_methods:
        if getattr(type(owner), name, None) is not getattr(origin, name, None) or name in getattr(owner, "__dict__", {}):
            return None
    return fill


def clarray(aps, lmax, zarray, zromb=3, zwidth=None, device_out=False, _lower_only=False):
    """Calculate an array of C_l(z, z') averaged over each frequency channel.

    Same contract as ``cora/core/skysim.py:10-69``: ``aps(l, z1, z2)`` is the angular power
    spectrum, channels are sampled at ``2**zromb + 1`` points and Romberg-integrated in both
    arguments; returns ``float64[lmax+1, len(zarray), len(zarray)]``.

    If ``aps`` is the bound ``angular_powerspectrum`` of one of this package's spectrum
    classes (SCK foregrounds, ``Corr21cm``) the whole table is filled by one fused CUDA kernel.
    Any other callable is evaluated in l-sections (the reference's chunking) and the Romberg
    average runs on the GPU.
    """
    t = _dev.torch()
    zarray = np.asarray(zarray, dtype=np.float64)
    nz = zarray.size
    owner = getattr(aps, "__self__", None)
    fused = _fused_fill(aps, owner)

    if zromb != 0 and lmax < 5:
        # the reference dies in np.array_split(..., 0) (skysim.py:51); keep the error type
        raise ValueError("number sections must be larger than 0.")

    za, zint = _sample_frequencies(zarray, zromb, zwidth)
    w = romberg_weights(zromb)

    if fused is not None:
        out = _dev.empty((lmax + 1, nz, nz), t.float64)
        # the 21cm kernel fills the lower triangles (full 32-byte sectors); the mirror pass makes the table symmetric
        # as the reference's is, unless the caller only feeds mkfullsky (which reads the lower triangle, like LAPACK)
        lower = getattr(owner, "_b200_fill_lower", False)
        if lower and _lower_only:
            out.zero_()
        fused(owner._b200_fill_inputs(za, w), 0, 1, lmax + 1, nz, zint, out, lower_only=lower)
        if lower and not _lower_only:
            _lib.call("cora_b200_cl_symmetrize", _lib.ptr(out), lmax + 1, nz, _lib.stream_ptr())
        return out if device_out else _dev.to_host(out)

    if zromb == 0:
        res = aps(np.arange(lmax + 1)[:, np.newaxis, np.newaxis], zarray[np.newaxis, :, np.newaxis],
                  zarray[np.newaxis, np.newaxis, :])
        return _dev.to_device(res, t.float64) if device_out else res

    out = _dev.empty((lmax + 1, nz, nz), t.float64)
    wd = _dev.to_device(w, t.float64)
    for lsec in np.array_split(np.arange(lmax + 1), lmax // 5):
        l_first = int(lsec[0])  # before the call: reference-style callables may mutate lsec
        clt = aps(lsec[:, np.newaxis, np.newaxis], za[np.newaxis, :, np.newaxis], za[np.newaxis, np.newaxis, :])
        blk = _dev.to_device(np.ascontiguousarray(np.broadcast_to(clt, (len(lsec), nz * zint, nz * zint))), t.float64)
        _lib.call("cora_b200_cl_romberg_reduce", _lib.ptr(blk), _lib.ptr(wd), len(lsec), nz, zint,
                  _lib.ptr(out[l_first:]), _lib.stream_ptr())
    t.cuda.current_stream().synchronize()
    return out if device_out else _dev.to_host(out)


def draw_apply_device(root, l_list, dense_flag, nz, lmax, panel, seed=0, gauss=None, chan0=0, nu0=0, nnu=None,
                      stream=None):
    """Draw + apply on device buffers (see ``cora_b200_draw_apply`` in include/cora_b200.h)."""
    t = _dev.torch()
    lib = _lib.load()
    l_list = np.ascontiguousarray(l_list, dtype=np.int32)
    nl = len(l_list)
    nnu = nz if nnu is None else nnu
    if gauss is None:
        full = lib.cora_b200_draw_apply_workspace_bytes(nz, int(l_list.max()), nl)
        one = lib.cora_b200_draw_apply_workspace_bytes(nz, int(l_list.max()), 1)
        nbytes = min(full, max(one, _dev.free_bytes() - (2 << 30)))
        gauss_ld = 0
    else:
        nbytes = 64 * nl + 4096
        gauss_ld = int(gauss.shape[-1])
    ws = _dev.workspace(nbytes)
    _lib.call("cora_b200_draw_apply", _lib.ptr(root), _lib.ptr(l_list), _lib.ptr(dense_flag), nl, int(nz), int(lmax),
              ctypes.c_ulonglong(int(seed)), _lib.ptr(gauss), gauss_ld, _lib.ptr(panel), int(panel.shape[1]), int(chan0),
              int(nu0), int(nnu), _lib.ptr(ws), int(nbytes), _lib.stream_ptr(stream))
    return panel


def mkfullsky(corr, nside, alms=False, rng=None, *, seed=None, roots=None, gauss=None, device_out=False):
    """Construct a set of correlated Healpix maps (``cora/core/skysim.py:72-136``).

    Parameters
    ----------
    corr : np.ndarray or CUDA tensor (lmax+1, numz, numz)
        The correlation matrix C_l(z, z').  A distributed array (``cora_b200.mpiarray.MPIArray`` or caput's,
        split over l) takes the multi-GPU path and returns the maps distributed over frequency, as the
        reference does (``skysim.py:97-103,128-134``).
    nside : integer
        The resolution of the Healpix maps.
    alms : boolean, optional
        If True return the alms ``complex128[numz, 1, lmax+1, lmax+1]`` instead of the maps.
    rng : numpy Generator, optional
        If given, the Gaussian draws are taken from it exactly as the reference does (per l,
        ascending: real block then imaginary block of shape (numz, l+1)) and applied on the GPU
        -- the identical-draw parity path.  If None, draws come from the device Philox
        generator (the reference would use numpy's legacy global state).

    Keyword-only extensions (not in the reference signature): ``seed`` for the Philox path;
    ``roots`` (float64[L, numz, numz]) / ``gauss`` (complex128[L, numz, L]) to inject the
    reference's own matrix roots / draws; ``device_out`` to get CUDA tensors back.

    Returns
    -------
    hpmaps : np.ndarray (numz, npix)   -- or the alm array if ``alms``.
    """
    t = _dev.torch()
    if hasattr(corr, "local_array") and hasattr(corr, "global_shape"):
        # distributed corr (caput MPIArray / cora_b200.mpiarray.MPIArray, split over l): skysim.py:97-134
        from . import dist as _cdist

        return _cdist.mkfullsky_mpi(corr, nside, alms=alms, rng=rng, seed=seed)
    numz = corr.shape[1]
    maxl = corr.shape[0] - 1
    if corr.shape[2] != numz:
        raise Exception("Correlation matrix is incorrect shape.")
    L = maxl + 1
    nalm = L * (L + 1) // 2

    if roots is None:
        cl = _dev.to_device(corr, t.float64)
        root, used, _ = nputil.root_batched_device(cl, jitter_rel=1e-14, clip_rel=1e-16)
        del cl
    else:
        root = _dev.to_device(roots, t.float64)
        used = None  # injected roots: treat as dense

    gdev = None
    if gauss is not None:
        gdev = _dev.to_device(gauss, t.complex128)
    elif rng is not None:
        g = np.zeros((L, numz, L), dtype=np.complex128)
        for l in range(L):
            g[l, :, : l + 1] = nputil.complex_std_normal((numz, l + 1), rng=rng)
        gdev = _dev.to_device(g, t.complex128)
    elif seed is None:
        seed = int(np.random.randint(0, 2**31 - 1))

    panel = _dev.empty((nalm, numz), t.complex128)
    draw_apply_device(root, np.arange(L), used, numz, maxl, panel, seed=seed or 0, gauss=gdev)
    del root, gdev

    if alms:
        dense = hputil.panel_to_dense(panel, maxl, numz)  # [numz, L, L]
        out = dense.reshape(numz, 1, L, L)
        return out if device_out else _dev.to_host(out)

    if not device_out:
        return hputil.alm2map_to_host(panel, nside, maxl, numz)
    return hputil.alm2map_device(panel, nside, maxl, _lib.ALM_PANEL, numz, numz)


def mkconstrained(corr, constraints, nside, device_out=False):
    """Construct a set of Healpix maps satisfying given constraints on specified frequency
    slices, by using the lowest eigenmodes (``cora/core/skysim.py:139-201``).

    ``corr``: ``float64[lmax+1, numz, numz]``; ``constraints``: ``[[frequency_index, healpix map], ...]``.
    Per l the ``nmodes = len(constraints)`` eigenvectors of the largest eigenvalues are found
    (batched Jacobi eigh on the GPU), the constraint maps go to harmonic space
    (``healpy.map2alm`` defaults: no ring weights, 3 refinements), the mode amplitudes that
    reproduce the constrained slices are solved for and projected across all frequencies
    (``cv = trans^T solve(tmat^T, cmap)``, l = 0 set to zero), and the batched inverse SHT makes
    the maps.  The small ``nmodes x nmodes`` solves run on the host; everything that scales with
    the map or with ``numz^2`` runs in the CUDA kernels.  Returns ``float64[numz, npix]``."""
    t = _dev.torch()
    numz = corr.shape[1]
    maxl = corr.shape[0] - 1
    if corr.shape[2] != numz:
        raise Exception("Correlation matrix is incorrect shape.")
    nmodes = len(constraints)
    f_ind = [int(c[0]) for c in constraints]
    L = maxl + 1
    cl = _dev.to_device(corr, t.float64)
    _, evecs = nputil.eigh_batched_device(cl)                      # columns ascending: the last nmodes are the largest
    trans = evecs[:, :, numz - nmodes:].transpose(1, 2).contiguous().cpu().numpy()   # [L, nmodes, numz]
    del evecs
    tmat = trans[:, :, f_ind]                                      # [L, nmodes, nmodes]
    # W_l = trans_l^T inv(tmat_l^T): cv[:, (l, m)] = W_l cmap[(l, m), :]
    W = np.zeros((L, numz, numz))
    for l in range(1, L):
        W[l, :, :nmodes] = np.linalg.solve(tmat[l], trans[l]).T     # (inv(tmat) trans)^T = trans^T inv(tmat^T)
    cmaps = _dev.to_device(np.ascontiguousarray([np.asarray(c[1], dtype=np.float64) for c in constraints]), t.float64)
    cpanel = hputil.map2alm_device(cmaps, nside, maxl, iter=3)     # PANEL [nalm, nmodes]
    cdense = hputil.panel_to_dense(cpanel, maxl, nmodes)           # [nmodes, L, L]
    gauss = _dev.zeros((L, numz, L), t.complex128)
    gauss[:, :nmodes, :] = cdense.permute(1, 0, 2)
    nalm = L * (L + 1) // 2
    panel = _dev.empty((nalm, numz), t.complex128)
    draw_apply_device(_dev.to_device(W, t.float64), np.arange(L), None, numz, maxl, panel, gauss=gauss)
    sky = hputil.alm2map_device(panel, nside, maxl, _lib.ALM_PANEL, numz, numz)
    return sky if device_out else _dev.to_host(sky)
