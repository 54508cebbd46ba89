"""HEALPix helpers: alm packing and the batched inverse spherical-harmonic transform.

Mirrors the inverse-transform half of ``cora/util/hputil.py`` (``pack_alm``/``unpack_alm``
``:93-152``, ``sphtrans_inv_real`` ``:369-391``, ``sphtrans_inv_real_pol`` ``:394-432``,
``sphtrans_inv_sky`` ``:500-531``).  ``healpy.alm2map`` is replaced by the CUDA kernels in
``csrc/sht.cu`` -- all frequency channels in one batched call instead of a Python loop.
"""

import numpy as np

from . import _dev, _lib


def unpack_alm(alm, lmax, fullm=False):
    """Healpix-packed a_lm -> 2-D [l, m] array (``hputil.py:93-121``).  Pure index shuffle."""
    almarray = np.zeros((lmax + 1, lmax + 1), dtype=alm.dtype)
    (almarray.T)[np.triu_indices(lmax + 1)] = alm
    if fullm:
        almarray = _make_full_alm(almarray)
    return almarray


def pack_alm(almarray, lmax=None):
    """2-D [l, m] a_lm -> Healpix packing, ``idx(l,m) = m(2 lmax+1-m)/2 + l`` (``hputil.py:124-152``)."""
    if (2 * almarray.shape[1] - 1) == almarray.shape[0]:
        almarray = _make_half_alm(almarray)
    if not lmax:
        lmax = almarray.shape[0] - 1
    return (almarray.T)[np.triu_indices(lmax + 1)]


def _make_full_alm(alm_half, centered=False):
    """[l, m>=0] -> [l, all m] via a_{l,-m} = (-1)^m conj(a_lm); negative m wrap to the end of
    the axis (index -m), or sit left of m=0 when ``centered`` (behaviour of ``hputil.py:155-175``)."""
    nl, nm = alm_half.shape[-2:]
    full = np.zeros(alm_half.shape[:-2] + (nl, 2 * nm - 1), dtype=alm_half.dtype)
    ms = np.arange(1, nm)
    neg = np.where(ms % 2 == 0, 1.0, -1.0) * np.conj(alm_half[..., ms])
    if centered:
        full[..., nm - 1 :] = alm_half
        full[..., nm - 1 - ms] = neg
    else:
        full[..., :nm] = alm_half
        full[..., 2 * nm - 1 - ms] = neg
    return full


def _make_half_alm(alm_full):
    """[l, all m] (wrapped) -> [l, m>=0], keeping the part consistent with a real field:
    a_lm <- (a_lm + (-1)^m conj(a_{l,-m})) / 2   (behaviour of ``hputil.py:177-193``)."""
    nl = alm_full.shape[-2]
    half = np.zeros(alm_full.shape[:-2] + (nl, nl), dtype=alm_full.dtype)
    half[..., 0] = alm_full[..., 0]
    ms = np.arange(1, nl)
    sign = np.where(ms % 2 == 0, 1.0, -1.0)
    half[..., ms] = 0.5 * (alm_full[..., ms] + sign * np.conj(alm_full[..., -ms]))
    return half


# ------------------------------------------------------------------ device-level SHT
def alm2map_device(alm_dev, nside, lmax, layout, alm_stride, nchan, out=None, stream=None, ws=None):
    """Scalar synthesis on device buffers.  ``alm_dev``: CUDA complex128 tensor in PACKED
    ([chan][nalm]) or PANEL ([nalm][stride]) layout; returns CUDA float64 [nchan, npix]."""
    t = _dev.torch()
    plan = _dev.sht_plan(nside, lmax)
    npix = 12 * nside * nside
    if out is None:
        out = _dev.empty((nchan, npix), t.float64)
    if ws is None:
        ws, nbytes = _dev.sht_workspace(plan, layout, nchan)
    else:
        nbytes = ws.numel()
    _lib.call("cora_b200_alm2map", plan, _lib.ptr(alm_dev), int(layout), int(alm_stride), int(nchan),
              _lib.ptr(out), _lib.ptr(ws), int(nbytes), _lib.stream_ptr(stream))
    return out


def alm2map_to_host(alm_panel, nside, lmax, nchan, nbatch=None, host=None, cache=None):
    """Scalar synthesis of a PANEL alm array straight into a pinned host array: the channels are
    transformed in ``nbatch`` batches and every finished batch is copied back on a second
    stream while the next one is computed (the PCIe copy of the float64 maps takes longer than
    the transform itself).  ``host``: a pinned ``float64[nchan, npix]`` tensor to fill (default: a new
    one); ``cache``: a dict that keeps the device staging buffers and the workspace between calls."""
    t = _dev.torch()
    npix = 12 * nside * nside
    plan = _dev.sht_plan(nside, lmax)
    if host is None:
        host = t.empty((nchan, npix), dtype=t.float64, pin_memory=True)
    if nbatch is None:
        # the call is bound by the PCIe copy, which can only start when the first batch exists: batches of ~64 channels
        # (measured at nside 512 x 1024 channels: 4 batches 720 ms per call, first maps ready 50 ms after the a_lm)
        nbatch = max(4, min(16, nchan // 64))
    cb = max(16, -(-nchan // nbatch))
    cb += (-cb) % 16
    main, side = t.cuda.current_stream(), _dev.copy_stream()
    cache = {} if cache is None else cache
    key = ("a2m_host", nside, lmax, min(cb, nchan))
    if key not in cache:
        bufs = [_dev.empty((min(cb, nchan), npix), t.float64) for _ in range(2)]
        cache[key] = (bufs,) + _dev.sht_workspace(plan, _lib.ALM_PANEL, min(cb, nchan))
    bufs, ws, nbytes = cache[key]
    freed = [None, None]
    stride = int(alm_panel.shape[1])
    for i, c0 in enumerate(range(0, nchan, cb)):
        n = min(cb, nchan - c0)
        buf = bufs[i & 1]
        if freed[i & 1] is not None:
            main.wait_event(freed[i & 1])          # the copy out of this buffer has finished
        _lib.call("cora_b200_alm2map", plan, _lib.ptr_off(alm_panel, 16 * c0), _lib.ALM_PANEL, stride, n, _lib.ptr(buf),
                  _lib.ptr(ws), int(nbytes), _lib.stream_ptr(main))
        done = t.cuda.Event()
        done.record(main)
        side.wait_event(done)
        with t.cuda.stream(side):
            host[c0 : c0 + n].copy_(buf[:n], non_blocking=True)
            ev = t.cuda.Event()
            ev.record(side)
        freed[i & 1] = ev
    side.synchronize()
    main.wait_stream(side)
    _dev.traffic["d2h"] += host.numel() * 8
    return host.numpy()


def alm2map_spin2_device(almE_dev, almB_dev, nside, lmax, layout, alm_stride, nchan, outQ=None, outU=None, stream=None):
    t = _dev.torch()
    plan = _dev.sht_plan(nside, lmax)
    npix = 12 * nside * nside
    if outQ is None:
        outQ = _dev.empty((nchan, npix), t.float64)
    if outU is None:
        outU = _dev.empty((nchan, npix), t.float64)
    ws, nbytes = _dev.sht_workspace(plan, layout, nchan, mult=2)
    _lib.call("cora_b200_alm2map_spin2", plan, _lib.ptr(almE_dev), _lib.ptr(almB_dev), int(layout), int(alm_stride),
              int(nchan), _lib.ptr(outQ), _lib.ptr(outU), _lib.ptr(ws), int(nbytes), _lib.stream_ptr(stream))
    return outQ, outU


def dense_to_panel(alm_dense_dev, lmax):
    """cora dense complex128 [nchan, L, L] (device) -> PANEL [nalm, nchan] (device)."""
    t = _dev.torch()
    nchan = alm_dense_dev.shape[0]
    nalm = (lmax + 1) * (lmax + 2) // 2
    panel = _dev.empty((nalm, nchan), t.complex128)
    _lib.call("cora_b200_alm_dense_to_panel", _lib.ptr(alm_dense_dev), int(nchan), int(lmax), _lib.ptr(panel),
              int(nchan), 0, _lib.stream_ptr())
    return panel


def panel_to_dense(panel_dev, lmax, nchan, chan0=0):
    t = _dev.torch()
    L = lmax + 1
    dense = _dev.empty((nchan, L, L), t.complex128)
    _lib.call("cora_b200_alm_panel_to_dense", _lib.ptr(panel_dev), int(panel_dev.shape[1]), int(chan0), int(nchan),
              int(lmax), _lib.ptr(dense), _lib.stream_ptr())
    return dense


# ------------------------------------------------------------------ reference API
def sphtrans_inv_real(alm, nside):
    """Inverse SHT of one real field, ``alm[l, m]`` (``hputil.py:369-391``)."""
    if alm.shape[1] != alm.shape[0]:
        raise Exception("a_lm array wrong shape.")
    t = _dev.torch()
    lmax = alm.shape[0] - 1
    dense = _dev.to_device(np.asarray(alm, dtype=np.complex128)[np.newaxis], t.complex128)
    panel = dense_to_panel(dense, lmax)
    m = alm2map_device(panel, nside, lmax, _lib.ALM_PANEL, 1, 1)
    return m[0].cpu().numpy()


def sphtrans_inv_real_pol(alm, nside):
    """Inverse SHT of a polarised field ``alm[npol, l, m]``, npol = 3 or 4 (``hputil.py:394-432``)."""
    npol = alm.shape[0]
    if alm.shape[1] != alm.shape[2] or not (npol == 3 or npol == 4):
        raise Exception("a_lm array wrong shape.")
    return sphtrans_inv_sky(np.asarray(alm)[np.newaxis], nside)[0]


def sphtrans_inv_sky(alm, nside, device_out=False):
    """``alm[freq, pol, l, m]`` -> ``skymaps[freq, pol, pixel]`` (``hputil.py:500-531``).

    Polarised (T scalar, (E,B)->(Q,U) spin-2, V scalar) iff ``npol >= 3``; otherwise only
    polarisation 0 is transformed, exactly like the reference (entries of the other
    polarisations are left at 0 where the reference leaves them uninitialised).
    ``alm`` may be a numpy array or a CUDA tensor; the result is numpy unless ``device_out``.
    """
    t = _dev.torch()
    if alm.ndim != 4 or alm.shape[2] != alm.shape[3]:
        raise Exception("a_lm array wrong shape.")
    nfreq, npol = alm.shape[0], alm.shape[1]
    if npol >= 3 and not (npol == 3 or npol == 4):
        raise Exception("a_lm array wrong shape.")
    lmax = alm.shape[2] - 1
    npix = 12 * nside * nside
    pol = npol >= 3
    a = _dev.to_device(alm, t.complex128)  # [nfreq, npol, L, L]
    sky = _dev.zeros((nfreq, npol, npix), t.float64)
    scal = [0, 3] if (pol and npol == 4) else [0]
    for p in scal:
        panel = dense_to_panel(a[:, p].contiguous(), lmax)
        m = alm2map_device(panel, nside, lmax, _lib.ALM_PANEL, nfreq, nfreq)
        sky[:, p] = m
        del panel, m
    if pol:
        pe = dense_to_panel(a[:, 1].contiguous(), lmax)
        pb = dense_to_panel(a[:, 2].contiguous(), lmax)
        q, u = alm2map_spin2_device(pe, pb, nside, lmax, _lib.ALM_PANEL, nfreq, nfreq)
        sky[:, 1] = q
        sky[:, 2] = u
    if device_out:
        return sky
    return _dev.to_host(sky)


# ------------------------------------------------------------------ forward transform
# cora calls healpy.map2alm with use_weights=True, iter=2 (``hputil.py:46-47``).  healpy's ring
# weights are data files of the healpy distribution (not available here), so every forward transform below
# is UNWEIGHTED unless the caller passes ``ring_weights`` (absolute weights of the 2*nside northern rings
# incl. the equator): results then differ from cora.util.hputil at the quadrature-error level that the
# ``iter`` Jacobi refinements leave (``mkconstrained`` is unaffected: healpy's default there is unweighted).
_iter = 2


def map2alm_device(maps_dev, nside, lmax=None, iter=None, ring_weights=None, panel=None):
    """Batched scalar analysis on device: CUDA float64 ``[nchan, npix]`` -> PANEL alm
    ``complex128[nalm, nchan]``.  One quadrature pass plus ``iter`` Jacobi refinements
    ``a += A(map - S a)`` (what ``healpy.map2alm(..., iter=iter)`` does)."""
    t = _dev.torch()
    nchan = int(maps_dev.shape[0])
    lmax = 3 * nside - 1 if lmax is None else int(lmax)
    iter = _iter if iter is None else int(iter)
    plan = _dev.sht_plan(nside, lmax)
    lib = _lib.load()
    nalm = (lmax + 1) * (lmax + 2) // 2
    if panel is None:
        panel = _dev.empty((nalm, nchan), t.complex128)
    rw = None
    if ring_weights is not None:
        rw = np.asarray(ring_weights, dtype=np.float64)
        if rw.shape != (2 * nside,):
            raise ValueError("ring_weights must have 2*nside entries")
        rw = _dev.to_device(rw, t.float64)
    need = max(lib.cora_b200_map2alm_workspace_bytes(plan, nchan), lib.cora_b200_alm2map_workspace_bytes(plan, _lib.ALM_PANEL, nchan))
    per16 = max(lib.cora_b200_map2alm_workspace_bytes(plan, 16), lib.cora_b200_alm2map_workspace_bytes(plan, _lib.ALM_PANEL, 16))
    nbytes = min(need, max(per16, _dev.free_bytes() - (2 << 30)))
    ws = _dev.workspace(nbytes)

    def analyse(src, accumulate):
        _lib.call("cora_b200_map2alm", plan, _lib.ptr(src), nchan, _lib.ptr(rw), int(accumulate), _lib.ptr(panel),
                  int(panel.shape[1]), 0, _lib.ptr(ws), int(nbytes), _lib.stream_ptr())

    analyse(maps_dev, 0)
    if iter > 0:
        res = _dev.empty(tuple(maps_dev.shape), t.float64)
        for _ in range(iter):
            _lib.call("cora_b200_alm2map", plan, _lib.ptr(panel), _lib.ALM_PANEL, int(panel.shape[1]), nchan, _lib.ptr(res),
                      _lib.ptr(ws), int(nbytes), _lib.stream_ptr())
            _lib.call("cora_b200_map_sub", _lib.ptr(maps_dev), _lib.ptr(res), int(maps_dev.numel()), _lib.ptr(res),
                      _lib.stream_ptr())
            analyse(res, 1)
    return panel


def map2alm_spin2_device(mapQ_dev, mapU_dev, nside, lmax=None, iter=None, ring_weights=None):
    """Batched polarised analysis on device: (Q, U) CUDA float64 ``[nchan, npix]`` -> PANEL
    (aE, aB) ``complex128[nalm, nchan]``; quadrature pass + ``iter`` Jacobi refinements."""
    t = _dev.torch()
    nchan = int(mapQ_dev.shape[0])
    lmax = 3 * nside - 1 if lmax is None else int(lmax)
    iter = _iter if iter is None else int(iter)
    plan = _dev.sht_plan(nside, lmax)
    lib = _lib.load()
    nalm = (lmax + 1) * (lmax + 2) // 2
    pE, pB = _dev.empty((nalm, nchan), t.complex128), _dev.empty((nalm, nchan), t.complex128)
    rw = None
    if ring_weights is not None:
        rw = np.asarray(ring_weights, dtype=np.float64)
        if rw.shape != (2 * nside,):
            raise ValueError("ring_weights must have 2*nside entries")
        rw = _dev.to_device(rw, t.float64)
    need = max(lib.cora_b200_map2alm_spin2_workspace_bytes(plan, nchan), 2 * lib.cora_b200_alm2map_workspace_bytes(plan, _lib.ALM_PANEL, nchan))
    per16 = max(lib.cora_b200_map2alm_spin2_workspace_bytes(plan, 16), 2 * lib.cora_b200_alm2map_workspace_bytes(plan, _lib.ALM_PANEL, 16))
    nbytes = min(need, max(per16, _dev.free_bytes() - (2 << 30)))
    ws = _dev.workspace(nbytes)

    def analyse(q, u, accumulate):
        _lib.call("cora_b200_map2alm_spin2", plan, _lib.ptr(q), _lib.ptr(u), nchan, _lib.ptr(rw), int(accumulate), _lib.ptr(pE),
                  _lib.ptr(pB), nchan, 0, _lib.ptr(ws), int(nbytes), _lib.stream_ptr())

    analyse(mapQ_dev, mapU_dev, 0)
    if iter > 0:
        rq, ru = _dev.empty(tuple(mapQ_dev.shape), t.float64), _dev.empty(tuple(mapU_dev.shape), t.float64)
        for _ in range(iter):
            _lib.call("cora_b200_alm2map_spin2", plan, _lib.ptr(pE), _lib.ptr(pB), _lib.ALM_PANEL, nchan, nchan, _lib.ptr(rq),
                      _lib.ptr(ru), _lib.ptr(ws), int(nbytes), _lib.stream_ptr())
            _lib.call("cora_b200_map_sub", _lib.ptr(mapQ_dev), _lib.ptr(rq), int(rq.numel()), _lib.ptr(rq), _lib.stream_ptr())
            _lib.call("cora_b200_map_sub", _lib.ptr(mapU_dev), _lib.ptr(ru), int(ru.numel()), _lib.ptr(ru), _lib.stream_ptr())
            analyse(rq, ru, 1)
    return pE, pB


def _nside_of(npix):
    nside = int(round(np.sqrt(npix / 12.0)))
    if 12 * nside * nside != npix:
        raise ValueError("Wrong pixel number (it is not 12*nside**2)")   # healpy.npix2nside's message
    return nside


def sphtrans_real(hpmap, lmax=None, lside=None, ring_weights=None):
    """Spherical harmonic transform of a real map -> ``alm[l, m]``, m >= 0 (``hputil.py:195-234``).
    Unweighted quadrature + 2 refinements unless ``ring_weights`` is given (see the note above ``_iter``)."""
    t = _dev.torch()
    hpmap = np.ascontiguousarray(hpmap, dtype=np.float64)
    nside = _nside_of(hpmap.size)
    if lmax is None:
        lmax = 3 * nside - 1
    if lside is None or lside < lmax:
        lside = lmax
    panel = map2alm_device(_dev.to_device(hpmap[np.newaxis], t.float64), nside, lmax, ring_weights=ring_weights)
    dense = panel_to_dense(panel, lmax, 1)[0].cpu().numpy()
    alm = np.zeros([lside + 1, lside + 1], dtype=np.complex128)
    alm[: lmax + 1, : lmax + 1] = dense
    return alm


def sphtrans_complex(hpmap, lmax=None, centered=False, lside=None):
    """Transform of a complex function (``hputil.py:237-271``): real and imaginary parts separately."""
    if lmax is None:
        lmax = 3 * _nside_of(hpmap.size) - 1
    alm = _make_full_alm(sphtrans_real(hpmap.real, lmax=lmax, lside=lside), centered=centered)
    alm = alm + 1.0j * _make_full_alm(sphtrans_real(hpmap.imag, lmax=lmax, lside=lside), centered=centered)
    return alm


def _pol_analysis(maps, nside, lmax, ring_weights):
    """maps: CUDA float64 [nfreq, npol (3 or 4), npix] -> dense alms CUDA complex128 [nfreq, npol, L, L]."""
    t = _dev.torch()
    nfreq, npol = int(maps.shape[0]), int(maps.shape[1])
    L = lmax + 1
    out = _dev.empty((nfreq, npol, L, L), t.complex128)
    for p in ([0, 3] if npol == 4 else [0]):    # T (and V): scalar transforms
        out[:, p] = panel_to_dense(map2alm_device(maps[:, p].contiguous(), nside, lmax, ring_weights=ring_weights), lmax, nfreq)
    pE, pB = map2alm_spin2_device(maps[:, 1].contiguous(), maps[:, 2].contiguous(), nside, lmax, ring_weights=ring_weights)
    out[:, 1] = panel_to_dense(pE, lmax, nfreq)
    out[:, 2] = panel_to_dense(pB, lmax, nfreq)
    return out


def sphtrans_real_pol(hpmaps, lmax=None, lside=None, ring_weights=None):
    """Transform of real T, Q, U (and optionally V) maps -> ``alms[npol, l, m]`` (T, E, B, V),
    m >= 0 (``hputil.py:274-323``)."""
    t = _dev.torch()
    hpmaps = np.ascontiguousarray(hpmaps, dtype=np.float64)
    npol = len(hpmaps)
    if npol not in (3, 4):
        raise Exception("hpmaps must hold 3 or 4 polarisations.")
    nside = _nside_of(hpmaps[0].size)
    if lmax is None:
        lmax = 3 * nside - 1
    if lside is None or lside < lmax:
        lside = lmax
    dense = _pol_analysis(_dev.to_device(hpmaps[np.newaxis], t.float64), nside, lmax, ring_weights)[0].cpu().numpy()
    alms = np.zeros([npol, lside + 1, lside + 1], dtype=np.complex128)
    alms[:, : lmax + 1, : lmax + 1] = dense
    return alms


def sphtrans_sky(skymap, lmax=None, device_out=False, ring_weights=None):
    """Transform a 3-D sky map channel by channel (``hputil.py:460-497``), all channels batched.

    ``skymap[freq, pixel]`` -> ``alms[freq, l, m]``; polarised if there are three dimensions and
    3 or 4 polarisations: ``skymap[freq, pol, pixel]`` -> ``alms[freq, pol, l, m]`` (T, E, B, V)."""
    t = _dev.torch()
    if len(skymap.shape) == 3 and skymap.shape[1] >= 3:
        if skymap.shape[1] > 4:
            raise Exception("skymap wrong shape.")
        nside = _nside_of(skymap.shape[-1])
        if lmax is None:
            lmax = 3 * nside - 1
        dense = _pol_analysis(_dev.to_device(skymap, t.float64), nside, lmax, ring_weights)
        return dense if device_out else _dev.to_host(dense)
    if len(skymap.shape) == 3:
        raise Exception("skymap wrong shape.")   # [freq, pol < 3, pix]: the reference hands a 2-D block to healpy and fails there
    nside = _nside_of(skymap.shape[-1])
    if lmax is None:
        lmax = 3 * nside - 1
    maps = _dev.to_device(skymap, t.float64)
    panel = map2alm_device(maps, nside, lmax, ring_weights=ring_weights)
    dense = panel_to_dense(panel, lmax, int(maps.shape[0]))
    return dense if device_out else _dev.to_host(dense)


def sph_ps(map1, map2=None, lmax=None):
    """Angular (cross) power spectrum of real maps (``hputil.py:607-619``):
    ``C_l = (a1_l0 conj(a2_l0) + 2 Re sum_{m>0} a1_lm conj(a2_lm)) / (2l + 1)``.
    (The reference tests ``if map is not None`` -- the builtin, always true -- so its auto-spectrum
    call dies inside ``sphtrans_real(None)``; here ``map2=None`` means the auto-spectrum.)"""
    lmax = lmax if lmax is not None else (3 * _nside_of(np.asarray(map1).size) - 1)
    alm1 = sphtrans_real(map1, lmax)
    alm2 = sphtrans_real(map2, lmax) if map2 is not None else alm1
    prod = alm1 * alm2.conj()
    s = prod[:, 0] + 2 * prod[:, 1:].sum(axis=1).real
    return s / (2.0 * np.arange(lmax + 1) + 1.0)


# ------------------------------------------------------------------ small helpers of the reference module
def ang_positions(nside):
    """(theta, phi) of every RING pixel, ``float64[npix, 2]`` (``hputil.py:53-73``; the HEALPix ring
    geometry is evaluated here instead of through ``healpy.pix2ang``)."""
    n = int(nside)
    npix = 12 * n * n
    angpos = np.empty([npix, 2], dtype=np.float64)
    ring = np.arange(1, 4 * n)
    north = np.minimum(ring, 4 * n - ring).astype(np.float64)
    cap = north < n
    z = np.where(cap, 1.0 - north**2 / (3.0 * n * n), 4.0 / 3.0 - 2.0 * north / (3.0 * n))
    z = np.where(ring > 2 * n, -z, z)
    nph = np.where(cap, 4 * north, 4 * n).astype(np.int64)
    shifted = cap | ((north.astype(np.int64) - n) % 2 == 0)
    start = 0
    for r in range(4 * n - 1):
        k = int(nph[r])
        angpos[start : start + k, 0] = np.arccos(z[r])
        angpos[start : start + k, 1] = (np.arange(k) + (0.5 if shifted[r] else 0.0)) * (2.0 * np.pi / k)
        start += k
    return angpos


def nside_for_lmax(lmax, accuracy_boost=1):
    """A power-of-two nside appropriate for a decomposition up to ``lmax`` (``hputil.py:76-90``)."""
    return int(2 ** (accuracy_boost + np.ceil(np.log((lmax + 1) / 3.0) / np.log(2.0))))


def sphtrans_complex_pol(hpmaps, lmax=None, centered=False, lside=None):
    """Transform of complex T, Q, U (V) maps: real and imaginary parts separately, full-m layout
    (``hputil.py:326-366``)."""
    hpmaps = np.asarray(hpmaps)
    if lmax is None:
        lmax = 3 * _nside_of(hpmaps[0].size) - 1
    alm = _make_full_alm(sphtrans_real_pol(hpmaps.real, lmax=lmax, lside=lside), centered=centered)
    alm = alm + 1.0j * _make_full_alm(sphtrans_real_pol(hpmaps.imag, lmax=lmax, lside=lside), centered=centered)
    return alm


def sphtrans_inv_complex(alm, nside):
    """Inverse transform onto a complex field; ``alm[l, all m]`` in the wrapped layout
    (``hputil.py:435-457``).  Mirrors the reference line by line, including that it is not an exact inverse
    of ``sphtrans_complex`` (the imaginary parts of the m = 0 column are dropped by the real transform and
    ``almi = +1j (alm - almr)`` carries a sign)."""
    if alm.shape[1] != (2 * alm.shape[0] - 1):
        raise Exception("a_lm array wrong shape: " + repr(alm.shape))
    almr = _make_half_alm(alm)
    almi = 1.0j * (alm[:, : almr.shape[1]] - almr)
    return sphtrans_inv_real(almr, nside) + 1.0j * sphtrans_inv_real(almi, nside)


# ------------------------------------------------------------------ smoothing and coordinate rotation
def smoothing(hpmap, fwhm=0.0, sigma=None, iter=3, lmax=None):
    """``healpy.smoothing`` with the defaults cora relies on (``cora/foreground/galaxy.py:101-104,165-185``): analysis
    without ring weights and ``iter`` refinements up to ``lmax = 3 nside - 1``, multiplication by the Gaussian beam
    ``exp(-l (l + 1) sigma^2 / 2)`` (``sigma = fwhm / sqrt(8 ln 2)``), synthesis.  ``hpmap``: ``[npix]`` or
    ``[nmap, npix]``; ``fwhm`` / ``sigma`` in radians, scalars or one per map.  Both transforms and the beam run on the
    GPU (``cora_b200_map2alm`` -> ``cora_b200_alm_scale_l`` -> ``cora_b200_alm2map``)."""
    t = _dev.torch()
    m = np.asarray(hpmap, dtype=np.float64)
    single = m.ndim == 1
    m2 = np.ascontiguousarray(m.reshape(-1, m.shape[-1]))
    nmap, npix = m2.shape
    nside = _nside_of(npix)
    lmax = 3 * nside - 1 if lmax is None else int(lmax)
    sig = np.broadcast_to(np.asarray(fwhm, dtype=np.float64) / np.sqrt(8.0 * np.log(2.0)) if sigma is None
                          else np.asarray(sigma, dtype=np.float64), (nmap,))
    ell = np.arange(lmax + 1, dtype=np.float64)
    beam = np.exp(-0.5 * ell[:, None] * (ell[:, None] + 1.0) * sig[None, :] ** 2)           # [L, nmap]
    panel = map2alm_device(_dev.to_device(m2, t.float64), nside, lmax, iter=iter)
    _lib.call("cora_b200_alm_scale_l", _lib.ptr(panel), int(panel.shape[1]), 0, nmap, lmax,
              _lib.ptr(_dev.to_device(np.ascontiguousarray(beam), t.float64)), _lib.stream_ptr())
    out = _dev.to_host(alm2map_device(panel, nside, lmax, _lib.ALM_PANEL, int(panel.shape[1]), nmap))
    return out[0] if single else out.reshape(m.shape)


def coord_x2y(map, x, y):
    """Rotate a map between coordinate systems (``hputil.py:534-566``): every output pixel takes the bilinear
    interpolation of the input map at its own direction rotated by ``Rotator(coord=[y, x])``.  Works on a series of
    maps (Healpix index last), in place like the reference.  'C' celestial, 'G' Galactic ('E' ecliptic is not
    provided)."""
    from . import healpix

    if x not in ["C", "G", "E"] or y not in ["C", "G", "E"]:
        raise Exception("Co-ordinate system invalid.")
    npix = map.shape[-1]
    angpos = ang_positions(_nside_of(npix))
    theta, phi = healpix.rotate_angles(angpos[:, 0], angpos[:, 1], y, x)
    pix, wgt = healpix.get_interp_weights(_nside_of(npix), theta, phi)
    map_flat = map.reshape((-1, npix))
    for i in range(map_flat.shape[0]):
        map_flat[i] = np.sum(map_flat[i][pix] * wgt, axis=0)
    return map_flat.reshape(map.shape)


def coord_g2c(map_):
    """Galactic -> celestial (``hputil.py:569-585``)."""
    return coord_x2y(map_, "G", "C")


def coord_c2g(map_):
    """Celestial -> Galactic (``hputil.py:588-604``)."""
    return coord_x2y(map_, "C", "G")
