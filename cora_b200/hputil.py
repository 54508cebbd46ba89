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


def alm2map_to_host(alm_panel, nside, lmax, nchan, nbatch=4):
    """Scalar synthesis of a PANEL alm array straight into a pinned host array: the channels are
    transformed in ``nbatch`` batches and every finished batch is copied back on a second
    stream while the next one is computed (the PCIe copy of the float64 maps takes longer than
    the transform itself)."""
    t = _dev.torch()
    npix = 12 * nside * nside
    plan = _dev.sht_plan(nside, lmax)
    host = t.empty((nchan, npix), dtype=t.float64, pin_memory=True)
    cb = max(16, -(-nchan // nbatch))
    cb += (-cb) % 16
    main, side = t.cuda.current_stream(), _dev.copy_stream()
    bufs = [_dev.empty((min(cb, nchan), npix), t.float64) for _ in range(2)]
    ws, nbytes = _dev.sht_workspace(plan, _lib.ALM_PANEL, min(cb, nchan))
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
