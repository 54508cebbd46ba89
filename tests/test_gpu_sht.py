"""GPU parity: batched inverse SHT (csrc/sht.cu through the C ABI) vs the oracle restatement.
Tolerance: float64, max-abs over pixels relative to the map's max <= 1e-10 (north star)."""

import numpy as np
import pytest

from oracle import hputil as ohp
from oracle import sht as osht

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _rand_alm(rng, nchan, lmax, spectrum=True):
    nalm = (lmax + 1) * (lmax + 2) // 2
    a = rng.standard_normal((nchan, nalm)) + 1j * rng.standard_normal((nchan, nalm))
    if spectrum:  # red spectrum, like the sky models
        l = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
        a = a * (1.0 + l) ** -1.2
    return a


def _relerr(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("nside,lmax,nchan", [(1, 2, 1), (2, 5, 3), (4, 11, 16), (8, 23, 5), (8, 24, 33), (16, 47, 2),
                                              (32, 95, 17), (3, 8, 4), (12, 35, 6)])
def test_alm2map_packed_vs_oracle(nside, lmax, nchan):
    import torch
    from cora_b200 import _lib, hputil

    rng = np.random.default_rng(nside * 1000 + lmax)
    alm = _rand_alm(rng, nchan, lmax)
    ref = osht.alm2map(alm, nside, lmax)
    d = torch.from_numpy(alm).cuda()
    out = hputil.alm2map_device(d, nside, lmax, _lib.ALM_PACKED, alm.shape[1], nchan).cpu().numpy()
    assert out.shape == ref.shape
    assert _relerr(out, ref) < TOL


def test_alm2map_config1_shape():
    """BASELINE config 1: nside 64, 32 channels, lmax = 3 nside (gaussianfg CLI) and 3 nside - 1."""
    import torch
    from cora_b200 import _lib, hputil

    for lmax in (191, 192):
        rng = np.random.default_rng(lmax)
        alm = _rand_alm(rng, 32, lmax)
        ref = osht.alm2map(alm, 64, lmax)
        out = hputil.alm2map_device(torch.from_numpy(alm).cuda(), 64, lmax, _lib.ALM_PACKED, alm.shape[1], 32).cpu().numpy()
        assert _relerr(out, ref) < TOL


def test_alm2map_small_workspace_batches():
    """A workspace that only fits a few channels must give the same maps (channel batching)."""
    import torch
    from cora_b200 import _dev, _lib

    nside, lmax, nchan = 8, 23, 21
    rng = np.random.default_rng(3)
    alm = _rand_alm(rng, nchan, lmax)
    ref = osht.alm2map(alm, nside, lmax)
    plan = _dev.sht_plan(nside, lmax)
    lib = _lib.load()
    nbytes = lib.cora_b200_alm2map_workspace_bytes(plan, _lib.ALM_PACKED, 4)
    ws = _dev.workspace(nbytes)
    d = torch.from_numpy(alm).cuda()
    out = torch.empty((nchan, 12 * nside * nside), dtype=torch.float64, device="cuda")
    _lib.call("cora_b200_alm2map", plan, _lib.ptr(d), _lib.ALM_PACKED, alm.shape[1], nchan, _lib.ptr(out), _lib.ptr(ws),
              int(nbytes), _lib.stream_ptr())
    assert _relerr(out.cpu().numpy(), ref) < TOL
    # too small for even one channel -> error code, not a crash
    with pytest.raises(_lib.CoraB200Error, match="workspace too small"):
        _lib.call("cora_b200_alm2map", plan, _lib.ptr(d), _lib.ALM_PACKED, alm.shape[1], nchan, _lib.ptr(out), _lib.ptr(ws),
                  1024, _lib.stream_ptr())


def test_monopole_and_m0_imag_ignored():
    from cora_b200 import hputil

    nside, lmax = 16, 47
    a = np.zeros((lmax + 1, lmax + 1), dtype=complex)
    a[0, 0] = np.sqrt(4 * np.pi)
    np.testing.assert_allclose(hputil.sphtrans_inv_real(a, nside), 1.0, rtol=1e-13)
    rng = np.random.default_rng(0)
    a = np.tril(rng.standard_normal((lmax + 1, lmax + 1)) + 1j * rng.standard_normal((lmax + 1, lmax + 1)))
    b = a.copy()
    b[:, 0] = b[:, 0].real
    np.testing.assert_array_equal(hputil.sphtrans_inv_real(a, nside), hputil.sphtrans_inv_real(b, nside))


def test_sphtrans_inv_real_errors():
    from cora_b200 import hputil

    with pytest.raises(Exception, match="a_lm array wrong shape"):
        hputil.sphtrans_inv_real(np.zeros((4, 5), dtype=complex), 2)
    with pytest.raises(Exception, match="a_lm array wrong shape"):
        hputil.sphtrans_inv_real_pol(np.zeros((2, 4, 4), dtype=complex), 2)


@pytest.mark.parametrize("nside,lmax,nchan", [(2, 5, 1), (4, 12, 3), (8, 23, 9), (16, 48, 4), (32, 95, 2), (6, 17, 2)])
def test_spin2_vs_oracle(nside, lmax, nchan):
    import torch
    from cora_b200 import _lib, hputil

    rng = np.random.default_rng(77 + lmax)
    aE = _rand_alm(rng, nchan, lmax)
    aB = _rand_alm(rng, nchan, lmax)
    Q, U = osht.alm2map_spin2(aE, aB, nside, lmax)
    q, u = hputil.alm2map_spin2_device(torch.from_numpy(aE).cuda(), torch.from_numpy(aB).cuda(), nside, lmax,
                                       _lib.ALM_PACKED, aE.shape[1], nchan)
    scale = max(np.abs(Q).max(), np.abs(U).max())
    assert np.max(np.abs(q.cpu().numpy() - Q)) / scale < TOL
    assert np.max(np.abs(u.cpu().numpy() - U)) / scale < TOL


@pytest.mark.parametrize("npol", [1, 2, 3, 4])
def test_sphtrans_inv_sky_vs_oracle(npol):
    from cora_b200 import hputil

    nside, lmax, nfreq = 8, 24, 3
    rng = np.random.default_rng(npol)
    L = lmax + 1
    alm = np.tril(rng.standard_normal((nfreq, npol, L, L)) + 1j * rng.standard_normal((nfreq, npol, L, L)))
    ref = ohp.sphtrans_inv_sky(alm, nside)
    out = hputil.sphtrans_inv_sky(alm, nside)
    assert out.shape == (nfreq, npol, 12 * nside * nside)
    assert np.max(np.abs(out - ref)) / np.max(np.abs(ref)) < TOL


def test_alm2map_nside256_property():
    """Full-size geometry (config 2 resolution): lmax = 767, a few channels.  Size-independent
    checks: monopole -> constant; linearity; north/south mirror of an even-parity mode; a single
    high-(l, m) mode against scipy on a sample of rings."""
    import torch
    from scipy.special import sph_harm_y
    from cora_b200 import _lib, hputil

    nside, lmax = 256, 767
    nalm = (lmax + 1) * (lmax + 2) // 2
    rng = np.random.default_rng(5)
    alm = np.zeros((4, nalm), dtype=complex)
    alm[0, 0] = np.sqrt(4 * np.pi)
    l1, m1 = 700, 650
    alm[1, osht.alm_index(lmax, l1, m1)] = 0.3 - 0.8j
    alm[2] = _rand_alm(rng, 1, lmax)[0]
    alm[3] = 2.0 * alm[1] - 0.5 * alm[2]
    out = hputil.alm2map_device(torch.from_numpy(alm).cuda(), nside, lmax, _lib.ALM_PACKED, nalm, 4).cpu().numpy()
    np.testing.assert_allclose(out[0], 1.0, rtol=1e-12)
    scale = np.abs(out[3]).max()
    assert np.max(np.abs(out[3] - (2.0 * out[1] - 0.5 * out[2]))) / scale < 1e-12
    g = osht.ring_geometry(nside)
    lam = osht.lambda_lm(lmax, m1, g["cth"], g["sth"])[l1 - m1]  # scipy's Y_lm returns NaN at this (l, m)
    th_eq = np.arctan2(g["sth"][511], g["cth"][511])
    lam_lo = osht.lambda_lm(lmax, 40, g["cth"][511:512], g["sth"][511:512])[300 - 40, 0]
    assert abs(lam_lo - sph_harm_y(300, 40, th_eq, 0.0).real) < 1e-12  # the recurrence itself is pinned to scipy
    for r in (0, 5, 100, 255, 256, 400, 511, 700, 1022):
        s, n = int(g["start"][r]), int(g["nph"][r])
        ph = g["phi0"][r] + 2 * np.pi * np.arange(n) / n
        ref = 2.0 * ((0.3 - 0.8j) * lam[r] * np.exp(1j * m1 * ph)).real
        assert np.max(np.abs(out[1, s : s + n] - ref)) < 1e-10
