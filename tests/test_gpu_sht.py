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


# ------------------------------------------------------------------ forward transform (map2alm)
def _panel_to_packed(panel):
    return panel.cpu().numpy().T.copy()   # [nchan, nalm], healpy order (PANEL rows are idx(l, m))


@pytest.mark.parametrize("nside,lmax,nchan", [(1, 2, 1), (2, 5, 3), (4, 11, 9), (8, 23, 5), (8, 16, 17), (16, 47, 2), (6, 17, 4),
                                               (32, 95, 3)])
def test_map2alm_quadrature_pass_vs_oracle(nside, lmax, nchan):
    """One analysis pass (iter=0) against the oracle's quadrature, 1e-12 relative (max-abs)."""
    import torch
    from cora_b200 import hputil

    rng = np.random.default_rng(nside * 100 + lmax)
    maps = rng.standard_normal((nchan, 12 * nside**2))
    got = _panel_to_packed(hputil.map2alm_device(torch.from_numpy(maps).cuda(), nside, lmax, iter=0))
    ref = osht.map2alm_adjoint(maps, nside, lmax)
    assert _relerr(got, ref) < 1e-12


def test_map2alm_iterations_and_weights_vs_oracle():
    import torch
    from cora_b200 import hputil

    rng = np.random.default_rng(11)
    nside, lmax, nchan = 8, 20, 6
    maps = rng.standard_normal((nchan, 12 * nside**2))
    w = 1.0 + 0.05 * rng.standard_normal(2 * nside)
    for it, rw in ((2, None), (3, w)):
        got = _panel_to_packed(hputil.map2alm_device(torch.from_numpy(maps).cuda(), nside, lmax, iter=it, ring_weights=rw))
        ref = osht.map2alm(maps, nside, lmax, iter=it, ring_weights=rw)
        assert _relerr(got, ref) < 1e-11


def test_map2alm_many_ring_blocks_nside256():
    """nside 256 (one 512-ring block) and nside 512-like block splitting are exercised through the
    adjoint identity at full size: <A f, a> = (4 pi / npix) <f, S a> with the GPU synthesis."""
    import torch
    from cora_b200 import _lib, hputil

    for nside, lmax, nchan in ((256, 767, 4), (320, 600, 2)):   # 320: 640 north rings -> two ring blocks (atomics)
        rng = np.random.default_rng(nside)
        npix = 12 * nside**2
        f = torch.from_numpy(rng.standard_normal((nchan, npix))).cuda()
        nalm = (lmax + 1) * (lmax + 2) // 2
        a = rng.standard_normal((nalm, nchan)) + 1j * rng.standard_normal((nalm, nchan))
        a[: lmax + 1] = a[: lmax + 1].real
        a_dev = torch.from_numpy(a).cuda()
        Af = hputil.map2alm_device(f, nside, lmax, iter=0).cpu().numpy()
        Sa = hputil.alm2map_device(a_dev, nside, lmax, _lib.ALM_PANEL, nchan, nchan).cpu().numpy()
        lhs = (f.cpu().numpy() * Sa).sum(axis=1) * 4.0 * np.pi / npix
        wm = np.full(nalm, 2.0)
        wm[: lmax + 1] = 1.0
        rhs = (wm[:, None] * (np.conj(a) * Af).real).sum(axis=0)
        np.testing.assert_allclose(lhs, rhs, rtol=1e-10)


def test_sphtrans_real_sky_and_sph_ps_vs_oracle():
    from cora_b200 import hputil

    rng = np.random.default_rng(12)
    nside = 8
    lmax = 3 * nside - 1
    m = rng.standard_normal(12 * nside**2)
    ref = ohp.unpack_alm(osht.map2alm(m, nside, lmax, iter=2), lmax) if hasattr(ohp, "unpack_alm") else None
    alm = hputil.sphtrans_real(m)
    assert alm.shape == (lmax + 1, lmax + 1)
    packed = hputil.pack_alm(alm)
    assert _relerr(packed, osht.map2alm(m, nside, lmax, iter=2)) < 1e-11
    if ref is not None:
        assert _relerr(alm, ref) < 1e-11
    big = hputil.sphtrans_real(m, lmax=10, lside=14)
    assert big.shape == (15, 15) and np.all(big[11:] == 0) and np.all(big[:, 11:] == 0)
    sky = rng.standard_normal((3, 12 * nside**2))
    alms = hputil.sphtrans_sky(sky)
    assert alms.shape == (3, lmax + 1, lmax + 1)
    assert _relerr(hputil.pack_alm(alms[1]), osht.map2alm(sky[1], nside, lmax, iter=2)) < 1e-11
    cl = hputil.sph_ps(m)
    assert _relerr(cl, osht.anafast(m, lmax=lmax, iter=2)) < 1e-10
    with pytest.raises(Exception, match="wrong shape"):
        hputil.sphtrans_sky(np.zeros((2, 2, 12 * nside**2)))
    with pytest.raises(ValueError):
        hputil.sphtrans_real(np.zeros(100))


def test_mkfullsky_reproduces_cl_via_anafast():
    """North-star check 2: GPU-RNG maps -> GPU anafast reproduce the input C_l(nu, nu') within
    cosmic variance (config 1 shape: gaussianfg nside 64, lmax 192, 8 of the 32 channels analysed)."""
    import torch
    from cora_b200 import galaxy, hputil, skysim

    nside, nfreq = 64, 32
    lmax = 3 * nside
    freq = np.linspace(800.0, 400.0, nfreq, endpoint=False)
    cl = skysim.clarray(galaxy.FullSkySynchrotron().angular_powerspectrum, lmax, freq)
    sky = skysim.mkfullsky(cl, nside, seed=2024, device_out=True)
    la = 2 * nside     # analysis band limit (HEALPix quadrature is accurate to ~2 nside)
    sel = [0, 5, 13, 31]
    alm = hputil.panel_to_dense(hputil.map2alm_device(sky[sel].contiguous(), nside, la, iter=3), la, len(sel)).cpu().numpy()
    for a_i, i in enumerate(sel):
        for b_i, j in enumerate(sel):
            prod = alm[a_i] * np.conj(alm[b_i])
            est = (prod[:, 0].real + 2.0 * prod[:, 1:].sum(axis=1).real) / (2.0 * np.arange(la + 1) + 1.0)
            l = np.arange(8, la + 1)
            # m = 0 quirk of the reference: E[est] = C_l (2l + 1/2)/(2l + 1)
            want = cl[l, i, j] * (2.0 * l + 0.5) / (2.0 * l + 1.0)
            sig = np.sqrt((cl[l, i, i] * cl[l, j, j] + cl[l, i, j] ** 2) / (2.0 * l + 1.0))
            assert np.all(np.abs(est[l] - want) < 6.0 * sig), (i, j)
            # and on average over l the estimator is unbiased at the few-percent level
            assert abs(np.mean((est[l] - want) / sig)) < 0.5


# ------------------------------------------------------------------ polarised analysis
@pytest.mark.parametrize("nside,lmax,nchan", [(2, 5, 1), (4, 11, 5), (8, 23, 3), (8, 14, 9), (16, 47, 2), (6, 17, 4)])
def test_map2alm_spin2_pass_vs_oracle(nside, lmax, nchan):
    import torch
    from cora_b200 import hputil

    rng = np.random.default_rng(nside * 31 + lmax)
    Q, U = rng.standard_normal((2, nchan, 12 * nside**2))
    pE, pB = hputil.map2alm_spin2_device(torch.from_numpy(Q).cuda(), torch.from_numpy(U).cuda(), nside, lmax, iter=0)
    rE, rB = osht.map2alm_spin2_adjoint(Q, U, nside, lmax)
    scale = max(np.abs(rE).max(), np.abs(rB).max())
    assert np.abs(_panel_to_packed(pE) - rE).max() / scale < 1e-12
    assert np.abs(_panel_to_packed(pB) - rB).max() / scale < 1e-12


def test_sphtrans_real_pol_and_sky_vs_oracle():
    from cora_b200 import hputil

    rng = np.random.default_rng(21)
    nside = 8
    lmax = 3 * nside - 1
    maps = rng.standard_normal((4, 12 * nside**2))
    alms = hputil.sphtrans_real_pol(maps)
    assert alms.shape == (4, lmax + 1, lmax + 1)
    rE, rB = osht.map2alm_spin2(maps[1], maps[2], nside, lmax, iter=2)
    assert _relerr(hputil.pack_alm(alms[0]), osht.map2alm(maps[0], nside, lmax, iter=2)) < 1e-11
    assert _relerr(hputil.pack_alm(alms[3]), osht.map2alm(maps[3], nside, lmax, iter=2)) < 1e-11
    scale = max(np.abs(rE).max(), np.abs(rB).max())
    assert np.abs(hputil.pack_alm(alms[1]) - rE[0]).max() / scale < 1e-11
    assert np.abs(hputil.pack_alm(alms[2]) - rB[0]).max() / scale < 1e-11
    sky = rng.standard_normal((2, 3, 12 * nside**2))
    a = hputil.sphtrans_sky(sky, lmax=12)
    assert a.shape == (2, 3, 13, 13)
    rE, rB = osht.map2alm_spin2(sky[1, 1], sky[1, 2], nside, 12, iter=2)
    assert np.abs(hputil.pack_alm(a[1, 1]) - rE[0]).max() / np.abs(rE).max() < 1e-11
    with pytest.raises(Exception):
        hputil.sphtrans_real_pol(maps[:2])


def test_pol_roundtrip_band_limited_gpu():
    """alm (T,E,B) -> maps (GPU synthesis) -> alm (GPU analysis, iter=3) for a band-limited field."""
    from cora_b200 import hputil

    rng = np.random.default_rng(22)
    nside, lmax = 16, 20
    L = lmax + 1
    alm = np.zeros((1, 3, L, L), dtype=np.complex128)
    for l in range(2, L):
        alm[0, :, l, : l + 1] = rng.standard_normal((3, l + 1)) + 1j * rng.standard_normal((3, l + 1))
        alm[0, :, l, 0] = alm[0, :, l, 0].real
    sky = hputil.sphtrans_inv_sky(alm, nside)
    old = hputil._iter
    try:
        hputil._iter = 3
        back = hputil.sphtrans_sky(sky, lmax=lmax)
    finally:
        hputil._iter = old
    assert np.abs(back - alm).max() / np.abs(alm).max() < 2e-3


def test_complex_field_roundtrip_and_full_size_adjoint():
    from cora_b200 import _lib, hputil
    import torch

    # sphtrans_inv_complex / sphtrans_complex on a band-limited complex field
    rng = np.random.default_rng(33)
    nside, lmax = 8, 10
    L = lmax + 1
    full = np.zeros((L, 2 * L - 1), dtype=np.complex128)
    for l in range(L):
        for m in range(-l, l + 1):
            full[l, m] = rng.standard_normal() + 1j * rng.standard_normal()
    field = hputil.sphtrans_inv_complex(full, nside)
    assert field.shape == (12 * nside**2,) and np.iscomplexobj(field)
    # the reference's definition, line by line (hputil.py:452-457), on the oracle's real transforms.  (It is not
    # the inverse of sphtrans_complex: the m = 0 imaginary parts are dropped and almi carries a sign.)
    almr = hputil._make_half_alm(full)
    almi = 1.0j * (full[:, :L] - almr)
    ref = osht.alm2map(hputil.pack_alm(almr), nside, lmax) + 1.0j * osht.alm2map(hputil.pack_alm(almi), nside, lmax)
    assert np.abs(field - ref).max() / np.abs(ref).max() < 1e-10
    # sphtrans_complex of a complex map = transforms of the real and imaginary parts in the full-m layout
    cm = rng.standard_normal(12 * nside**2) + 1j * rng.standard_normal(12 * nside**2)
    got = hputil.sphtrans_complex(cm, lmax=lmax)
    want = hputil._make_full_alm(ohp.unpack_alm(osht.map2alm(cm.real, nside, lmax, iter=2), lmax)) + 1.0j * hputil._make_full_alm(
        ohp.unpack_alm(osht.map2alm(cm.imag, nside, lmax, iter=2), lmax))
    assert got.shape == (L, 2 * L - 1) and np.abs(got - want).max() / np.abs(want).max() < 1e-10
    with pytest.raises(Exception, match="wrong shape"):
        hputil.sphtrans_inv_complex(np.zeros((4, 4)), nside)

    # nside 1024, lmax 3071 (config 5 geometry, 8 ring blocks in the analysis kernel): adjoint identity
    nside, lmax, nchan = 1024, 3071, 2
    npix = 12 * nside**2
    nalm = (lmax + 1) * (lmax + 2) // 2
    f = torch.from_numpy(rng.standard_normal((nchan, npix))).cuda()
    a = rng.standard_normal((nalm, nchan)) + 1j * rng.standard_normal((nalm, nchan))
    a[: lmax + 1] = a[: lmax + 1].real
    a_dev = torch.from_numpy(a).cuda()
    Af = hputil.map2alm_device(f, nside, lmax, iter=0).cpu().numpy()
    Sa = hputil.alm2map_device(a_dev, nside, lmax, _lib.ALM_PANEL, nchan, nchan).cpu().numpy()
    lhs = (f.cpu().numpy() * Sa).sum(axis=1) * 4.0 * np.pi / npix
    wm = np.full(nalm, 2.0)
    wm[: lmax + 1] = 1.0
    rhs = (wm[:, None] * (np.conj(a) * Af).real).sum(axis=0)
    np.testing.assert_allclose(lhs, rhs, rtol=1e-9)


# ------------------------------------------------------------------ production sizes: GPU vs oracle on ring subsets
# (VERDICT r01 weak #1).  The oracle evaluates the same direct-sum definition on ~36 rings (both poles, the cap/belt
# boundaries, the equator, random rings) with its l-major sweep; the GPU transforms everything.  Channels: the GPU
# batch is wider than one 32-channel block, the oracle checks the first and the last channel.  Tolerance 1e-10
# relative to the map's maximum (north star), for a red and for a white spectrum (the white one weights the
# high-l, high-m terms whose lambda_lm underflow near the poles and exercise the 2^-256 rescaling).
PROD = [(256, 767), (512, 1535), (1024, 3071)]


def _gpu_rand_alm(nchan, lmax, seed, white):
    """PANEL alm generated on the device (torch is plumbing); returns (panel CUDA [nalm, nchan], packed numpy of
    the first and last channel)."""
    import torch

    nalm = (lmax + 1) * (lmax + 2) // 2
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    panel = torch.view_as_complex(torch.randn((nalm, nchan, 2), dtype=torch.float64, device="cuda", generator=gen))
    if not white:
        l = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)]).astype(np.float64)
        panel = panel * torch.from_numpy((1.0 + l) ** -1.2).cuda()[:, None]
    panel = panel.contiguous()
    sel = panel[:, [0, nchan - 1]].cpu().numpy().T.copy()
    return panel, sel


def _ring_err(gpu_map, vals, start):
    """max |gpu - oracle| over the selected rings / max |oracle|, per channel (gpu_map: numpy [2, npix])."""
    num = max(np.max(np.abs(gpu_map[:, s : s + v.shape[1]] - v)) for v, s in zip(vals, start))
    den = max(np.max(np.abs(v)) for v in vals)
    return num / den


@pytest.mark.parametrize("white", [False, True])
@pytest.mark.parametrize("nside,lmax", PROD)
def test_alm2map_production_size_vs_oracle_rings(nside, lmax, white):
    from cora_b200 import _lib, hputil

    nchan = 34
    panel, sel = _gpu_rand_alm(nchan, lmax, 11 + nside, white)
    out = hputil.alm2map_device(panel, nside, lmax, _lib.ALM_PANEL, nchan, nchan)
    got = out[[0, nchan - 1]].cpu().numpy()
    del out, panel
    rs = osht.parity_rings(nside)
    vals, start = osht.alm2map_rings(sel, nside, lmax, rs)
    assert _ring_err(got, vals, start) < TOL


@pytest.mark.parametrize("white", [False, True])
@pytest.mark.parametrize("nside,lmax", PROD)
def test_spin2_production_size_vs_oracle_rings(nside, lmax, white):
    from cora_b200 import _lib, hputil

    nchan = 18
    pE, selE = _gpu_rand_alm(nchan, lmax, 21 + nside, white)
    pB, selB = _gpu_rand_alm(nchan, lmax, 31 + nside, white)
    q, u = hputil.alm2map_spin2_device(pE, pB, nside, lmax, _lib.ALM_PANEL, nchan, nchan)
    gq, gu = q[[0, nchan - 1]].cpu().numpy(), u[[0, nchan - 1]].cpu().numpy()
    del q, u, pE, pB
    rs = osht.parity_rings(nside)
    vq, vu, start = osht.alm2map_spin2_rings(selE, selB, nside, lmax, rs)
    scale = max(max(np.max(np.abs(v)) for v in vq), max(np.max(np.abs(v)) for v in vu))
    eq = max(np.max(np.abs(gq[:, s : s + v.shape[1]] - v)) for v, s in zip(vq, start)) / scale
    eu = max(np.max(np.abs(gu[:, s : s + v.shape[1]] - v)) for v, s in zip(vu, start)) / scale
    assert eq < TOL and eu < TOL


@pytest.mark.parametrize("nside,lmax", PROD)
def test_map2alm_production_size_vs_oracle_ms(nside, lmax):
    """One quadrature pass of the analysis at production size against the oracle for a handful of m (all l):
    m = 0, 1, a mid value, the cap-aliasing region and lmax."""
    import torch
    from cora_b200 import hputil

    nchan = 18
    npix = 12 * nside * nside
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5 + nside)
    maps = torch.randn((nchan, npix), dtype=torch.float64, device="cuda", generator=gen)
    panel = hputil.map2alm_device(maps, nside, lmax, iter=0)
    sel = maps[[0, nchan - 1]].cpu().numpy()
    got = panel[:, [0, nchan - 1]].cpu().numpy().T
    del maps, panel
    ms = [0, 1, lmax // 3, 2 * nside + 1, lmax - 1, lmax]
    ref = osht.map2alm_adjoint_ms(sel, nside, lmax, ms)
    scale = max(np.max(np.abs(v)) for v in ref.values())
    for m, v in ref.items():
        g = got[:, osht.alm_index(lmax, m, m) : osht.alm_index(lmax, lmax, m) + 1]
        assert np.max(np.abs(g - v)) / scale < TOL, m


def test_single_role_legendre_kernels_agree_with_warp_specialised():
    """Both Legendre implementations stay covered: the warp-specialised kernels (default) and the single-role
    kernel (cora_b200_set_legendre_ws(0)) give the same maps to rounding, scalar and spin 2."""
    import torch
    from cora_b200 import _lib, hputil

    lib = _lib.load()
    nside, lmax, nchan = 32, 95, 19
    rng = np.random.default_rng(5)
    aT, aE, aB = (torch.from_numpy(_rand_alm(rng, nchan, lmax)).cuda() for _ in range(3))
    old = lib.cora_b200_set_legendre_ws(3)
    try:
        t1 = hputil.alm2map_device(aT, nside, lmax, _lib.ALM_PACKED, aT.shape[1], nchan).clone()
        q1, u1 = (x.clone() for x in hputil.alm2map_spin2_device(aE, aB, nside, lmax, _lib.ALM_PACKED, aE.shape[1], nchan))
        lib.cora_b200_set_legendre_ws(0)
        t0 = hputil.alm2map_device(aT, nside, lmax, _lib.ALM_PACKED, aT.shape[1], nchan)
        q0, u0 = hputil.alm2map_spin2_device(aE, aB, nside, lmax, _lib.ALM_PACKED, aE.shape[1], nchan)
    finally:
        lib.cora_b200_set_legendre_ws(old)
    for a, b in ((t1, t0), (q1, q0), (u1, u0)):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-12
