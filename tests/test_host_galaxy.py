"""Host logic of ``galaxy.ConstrainedGalaxy`` on the CPU: every GPU call it makes (``skysim.clarray / mkfullsky /
mkconstrained``, ``hputil.smoothing / sphtrans_inv_real``) is replaced by the oracle's CPU counterpart, so the numpy
code between them -- template, amplitude and variance maps, constraint bookkeeping, Faraday-depth cube, rotation to
celestial coordinates -- runs without a device.  Pinned by ``tests/golden/constrained_galaxy_host.npz``, recorded with
this same harness from the implementation whose GPU stages were checked against the oracle pipeline on a B200
(``tests/test_gpu_galaxy.py``) before the class was restructured."""

import os

import numpy as np
import pytest

from oracle import skysim as osk
from oracle import sht as osht
from oracle import spectra as osp

GOLD = os.path.join(os.path.dirname(__file__), "golden", "constrained_galaxy_host.npz")


def _oracle_smoothing(hpmap, fwhm=0.0, sigma=None, iter=3, lmax=None):
    m = np.asarray(hpmap, dtype=np.float64)
    single = m.ndim == 1
    m2 = m.reshape(-1, m.shape[-1])
    nside = int(round(np.sqrt(m2.shape[1] / 12)))
    lmax = 3 * nside - 1 if lmax is None else lmax
    sig = np.broadcast_to(np.asarray(fwhm) / np.sqrt(8 * np.log(2)) if sigma is None else np.asarray(sigma), (m2.shape[0],))
    alm = osht.map2alm(m2, nside, lmax, iter=iter)
    ell = np.arange(lmax + 1)
    for c in range(m2.shape[0]):
        bl = np.exp(-0.5 * ell * (ell + 1) * sig[c] ** 2)
        for mm in range(lmax + 1):
            i0 = osht.alm_index(lmax, mm, mm)
            alm[c, i0 : i0 + lmax - mm + 1] *= bl[mm:]
    out = osht.alm2map(alm, nside, lmax)
    return out[0] if single else out.reshape(m.shape)


def _install(monkeypatch, seed):
    from cora_b200 import hputil, skysim

    rng = np.random.default_rng(seed)
    monkeypatch.setattr(hputil, "smoothing", _oracle_smoothing)
    monkeypatch.setattr(skysim, "clarray", lambda aps, lmax, z, zromb=3, **kw: osk.clarray(
        osp.full_sky_synchrotron().angular_powerspectrum, lmax, z, zromb=zromb))
    monkeypatch.setattr(skysim, "mkfullsky", lambda cla, nside, **kw: osk.mkfullsky(cla, nside, rng=rng))
    monkeypatch.setattr(skysim, "mkconstrained", lambda cla, cons, nside, **kw: osk.mkconstrained(cla, cons, nside))
    monkeypatch.setattr(hputil, "sphtrans_inv_real", lambda alm, nside: osht.alm2map(hputil.pack_alm(alm)[np.newaxis], nside)[0])


def _synthetic_data(nside_data=32, seed=0):
    from cora_b200 import healpix

    rng = np.random.default_rng(seed)
    th, ph = healpix.pix2ang(nside_data)
    b = np.pi / 2 - th
    haslam = 20.0 + 60.0 * np.exp(-(b / 0.35) ** 2) * (1.0 + 0.3 * np.cos(ph)) + rng.uniform(0.0, 2.0, th.size)
    spec = -2.8 + 0.1 * np.sin(b) + 0.02 * rng.standard_normal(th.size)
    faraday = 40.0 * np.exp(-(b / 0.5) ** 2) * np.sin(ph) + 5.0 * rng.standard_normal(th.size)
    return {"haslam": haslam, "spectral_md": spec, "spectral_gsm": spec + 0.05, "spectral_gd": spec - 0.05, "faraday": faraday}


def _check(name, got):
    g = np.load(GOLD)
    sub, sums = g[name + "_sub"], g[name + "_sum"]
    scale = np.abs(sub).max()
    assert np.max(np.abs(got[..., ::37] - sub)) / scale < 1e-10
    np.testing.assert_allclose([got.sum(), np.abs(got).sum(), (got * got).sum()], sums, rtol=1e-10)


def test_constrained_galaxy_host_logic(monkeypatch):
    from cora_b200 import galaxy

    class Small(galaxy.ConstrainedGalaxy):
        _amp_nside = 32
        _maxphi = 6.0
        _dphi = 1.0

    freqs3 = np.array([700.0, 600.0, 500.0])
    for smap in ("md", "gsm"):
        _install(monkeypatch, 5)
        cg = Small(_synthetic_data())
        cg.spectral_map, cg.nside, cg.frequencies = smap, 32, freqs3
        sky = cg.getsky(celestial=False)
        assert sky.shape == (3, 12 * 32 * 32) and sky.min() > 0.0
        _check("sky_gal_" + smap, sky)
    _install(monkeypatch, 5)
    cg = Small(_synthetic_data())
    cg.nside, cg.frequencies = 32, freqs3
    _check("sky_cel_md", cg.getsky(celestial=True))
    _install(monkeypatch, 7)
    np.random.seed(21)
    cg = Small(_synthetic_data(seed=2))
    cg.nside, cg.frequencies = 32, np.array([700.0, 650.0, 600.0, 550.0])
    pol = cg.getpolsky(celestial=True)
    assert pol.shape == (4, 4, 12 * 32 * 32) and np.all(pol[:, 3] == 0.0)
    assert np.all(pol[:, 1] ** 2 + pol[:, 2] ** 2 <= pol[:, 0] ** 2 * (1 + 1e-9) + 1e-12)
    _check("polsky_cel", pol)
    _check("amp_map", cg._amp_map)


def test_constrained_galaxy_debug_products(monkeypatch):
    """getsky(debug=True) returns (sky, fluctuations, constrained part, template, amplitude map, mean r.m.s.) and the
    pieces recombine into the sky."""
    from cora_b200 import galaxy

    class Small(galaxy.ConstrainedGalaxy):
        _amp_nside = 32

    _install(monkeypatch, 9)
    cg = Small(_synthetic_data())
    cg.nside, cg.frequencies = 32, np.array([650.0, 450.0])
    sky, fluct, large, template, amp, mean_rms = cg.getsky(debug=True, celestial=False)
    assert fluct.shape == large.shape == template.shape == (4, 12 * 32 * 32) and sky.shape[0] == 2
    x = (amp / mean_rms) * (fluct - large) / template
    np.testing.assert_allclose(sky, ((np.where(x < 0, np.tanh(x), x) + 1) * template)[2:], rtol=1e-13)
    np.testing.assert_allclose(template[0], cg._template(np.array([408.0]))[0], rtol=1e-14)
