"""ConstrainedGalaxy and the healpy-style helpers around it on the GPU path (cora/foreground/galaxy.py:43-345,
cora/util/hputil.py:534-604): smoothing = map2alm -> beam -> alm2map, coordinate rotation, the constrained realisation.
The reference's data file (skydata.npz) is not in its checkout: the tests use synthetic Haslam / spectral / Faraday maps."""

import numpy as np
import pytest

from oracle import sht as osht
from oracle import skysim as osk

pytestmark = pytest.mark.gpu


def _oracle_smoothing(m, sigma, nside):
    lmax = 3 * nside - 1
    alm = osht.map2alm(m[np.newaxis], nside, lmax, iter=3)
    ell = np.arange(lmax + 1)
    bl = np.exp(-0.5 * ell * (ell + 1) * sigma**2)
    for mm in range(lmax + 1):
        i0 = osht.alm_index(lmax, mm, mm)
        alm[:, i0 : i0 + lmax - mm + 1] *= bl[mm:]
    return osht.alm2map(alm, nside, lmax)[0]


def test_smoothing_vs_oracle_and_invariants():
    from cora_b200 import hputil

    nside = 8
    rng = np.random.default_rng(4)
    m = rng.standard_normal((3, 12 * nside * nside))
    sig = np.radians([2.0, 9.0, 0.0])
    got = hputil.smoothing(m, sigma=sig)
    for c in range(3):
        want = _oracle_smoothing(m[c], sig[c], nside)
        assert np.max(np.abs(got[c] - want)) / np.max(np.abs(want)) < 1e-10
    one = hputil.smoothing(m[0], fwhm=np.radians(2.0) * np.sqrt(8 * np.log(2)))
    np.testing.assert_allclose(one, got[0], rtol=0, atol=1e-13 * np.abs(got[0]).max())
    # a constant map passes (almost) unchanged: l = 0 is untouched, the nside-8 quadrature leaks 4e-4 of it into high l
    np.testing.assert_allclose(hputil.smoothing(np.full(768, 3.25), sigma=0.3), 3.25, rtol=2e-3)
    assert got[1].std() < got[0].std() < m[0].std()


def _synthetic_data(nside_data=32, seed=0):
    from cora_b200 import healpix

    rng = np.random.default_rng(seed)
    th, ph = healpix.pix2ang(nside_data)
    b = np.pi / 2 - th
    haslam = 20.0 + 60.0 * np.exp(-(b / 0.35) ** 2) * (1.0 + 0.3 * np.cos(ph)) + rng.uniform(0.0, 2.0, th.size)
    spec = -2.8 + 0.1 * np.sin(b) + 0.02 * rng.standard_normal(th.size)
    faraday = 40.0 * np.exp(-(b / 0.5) ** 2) * np.sin(ph) + 5.0 * rng.standard_normal(th.size)
    return {"haslam": haslam, "spectral_md": spec, "spectral_gsm": spec + 0.05, "spectral_gd": spec - 0.05, "faraday": faraday}


def test_constrained_galaxy_getsky_against_oracle_pipeline():
    """The GPU stages inside ConstrainedGalaxy.getsky (smoothing, mkconstrained) reproduce the oracle's on the same
    Gaussian realisation; the output is positive and carries the Haslam map's large-scale structure."""
    from cora_b200 import galaxy, healpix, hputil

    class Small(galaxy.ConstrainedGalaxy):
        _amp_nside = 32

    nside = 32                               # (the class needs nside > 16: it takes variances inside nside-16 pixels)
    cg = Small(_synthetic_data())
    cg.nside = nside
    cg.frequencies = np.array([600.0, 500.0, 400.0])
    np.random.seed(3)
    fgt, fg, fgs, fgsmooth, am, mv = cg.getsky(debug=True, celestial=False)
    npix = 12 * nside * nside
    assert fgt.shape == (3, npix) and fg.shape == (5, npix) and np.isfinite(fgt).all()
    assert fgt.min() > 0.0
    # the constrained maps reproduce the smoothed 408 MHz fluctuation map at the constraint frequency
    sub408 = hputil.smoothing(fg[0], fwhm=np.radians(1.0))
    # (up to what map2alm with three refinements recovers of a band-limited map at lmax = 3 nside - 1)
    assert np.max(np.abs(fgs[0] - sub408)) / np.max(np.abs(sub408)) < 1e-2
    # oracle pipeline on the same realisation
    efreq = np.concatenate(([408.0, 1420.0], cg.frequencies))
    from oracle import spectra as osp

    cla = osk.clarray(osp.full_sky_synchrotron().angular_powerspectrum, 3 * nside - 1, efreq, zromb=0)
    o408 = _oracle_smoothing(fg[0], np.radians(1.0) / np.sqrt(8 * np.log(2)), nside)
    ofgs = osk.mkconstrained(cla, [(0, o408)], nside)
    assert np.max(np.abs(fgs - ofgs)) / np.max(np.abs(ofgs)) < 1e-7
    haslam = healpix.ud_grade(cg._haslam, nside)
    sc = healpix.ud_grade(cg._sp_ind["md"], nside)
    osmooth = haslam[None, :] * ((efreq / 408.0)[:, None] ** sc)
    np.testing.assert_allclose(fgsmooth, osmooth, rtol=1e-13)
    x = (am / mv) * (fg - ofgs) / osmooth
    want = ((np.where(x < 0, np.tanh(x), x) + 1) * osmooth)[2:]
    assert np.max(np.abs(fgt - want)) / np.max(np.abs(want)) < 1e-7
    # celestial output = the Galactic one rotated
    np.random.seed(3)
    cel = cg.getsky(celestial=True)
    assert cel.shape == (3, npix) and np.isfinite(cel).all() and cel.min() > 0.0


def test_constrained_galaxy_getpolsky_small():
    from cora_b200 import galaxy

    class Small(galaxy.ConstrainedGalaxy):
        _amp_nside = 32
        _maxphi = 12.0
        _dphi = 1.0

    nside = 32
    cg = Small(_synthetic_data(seed=2))
    cg.nside = nside
    cg.frequencies = np.array([700.0, 650.0, 600.0, 550.0])
    np.random.seed(11)
    sky = cg.getpolsky(celestial=False)
    assert sky.shape == (4, 4, 12 * nside * nside) and np.isfinite(sky).all()
    assert np.all(sky[:, 3] == 0.0) and sky[:, 0].min() > 0.0
    # |P| = I tanh(|.|) < I
    assert np.all(sky[:, 1] ** 2 + sky[:, 2] ** 2 <= sky[:, 0] ** 2 * (1 + 1e-12))
    assert np.abs(sky[:, 1]).max() > 0.0


def test_map_variance_and_chunk_var():
    from cora_b200 import galaxy, healpix

    rng = np.random.default_rng(0)
    m = rng.standard_normal(12 * 8 * 8)
    v = galaxy.map_variance(m, 2)
    nest = healpix.reorder(m, r2n=True).reshape(48, 16)
    np.testing.assert_allclose(healpix.reorder(v, r2n=True), nest.var(axis=1), rtol=1e-14)
    z = rng.standard_normal(1000) + 1j * rng.standard_normal(1000)
    assert galaxy.chunk_var(z) == pytest.approx(np.var(z), rel=1e-12)
