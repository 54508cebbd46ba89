"""Pin the C_l oracle: the reference's own known-answer values (tests/test_corr.py) and
fixtures computed by the real reference (tests/golden/make_golden.py)."""

import numpy as np
import pytest

from conftest import golden
from oracle import skysim, spectra


def test_sck_reference_goldens():
    # /root/reference/tests/test_corr.py:34-57
    cr = spectra.full_sky_synchrotron()
    aps1 = cr.angular_powerspectrum(np.arange(1000), 800.0, 800.0)
    assert len(aps1) == 1000
    assert np.allclose(aps1.sum(), 75.47681191093129, rtol=1e-7)
    fa = np.linspace(400.0, 800.0, 64)
    aps2 = cr.angular_powerspectrum(np.arange(1000)[:, None, None], fa[None, :, None], fa[None, None, :])
    assert aps2.shape == (1000, 64, 64)
    assert np.allclose(aps2[400, 40, 40], 9.690708728692975e-06, rtol=1e-7)
    assert np.allclose(aps2[200, 10, 40], 0.00017630767166797886, rtol=1e-7)
    # stronger than the upstream rtol: bit-level agreement
    assert abs(aps1.sum() / 75.47681191093129 - 1) < 1e-15


def test_sck_clarray_fixture():
    g = golden("cl_sck.npz")
    freq, lmax = g["freq"], int(g["lmax"])
    cl = skysim.clarray(spectra.full_sky_synchrotron().angular_powerspectrum, lmax, freq)
    np.testing.assert_allclose(cl, g["cl"], rtol=1e-14, atol=0)
    clp = skysim.clarray(spectra.full_sky_polarised_synchrotron().angular_powerspectrum, lmax, freq)
    np.testing.assert_allclose(clp, g["cl_pol"], rtol=1e-14, atol=0)
    cl0 = skysim.clarray(spectra.full_sky_synchrotron().angular_powerspectrum, lmax, freq, zromb=0)
    np.testing.assert_allclose(cl0, g["cl_zromb0"], rtol=1e-14, atol=0)
    assert np.all(cl[0] == 0.0)  # C_0 = 0


def test_romberg_weights_closed_form():
    w = skysim.romberg_weights(3)
    ref = (4.0 / 14175.0) * np.array([1085, 5120, 1760, 5120, 2180, 5120, 1760, 5120, 1085.0]) / 8 * 8
    np.testing.assert_allclose(w, ref / ref.sum() * 8, rtol=1e-14)
    assert abs(w.sum() - 8.0) < 1e-13


def test_21cm_reference_goldens_planck2013():
    # /root/reference/tests/test_corr.py:7-31 -- generated under the Planck-2013 cosmology
    # (stale against the reference's current Planck-2018 defaults).
    cr = spectra.Corr21cm(cosmology=spectra.Cosmology.planck2013())
    aps1 = cr.angular_powerspectrum(np.arange(1000), 800.0, 800.0)
    assert np.allclose(aps1.sum(), 1.5963772205823096e-09, rtol=1e-7)
    fa = np.linspace(400.0, 800.0, 64)
    aps2 = cr.angular_powerspectrum(np.arange(1000)[:, None, None], fa[None, :, None], fa[None, None, :])
    assert np.allclose(aps2[400, 40, 40], 8.986790805379046e-13, rtol=1e-7)
    assert np.allclose(aps2[200, 10, 40], 1.1939298801340165e-18, rtol=1e-7)
    g = golden("cl_21cm.npz")
    np.testing.assert_allclose(aps1, g["p13_aps1"], rtol=1e-12)
    # far-off-diagonal entries are differences of large DCT terms: compare norm-wise
    diag = np.sqrt(np.abs(np.einsum("lii->li", aps2)))
    scale = (diag[:, :, None] * diag[:, None, :])[::37, ::7, ::5]
    assert np.max(np.abs(aps2[::37, ::7, ::5] - g["p13_aps2_sub"]) / scale) < 1e-12


def test_21cm_fixture_planck2018(oracle_corr21cm):
    cr = oracle_corr21cm
    g = golden("cl_21cm.npz")
    # host vectors
    z = g["vec_z"]
    chi, b, f, pf, D = cr.sample_vectors(z)
    np.testing.assert_allclose(chi, g["vec_chi"], rtol=1e-13)
    np.testing.assert_allclose(f, g["vec_f"], rtol=1e-14)
    np.testing.assert_allclose(D, g["vec_D"], rtol=1e-14)
    np.testing.assert_allclose(pf, g["vec_pf"], rtol=1e-14)
    # spline (+ extrapolation both sides)
    np.testing.assert_allclose(cr.ps_vv(g["ps_k"]), g["ps_vv"], rtol=1e-13)
    # DCT tables
    dd, dv, vv = cr.tables()
    ix = np.ix_(g["tab_rows"], g["tab_cols"])
    scale = np.abs(dd[g["tab_rows"]]).max(axis=1)[:, None]
    assert np.max(np.abs(dd[ix] - g["tab_dd"]) / scale) < 1e-14
    assert np.max(np.abs(dv[ix] - g["tab_dv"]) / scale) < 1e-14
    assert np.max(np.abs(vv[ix] - g["tab_vv"]) / scale) < 1e-14
    # spectra and Romberg-averaged tables
    aps1 = cr.angular_powerspectrum(np.arange(1000), 800.0, 800.0)
    np.testing.assert_allclose(aps1, g["p18_aps1"], rtol=1e-12)
    freq, lmax = g["freq"], int(g["lmax"])
    for key, zromb in (("p18_cl", 3), ("p18_cl_romb1", 1)):
        cl = skysim.clarray(cr.angular_powerspectrum, lmax, freq, zromb=zromb)
        ref = g[key]
        norm = np.sqrt(np.abs(np.einsum("lii->li", ref)))
        scale = norm[:, :, None] * norm[:, None, :] + 1e-300
        assert np.max(np.abs(cl - ref) / scale) < 1e-12


# ---------------------------------------------------------------- reference tests/test_cubicspline.py (LogInterpolater)
def _log_spline(x, y):
    from oracle import spectra

    return spectra.LogSpline(np.dstack((x, y))[0])


def test_logspline_constant_kat():
    """tests/test_cubicspline.py:31-48 (LogInterpolater branch): f(x) = 1 inter- and extrapolated."""
    p = _log_spline(np.arange(1, 8), np.ones(7))
    assert (p(np.asarray([0.025, 1, 2.5, 4, 5.55, 7.01, 19])) == 1).all()


def test_logspline_linear_kat():
    """tests/test_cubicspline.py:51-71: f(x) = 10 x at and between the knots (and below the first one)."""
    p = _log_spline(np.arange(1, 5), np.asarray([10, 20, 30, 40]))
    for x_, want in ((0.5, 5), (1, 10), (1.75, 17.5), (2, 20), (2.2, 22), (3, 30), (4, 40)):
        assert p(x_) == pytest.approx(want)


def test_logspline_random_knots_kat():
    """tests/test_cubicspline.py:74-85: the spline passes through its knots to 1e-13."""
    x = np.arange(1, 5)
    y = np.asarray([1.67, 1.99, 0.465, 0.234])
    p = _log_spline(x, y)
    for i, x_ in enumerate(x):
        assert p(x_) == pytest.approx(y[i], rel=1e-13)


def test_oracle_21cm_clarray_config2_rows_vs_reference(oracle_corr21cm):
    """The oracle against the real reference at config 2's channel axis (256 channels, 8 l rows;
    tests/golden/make_golden.py c2rows)."""
    from conftest import golden
    from oracle import skysim as osk

    g = golden("cl_21cm_c2rows.npz")
    rows = g["rows"]

    def aps(l, z1, z2):
        return oracle_corr21cm.angular_powerspectrum(rows[np.asarray(l)].astype(np.float64), z1, z2)

    cl = osk.clarray(aps, len(rows) - 1, g["freq"])
    d = np.sqrt(np.abs(np.einsum("lii->li", cl)))
    scale = d[:, g["chan_rows"], None] * d[:, None, :] + 1e-300
    assert np.max(np.abs(cl[:, g["chan_rows"], :] - g["cl_rows"]) / scale) < 1e-12
