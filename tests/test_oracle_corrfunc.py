"""Oracle of the xi(r) -> C_l front end (oracle/corrfunc.py) against the fixture produced by the reference's own
``corr_to_clarray`` / ``legendre_array`` (tests/golden/make_golden.py corr_to_clarray)."""

import os

import numpy as np
import pytest

from oracle import corrfunc as ocf

GOLD = os.path.join(os.path.dirname(__file__), "golden", "corr_to_clarray.npz")


def corr_test_function(r):
    # the function the fixture was generated with (tests/golden/make_golden.py)
    return np.exp(-((r / 60.0) ** 2)) / (1.0 + (r / 15.0) ** 2) + 0.05 * np.cos(r / 35.0) * np.exp(-r / 400.0)


@pytest.mark.parametrize("key,kw", [("cl_romb2_q2", dict(xromb=2, q=2, chunksize=16)),
                                    ("cl_romb0_q3", dict(xromb=0, q=3, chunksize=50)),
                                    ("cl_romb1_w40", dict(xromb=1, xwidth=40.0, q=2, chunksize=20))])
def test_corr_to_clarray_matches_reference(key, kw):
    g = np.load(GOLD)
    got = ocf.corr_to_clarray(corr_test_function, int(g["lmax"]), g["xarray"], **kw)
    want = g[key]
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13 * np.abs(want).max())


def test_legendre_array_matches_reference():
    g = np.load(GOLD)
    np.testing.assert_allclose(ocf.legendre_array(60, g["leg_mu"]), g["leg"], rtol=1e-13, atol=1e-15)


def test_cosine_rule_is_the_law_of_cosines():
    mu = np.array([-1.0, 0.0, 0.5, 1.0])
    x = np.array([3.0, 4.0])
    r = ocf.cosine_rule(mu, x, x)
    assert r.shape == (4, 2, 2)
    np.testing.assert_allclose(r[1], [[np.sqrt(18.0), 5.0], [5.0, np.sqrt(32.0)]])
    np.testing.assert_allclose(r[0], [[6.0, 7.0], [7.0, 8.0]])
    np.testing.assert_allclose(r[3], [[0.0, 1.0], [1.0, 0.0]])
    np.testing.assert_allclose(r[2] ** 2, x[:, None] ** 2 + x[None, :] ** 2 - x[:, None] * x[None, :])


def test_constant_correlation_gives_monopole_only():
    # xi = const -> C_l = 4 pi const delta_{l0} (Gauss-Legendre integrates P_l exactly)
    cl = ocf.corr_to_clarray(lambda r: np.full(r.shape, 2.5), 12, np.linspace(100.0, 200.0, 4), xromb=1, q=4, chunksize=10)
    np.testing.assert_allclose(cl[0], 4 * np.pi * 2.5, rtol=1e-13)
    assert np.abs(cl[1:]).max() < 1e-12
