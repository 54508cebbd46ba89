"""GPU xi(r) -> C_l front end (cora_b200.corrfunc, csrc/corrfunc.cu) against the reference-generated fixture and
the oracle, and the result fed to mkfullsky."""

import os

import numpy as np
import pytest

from oracle import corrfunc as ocf

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "corr_to_clarray.npz")


def corr_test_function(r):
    return np.exp(-((r / 60.0) ** 2)) / (1.0 + (r / 15.0) ** 2) + 0.05 * np.cos(r / 35.0) * np.exp(-r / 400.0)


@pytest.mark.parametrize("key,kw", [("cl_romb2_q2", dict(xromb=2, q=2, chunksize=16)),
                                    ("cl_romb0_q3", dict(xromb=0, q=3, chunksize=50)),
                                    ("cl_romb1_w40", dict(xromb=1, xwidth=40.0, q=2, chunksize=20))])
def test_corr_to_clarray_matches_reference_fixture(key, kw):
    from cora_b200 import corrfunc

    g = np.load(GOLD)
    got = corrfunc.corr_to_clarray(corr_test_function, int(g["lmax"]), g["xarray"], **kw)
    want = g[key]
    assert got.shape == want.shape and got.dtype == np.float64
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.abs(want).max())


def test_legendre_array_matches_reference_fixture():
    from cora_b200 import corrfunc

    g = np.load(GOLD)
    np.testing.assert_allclose(corrfunc.legendre_array(60, g["leg_mu"]), g["leg"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(corrfunc.cosine_rule(g["leg_mu"], g["xarray"], g["xarray"]),
                               ocf.cosine_rule(g["leg_mu"], g["xarray"], g["xarray"]), rtol=0, atol=0)


def test_corr_to_clarray_larger_vs_oracle_odd_sizes():
    """lmax 95, 11 radial bins (odd nx^2: the GEMM's unaligned path), ragged chunks."""
    from cora_b200 import corrfunc

    x = np.linspace(2500.0, 3600.0, 11)
    want = ocf.corr_to_clarray(corr_test_function, 95, x, xromb=2, q=2, chunksize=37)
    got = corrfunc.corr_to_clarray(corr_test_function, 95, x, xromb=2, q=2, chunksize=37)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.abs(want).max())


def test_dgemm_against_numpy():
    import torch
    from cora_b200 import _lib

    rng = np.random.default_rng(2)
    for (m, n, k) in [(129, 70, 33), (64, 64, 16), (300, 131, 257), (1, 2, 1)]:
        a, b = rng.standard_normal((m, k)), rng.standard_normal((k, n))
        ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        cd = torch.full((m, n), 7.0, dtype=torch.float64, device="cuda")
        _lib.call("cora_b200_dgemm", _lib.ptr(ad), _lib.ptr(bd), _lib.ptr(cd), m, n, k, k, n, n, 0, _lib.stream_ptr())
        np.testing.assert_allclose(cd.cpu().numpy(), a @ b, rtol=0, atol=1e-12 * k)
        _lib.call("cora_b200_dgemm", _lib.ptr(ad), _lib.ptr(bd), _lib.ptr(cd), m, n, k, k, n, n, 1, _lib.stream_ptr())
        np.testing.assert_allclose(cd.cpu().numpy(), 2 * (a @ b), rtol=0, atol=2e-12 * k)


def test_corr_to_clarray_feeds_mkfullsky():
    """The pipeline order of cora/signal/lss.py: C_l(x, x') from xi(r), then mkfullsky; M M^T reproduces it."""
    from cora_b200 import corrfunc, skysim

    nside = 8
    lmax = 3 * nside - 1
    x = np.linspace(3000.0, 3300.0, 5)
    cla = corrfunc.corr_to_clarray(corr_test_function, lmax, x, xromb=1, q=4, chunksize=10)
    sky = skysim.mkfullsky(cla, nside, seed=3)
    assert sky.shape == (5, 12 * nside**2) and np.isfinite(sky).all()
    with pytest.raises(ValueError):     # fewer nodes than one chunk: the reference dies in np.array_split(..., 0)
        corrfunc.corr_to_clarray(corr_test_function, 8, x, xromb=0, q=2, chunksize=50)


@pytest.mark.parametrize("kind", ["linear", "log"])
def test_tabulated_correlation_fused_path_vs_oracle(kind):
    """A tabulated xi(r) takes the fused kernel (cosine rule + interpolation + radial quadrature on the GPU); the
    oracle evaluates the same object as a plain host callable."""
    from cora_b200 import corrfunc

    r = np.concatenate([[1e-3], np.logspace(-1, 4, 300)])
    tab = corrfunc.TabulatedCorrelation(r, corr_test_function(r), kind=kind)
    x = np.linspace(2800.0, 3300.0, 7)
    for xromb in (0, 2):
        want = ocf.corr_to_clarray(tab, 63, x, xromb=xromb, q=2, chunksize=25)
        got = corrfunc.corr_to_clarray(tab, 63, x, xromb=xromb, q=2, chunksize=25)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.abs(want).max())
    # host call semantics = numpy.interp (clamped), including r = 0 on the log grid
    np.testing.assert_allclose(tab(np.array([0.0, 5e-4, 2e4])), [tab.values[0], tab.values[0], tab.values[-1]])

    class Sub(corrfunc.TabulatedCorrelation):      # an overriding subclass must be evaluated through the callable
        def __call__(self, rr):
            return 2.0 * corrfunc.TabulatedCorrelation.__call__(self, rr)

    got2 = corrfunc.corr_to_clarray(Sub(r, corr_test_function(r), kind=kind), 63, x, xromb=2, q=2, chunksize=25)
    np.testing.assert_allclose(got2, 2.0 * want, rtol=0, atol=1e-12 * np.abs(want).max())
