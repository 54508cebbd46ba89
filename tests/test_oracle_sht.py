"""Analytic validation of the SHT restatement (parity at this boundary is UNPINNED:
healpy is absent; see oracle/sht.py).  Checks the HEALPix RING geometry, lambda_lm and
the synthesis against scipy's spherical harmonics and closed forms."""

import numpy as np
import pytest
from scipy.special import sph_harm_y

from oracle import sht
from oracle import sht as osht


def test_ring_geometry_basic():
    for nside in (1, 2, 4, 8, 3):
        g = sht.ring_geometry(nside)
        assert g["nph"].sum() == 12 * nside * nside
        # rings tile the pixel range exactly, in order
        assert g["start"][0] == 0
        assert np.all(g["start"][1:] == np.cumsum(g["nph"])[:-1])
        # equal-area: cos(theta) spacing. z of ring centres symmetric
        np.testing.assert_allclose(g["cth"], -g["cth"][::-1], atol=1e-15)
        np.testing.assert_allclose(g["cth"] ** 2 + g["sth"] ** 2, 1.0, rtol=1e-14)
    # nside=1 known pixel centres: 3 rings of 4; z = 2/3, 0, -2/3; phi0 = pi/4, 0, pi/4
    g = sht.ring_geometry(1)
    np.testing.assert_allclose(g["cth"], [2 / 3, 0, -2 / 3], atol=1e-15)
    np.testing.assert_allclose(g["phi0"], [np.pi / 4, 0, np.pi / 4])
    # nside=2, published HEALPix values: ring 1 z = 1 - 1/12, ring 2 (belt, i=nside) z = 2/3
    g = sht.ring_geometry(2)
    np.testing.assert_allclose(g["cth"][:4], [1 - 1 / 12.0, 2 / 3.0, 1 / 3.0, 0.0], atol=1e-15)
    np.testing.assert_allclose(g["phi0"][:4], [np.pi / 4, np.pi / 8, 0.0, np.pi / 8])


def test_lambda_matches_scipy():
    theta = np.array([0.05, 0.7, 1.3, np.pi / 2, 2.4, 3.1])
    lmax = 40
    for m in (0, 1, 2, 5, 17, 40):
        lam = sht.lambda_lm(lmax, m, np.cos(theta), np.sin(theta))
        for l in range(m, lmax + 1):
            ref = sph_harm_y(l, m, theta, 0.0).real
            np.testing.assert_allclose(lam[l - m], ref, rtol=2e-12, atol=1e-13)


def test_lambda_scaled_start_no_underflow():
    # sin^m underflows double: m = 1500 at theta ~ 1e-3 is ~1e-4500; values must come back
    # in range once l ~ m / sin(theta) is reached or stay exactly 0 -- never NaN/inf.
    theta = np.array([1e-3, 0.3, np.pi / 2])
    lam = sht.lambda_lm(1600, 1500, np.cos(theta), np.sin(theta))
    assert np.all(np.isfinite(lam))
    assert np.all(lam[:, 0] == 0.0)
    # on the equator the harmonic is O(1) near the top of the l range
    assert np.abs(lam[:, 2]).max() > 0.1
    # orthonormality spot check via high-order Gauss-Legendre: int lam_lm lam_l'm dcos = delta/(2 pi)
    x, w = np.polynomial.legendre.leggauss(400)
    lam = sht.lambda_lm(120, 100, x, np.sqrt(1 - x * x))
    gram = (lam * w) @ lam.T * 2 * np.pi
    np.testing.assert_allclose(gram, np.eye(gram.shape[0]), atol=5e-12)


def test_monopole_and_single_modes():
    nside = 4
    lmax = 3 * nside - 1
    nalm = (lmax + 1) * (lmax + 2) // 2
    alm = np.zeros(nalm, dtype=complex)
    alm[0] = np.sqrt(4 * np.pi)
    np.testing.assert_allclose(sht.alm2map(alm, nside), 1.0, rtol=1e-14)
    theta, phi = sht.pix2ang_ring(nside)
    for (l, m, val) in ((1, 0, 1.0), (2, 1, 0.3 - 0.7j), (5, 5, 1.0j), (11, 3, -2.0 + 0.5j)):
        alm[:] = 0
        alm[sht.alm_index(lmax, l, m)] = val
        ref = (val.real if m == 0 else 2.0) * 1.0
        y = sph_harm_y(l, m, theta, phi)
        ref = val.real * y.real if m == 0 else 2.0 * (val * y).real
        np.testing.assert_allclose(sht.alm2map(alm, nside), ref, atol=1e-13)


def test_imag_of_m0_ignored():
    nside, lmax = 2, 5
    rng = np.random.default_rng(0)
    nalm = (lmax + 1) * (lmax + 2) // 2
    alm = rng.standard_normal(nalm) + 1j * rng.standard_normal(nalm)
    a2 = alm.copy()
    a2[: lmax + 1] = a2[: lmax + 1].real
    np.testing.assert_array_equal(sht.alm2map(alm, nside), sht.alm2map(a2, nside))


@pytest.mark.parametrize("nside,lmax", [(2, 5), (4, 11), (4, 12), (8, 23), (3, 8)])
def test_alm2map_vs_direct_sum(nside, lmax):
    # lmax = 3 nside - 1 and 3 nside: every ring has nph < 2 lmax + 1 -> aliasing path
    rng = np.random.default_rng(nside * 100 + lmax)
    nalm = (lmax + 1) * (lmax + 2) // 2
    alm = rng.standard_normal(nalm) + 1j * rng.standard_normal(nalm)
    fast = sht.alm2map(alm, nside)
    slow = sht.alm2map_direct(alm, nside)
    assert np.max(np.abs(fast - slow)) / np.max(np.abs(slow)) < 1e-13


def test_multichannel_matches_single():
    nside, lmax = 4, 11
    rng = np.random.default_rng(7)
    nalm = (lmax + 1) * (lmax + 2) // 2
    alm = rng.standard_normal((3, nalm)) + 1j * rng.standard_normal((3, nalm))
    m = sht.alm2map(alm, nside)
    for c in range(3):
        np.testing.assert_allclose(m[c], sht.alm2map(alm[c], nside), rtol=0, atol=1e-14)


def _sylm(s, l, m, theta, phi):
    """Spin-weighted harmonic _sY_lm by Goldberg et al.'s closed sum (validation only)."""
    from math import comb, factorial as f

    pref = (-1.0) ** m * np.sqrt(f(l + m) * f(l - m) * (2 * l + 1) / (4 * np.pi * f(l + s) * f(l - s)))
    sh, ch = np.sin(theta / 2), np.cos(theta / 2)
    tot = np.zeros_like(theta)
    for r in range(0, l - s + 1):
        k = r + s - m
        if k < 0 or k > l + s:
            continue
        # sin^{2l}(t/2) cot^{2r+s-m}(t/2) = cos^{2r+s-m} sin^{2l-2r-s+m}
        tot = tot + comb(l - s, r) * comb(l + s, k) * (-1.0) ** (l - r - s) * ch ** (2 * r + s - m) * sh ** (2 * l - 2 * r - s + m)
    return pref * tot * np.exp(1j * m * phi)


def test_goldberg_helper_against_known():
    theta = np.linspace(0.1, 3.0, 7)
    phi = np.linspace(0.0, 5.0, 7)
    for l, m in ((0, 0), (2, 1), (3, -2), (5, 4)):
        np.testing.assert_allclose(_sylm(0, l, m, theta, phi), sph_harm_y(l, m, theta, phi), atol=1e-13)
    np.testing.assert_allclose(_sylm(2, 2, 2, theta, phi), np.sqrt(5 / (64 * np.pi)) * (1 - np.cos(theta)) ** 2 * np.exp(2j * phi), atol=1e-14)
    np.testing.assert_allclose(_sylm(-2, 2, 2, theta, phi), np.sqrt(5 / (64 * np.pi)) * (1 + np.cos(theta)) ** 2 * np.exp(2j * phi), atol=1e-14)


@pytest.mark.parametrize("nside,lmax", [(2, 5), (4, 11), (4, 12)])
def test_spin2_vs_direct_sum(nside, lmax):
    """Q + iU = sum_{l,m=-l..l} _2a_lm _2Y_lm with _2a_lm = -(aE + i aB) (HEALPix sign)."""
    rng = np.random.default_rng(11 + lmax)
    nalm = (lmax + 1) * (lmax + 2) // 2
    aE = rng.standard_normal(nalm) + 1j * rng.standard_normal(nalm)
    aB = rng.standard_normal(nalm) + 1j * rng.standard_normal(nalm)
    for l in (0, 1):  # no spin-2 content below l = 2
        for m in range(l + 1):
            aE[sht.alm_index(lmax, l, m)] = aB[sht.alm_index(lmax, l, m)] = 0
    theta, phi = sht.pix2ang_ring(nside)
    P = np.zeros(theta.shape, dtype=complex)
    for l in range(2, lmax + 1):
        for m in range(-l, l + 1):
            i = sht.alm_index(lmax, l, abs(m))
            e, b = aE[i], aB[i]
            if m == 0:
                e, b = e.real, b.real
            if m < 0:
                e, b = (-1) ** m * np.conj(e), (-1) ** m * np.conj(b)
            P += -(e + 1j * b) * _sylm(2, l, m, theta, phi)
    Q, U = sht.alm2map_spin2(aE, aB, nside)
    scale = np.max(np.abs(P))
    assert np.max(np.abs(Q[0] - P.real)) / scale < 1e-12
    assert np.max(np.abs(U[0] - P.imag)) / scale < 1e-12


# ------------------------------------------------------------------ analysis (map2alm)
def _rand_packed(rng, lmax, nchan=1):
    nalm = (lmax + 1) * (lmax + 2) // 2
    a = rng.standard_normal((nchan, nalm)) + 1j * rng.standard_normal((nchan, nalm))
    a[:, : lmax + 1] = a[:, : lmax + 1].real   # m = 0 real
    return a


@pytest.mark.parametrize("nside,lmax", [(4, 11), (8, 23), (8, 12)])
def test_map2alm_adjoint_identity(nside, lmax):
    """The quadrature pass is (4 pi / npix) times the adjoint of the (analytically validated)
    synthesis: sum_pix f S(a) (4 pi/npix) = sum_lm w_m Re(conj(a) A(f)), w_0 = 1, w_m>0 = 2."""
    rng = np.random.default_rng(5)
    f = rng.standard_normal((2, 12 * nside**2))
    a = _rand_packed(rng, lmax, 2)
    Af = sht.map2alm_adjoint(f, nside, lmax)
    Sa = sht.alm2map(a, nside, lmax)
    lhs = (f * Sa).sum(axis=1) * 4.0 * np.pi / (12 * nside**2)
    w = np.full(a.shape[1], 2.0)
    w[: lmax + 1] = 1.0
    rhs = (w * (np.conj(a) * Af).real).sum(axis=1)
    np.testing.assert_allclose(lhs, rhs, rtol=1e-12)


def test_map2alm_band_limited_roundtrip():
    """alm -> map -> alm converges with the Jacobi iterations for a band-limited field."""
    rng = np.random.default_rng(6)
    nside, lmax = 8, 12
    a = _rand_packed(rng, lmax, 1)[0]
    m = sht.alm2map(a, nside, lmax)
    e0 = np.abs(sht.map2alm(m, nside, lmax, iter=0) - a).max()
    e3 = np.abs(sht.map2alm(m, nside, lmax, iter=3) - a).max()
    e8 = np.abs(sht.map2alm(m, nside, lmax, iter=8) - a).max()
    assert e0 < 0.1 and e3 < e0 * 1e-2 and e8 < e3


def test_map2alm_monopole_and_weights():
    nside = 4
    m = np.full(12 * nside**2, 3.0)
    a = sht.map2alm(m, nside, 5, iter=0)
    assert abs(a[0] - 3.0 * np.sqrt(4 * np.pi)) < 1e-12
    assert np.abs(a[1:]).max() < 0.3   # HEALPix quadrature leakage at nside 4 (percent level)
    # uniform ring weights c scale the quadrature pass by c
    a2 = sht.map2alm(m, nside, 5, iter=0, ring_weights=np.full(2 * nside, 1.5))
    np.testing.assert_allclose(a2, 1.5 * a, atol=1e-13)
    with pytest.raises(ValueError):
        sht.map2alm(m, nside, 5, ring_weights=np.ones(3))


def test_anafast_recovers_spectrum():
    rng = np.random.default_rng(7)
    nside, lmax = 16, 32
    cl = 1.0 / (1.0 + np.arange(lmax + 1)) ** 2
    a = _rand_packed(rng, lmax, 1)[0]
    for mm in range(lmax + 1):
        sl = slice(sht.alm_index(lmax, mm, mm), sht.alm_index(lmax, lmax, mm) + 1)
        a[sl] *= np.sqrt(cl[mm:] / (1.0 if mm == 0 else 2.0))
    est = sht.anafast(sht.alm2map(a, nside, lmax), lmax=lmax, iter=3)
    l = np.arange(2, lmax + 1)
    assert np.all(np.abs(est[l] - cl[l]) < 6 * cl[l] * np.sqrt(2.0 / (2 * l + 1)))


@pytest.mark.parametrize("nside,lmax", [(4, 11), (8, 17)])
def test_map2alm_spin2_adjoint_identity(nside, lmax):
    """<(Q,U), S2(aE,aB)> 4 pi/npix = sum_lm w_m Re(conj(aE) AE + conj(aB) AB): the polarised
    quadrature pass is the adjoint of the (closed-form validated) spin-2 synthesis."""
    rng = np.random.default_rng(8)
    npix = 12 * nside**2
    Q, U = rng.standard_normal((2, 2, npix))
    aE, aB = _rand_packed(rng, lmax, 2), _rand_packed(rng, lmax, 2)
    for a in (aE, aB):   # l < 2 carries no spin-2 power
        for mm in range(2):
            for ll in range(mm, 2):
                a[:, sht.alm_index(lmax, ll, mm)] = 0
    AE, AB = sht.map2alm_spin2_adjoint(Q, U, nside, lmax)
    q, u = sht.alm2map_spin2(aE, aB, nside, lmax)
    lhs = ((Q * q).sum(axis=1) + (U * u).sum(axis=1)) * 4.0 * np.pi / npix
    w = np.full(aE.shape[1], 2.0)
    w[: lmax + 1] = 1.0
    rhs = (w * ((np.conj(aE) * AE).real + (np.conj(aB) * AB).real)).sum(axis=1)
    np.testing.assert_allclose(lhs, rhs, rtol=1e-11)
    # l = 0, 1 rows of the analysis are identically zero
    for mm in range(2):
        for ll in range(mm, 2):
            assert np.all(AE[:, sht.alm_index(lmax, ll, mm)] == 0) and np.all(AB[:, sht.alm_index(lmax, ll, mm)] == 0)


def test_map2alm_spin2_roundtrip():
    rng = np.random.default_rng(9)
    nside, lmax = 8, 12
    aE, aB = _rand_packed(rng, lmax, 1), _rand_packed(rng, lmax, 1)
    for a in (aE, aB):
        for mm in range(2):
            for ll in range(mm, 2):
                a[:, sht.alm_index(lmax, ll, mm)] = 0
    q, u = sht.alm2map_spin2(aE, aB, nside, lmax)
    e0 = max(np.abs(x - y).max() for x, y in zip(sht.map2alm_spin2(q, u, nside, lmax, iter=0), (aE, aB)))
    e4 = max(np.abs(x - y).max() for x, y in zip(sht.map2alm_spin2(q, u, nside, lmax, iter=4), (aE, aB)))
    assert e0 < 0.2 and e4 < e0 * 1e-2


def test_ang_positions_and_nside_for_lmax_host_helpers():
    """cora_b200.hputil helpers that need no device: pixel angles equal the oracle geometry."""
    from cora_b200 import hputil as bh

    for nside in (1, 2, 8):
        theta, phi = sht.pix2ang_ring(nside)
        ang = bh.ang_positions(nside)
        assert ang.shape == (12 * nside**2, 2)
        np.testing.assert_allclose(ang[:, 0], theta, atol=1e-14)
        np.testing.assert_allclose(ang[:, 1], phi, atol=1e-14)
    assert bh.nside_for_lmax(767) == 512 and bh.nside_for_lmax(95, accuracy_boost=0) == 32


# ---------------------------------------------------------------- ring-subset routines (production-size parity helpers)
def test_lambda_sweep_matches_lambda_lm():
    nside, lmax = 16, 47
    g = osht.ring_geometry(nside)
    rs = osht.parity_rings(nside)
    ref = {m: osht.lambda_lm(lmax, m, g["cth"][rs], g["sth"][rs]) for m in range(lmax + 1)}
    for l, lam, prev in osht.lambda_sweep(lmax, g["cth"][rs], g["sth"][rs]):
        want = np.array([ref[m][l - m] for m in range(l + 1)])
        np.testing.assert_allclose(lam, want, rtol=1e-13, atol=1e-14)   # lambda = O(1); near its zeros only the absolute error is meaningful
        if l:
            wprev = np.array([ref[m][l - 1 - m] if m < l else np.zeros(len(rs)) for m in range(l + 1)])
            np.testing.assert_allclose(prev, wprev, rtol=1e-13, atol=1e-14)


def test_ring_subset_transforms_match_full():
    nside, lmax = 16, 47
    rng = np.random.default_rng(1)
    nalm = (lmax + 1) * (lmax + 2) // 2
    aE = rng.standard_normal((3, nalm)) + 1j * rng.standard_normal((3, nalm))
    aB = rng.standard_normal((3, nalm)) + 1j * rng.standard_normal((3, nalm))
    rs = osht.parity_rings(nside)
    full = osht.alm2map(aE, nside, lmax)
    vals, start = osht.alm2map_rings(aE, nside, lmax, rs)
    for v, s in zip(vals, start):
        np.testing.assert_allclose(v, full[:, s : s + v.shape[1]], atol=1e-12 * np.abs(full).max())
    Q, U = osht.alm2map_spin2(aE, aB, nside, lmax)
    vq, vu, start = osht.alm2map_spin2_rings(aE, aB, nside, lmax, rs)
    for q, u, s in zip(vq, vu, start):
        np.testing.assert_allclose(q, Q[:, s : s + q.shape[1]], atol=1e-12 * np.abs(Q).max())
        np.testing.assert_allclose(u, U[:, s : s + u.shape[1]], atol=1e-12 * np.abs(U).max())
    ad = osht.map2alm_adjoint(full, nside, lmax)
    for m, v in osht.map2alm_adjoint_ms(full, nside, lmax, [0, 5, 47]).items():
        np.testing.assert_array_equal(v, ad[:, osht.alm_index(lmax, m, m) : osht.alm_index(lmax, lmax, m) + 1])


# ------------------------------------------------------------------ exact spin-2 known answers up to l = 64
def _sylm_exact(s, l, m, half):
    """_sY_lm(theta, 0) from Goldberg et al.'s closed sum in EXACT rational arithmetic.  ``half = (p, q, n)`` encodes the
    half angle as cos(theta/2) = p / sqrt(n), sin(theta/2) = q / sqrt(n) (p^2 + q^2 = n): every term of the sum is then
    the integer C(l-s, r) C(l+s, r+s-m) (-1)^(l-r-s) p^a q^b over n^l, so the alternating sum -- which loses all its
    digits in floating point beyond l ~ 30 -- is exact; only the final square-root prefactor is rounded."""
    from fractions import Fraction
    from math import comb, factorial as f

    p, q, n = half
    tot = 0
    for r in range(0, l - s + 1):
        k = r + s - m
        if k < 0 or k > l + s:
            continue
        a, b = 2 * r + s - m, 2 * l - 2 * r - s + m
        tot += comb(l - s, r) * comb(l + s, k) * (-1) ** (l - r - s) * p**a * q**b
    pref2 = Fraction(f(l + m) * f(l - m) * (2 * l + 1), f(l + s) * f(l - s))      # times 1 / (4 pi)
    val = Fraction(tot, n**l)
    return (-1.0) ** m * float(val) * np.sqrt(float(pref2) / (4 * np.pi)) if abs(val) < 1e300 else None


@pytest.mark.parametrize("half", [(2, 1, 5), (2, 3, 13), (4, 3, 25), (1, 7, 50)])
def test_spin2_functions_exact_known_answers_to_l64(half):
    """X1 + X2 = _{+2}Y_lm(theta, 0) and X1 - X2 = _{-2}Y_lm(theta, 0) for every (l, m) up to l = 64,
    against exact rational evaluations of the closed form -- independent of any recurrence."""
    p, q, n = half
    c = np.array([(p * p - q * q) / n])            # cos(theta) = cos^2 - sin^2 of the half angle
    s = np.array([2.0 * p * q / n])
    lmax = 64
    worst = 0.0
    for m in range(0, lmax + 1):
        X1, X2 = sht._spin2_X(lmax, m, c, s)
        for l in range(max(m, 2), lmax + 1):
            yp = _sylm_exact(2, l, m, half)
            ym = _sylm_exact(-2, l, m, half)
            scale = max(abs(yp), abs(ym), 1e-3)
            worst = max(worst, abs(X1[l - m, 0] + X2[l - m, 0] - yp) / scale, abs(X1[l - m, 0] - X2[l - m, 0] - ym) / scale)
    assert worst < 1e-11, worst


def test_scalar_lambda_exact_known_answers_to_l64():
    """lambda_lm(theta) = _0Y_lm(theta, 0) against the same exact closed form (l up to 64, all m)."""
    half = (2, 3, 13)
    p, q, n = half
    c = np.array([(p * p - q * q) / n])
    s = np.array([2.0 * p * q / n])
    worst = 0.0
    for m in range(0, 65):
        lam = sht.lambda_lm(64, m, c, s)
        for l in range(m, 65):
            y = _sylm_exact(0, l, m, half)
            worst = max(worst, abs(lam[l - m, 0] - y) / max(abs(y), 1e-3))
    assert worst < 1e-11, worst
