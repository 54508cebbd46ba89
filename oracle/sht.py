"""Oracle restatement of ``healpy.alm2map`` on the HEALPix RING grid (scalar and spin-2).

TEST INFRASTRUCTURE ONLY (see package doc).  **Parity unpinned**: healpy (>=1.17,
``pyproject.toml:30``; libsharp / HEALPix C++ inside) is third-party, absent from
/root/reference and not installable here, and the reference's tests hold no golden at
this boundary.  This file restates the *published* HEALPix definitions (Gorski et al.
2005, the HEALPix primer, and the ``alm2map``/``alm2map_pol`` facility documentation)
that the reference's call sites rely on:

* ``cora/util/hputil.py:388``      scalar ``healpy.alm2map(alm_packed, nside)``
* ``cora/util/hputil.py:420-423``  ``healpy.alm2map([T, E, B], nside)`` (``pol=True`` default)
* ``cora/util/hputil.py:426-430``  V through a separate scalar transform

Conventions: alm packed m-major, ``idx(l, m) = m (2 lmax + 1 - m) / 2 + l``
(``hputil.py:124-152``); ``lmax`` inferred from the packed length, ``mmax = lmax``; no beam
or pixel window; the imaginary part of ``a_l0`` is ignored; the map value is the direct sum

    T(ring, j) = sum_m w_m Re[ F_m(theta_ring) exp(i m phi_j) ],   w_0 = 1, w_{m>0} = 2,
    F_m(theta) = sum_{l>=m} a_lm lambda_lm(cos theta)

so harmonics with ``m >= nph`` alias onto ``m mod nph`` exactly as the sum dictates.

It is validated analytically in ``tests/test_oracle_sht.py`` (scipy ``sph_harm_y`` direct
sums, ``a_00`` -> constant map, spin-2 against the closed forms of ``_{+-2}Y_lm``).
"""

import numpy as np


# ------------------------------------------------------------------------ geometry
def nside2npix(nside):
    return 12 * nside * nside


def ring_geometry(nside):
    """Per-ring arrays for rings 1 .. 4 nside - 1 (north to south).

    Returns dict with ``nph`` (pixels in ring), ``cth``/``sth`` (cos/sin of colatitude),
    ``phi0`` (azimuth of first pixel) and ``start`` (index of first pixel, RING order).
    """
    n = int(nside)
    i = np.arange(1, 4 * n, dtype=np.int64)
    npix = 12 * n * n
    north = np.minimum(i, 4 * n - i)  # mirror ring index for the south
    cap = north < n
    nph = np.where(cap, 4 * north, 4 * n)
    nf = float(n)
    nn = north.astype(np.float64)
    # polar cap: cos = 1 - i^2 / (3 n^2); belt: cos = 4/3 - 2 i / (3 n)
    cth_cap = 1.0 - nn * nn / (3.0 * nf * nf)
    cth_belt = 4.0 / 3.0 - 2.0 * nn / (3.0 * nf)
    cth_n = np.where(cap, cth_cap, cth_belt)
    # sin(theta) without cancellation near the poles
    sth_cap = np.sqrt((nn * nn / (3.0 * nf * nf)) * (1.0 + cth_cap))
    sth_belt = np.sqrt(np.maximum((1.0 - cth_belt) * (1.0 + cth_belt), 0.0))
    sth = np.where(cap, sth_cap, sth_belt)
    cth = np.where(i > 2 * n, -cth_n, cth_n)
    shifted = cap | (((north - n) & 1) == 0)
    phi0 = np.where(shifted, np.pi / nph, 0.0)
    start_n = np.where(cap, 2 * north * (north - 1), 2 * n * (n - 1) + 4 * n * (north - n))
    start_s = np.where(cap, npix - 2 * north * (north + 1), 0)
    # southern belt rings: continue the belt numbering
    start_belt_s = 2 * n * (n - 1) + 4 * n * (i - n)
    start = np.where(i <= 2 * n, start_n, np.where(cap, start_s, start_belt_s))
    return {"nph": nph, "cth": cth, "sth": sth, "phi0": phi0, "start": start}


def pix2ang_ring(nside):
    """(theta, phi) of every pixel in RING order."""
    g = ring_geometry(nside)
    theta = np.empty(nside2npix(nside))
    phi = np.empty_like(theta)
    for r in range(len(g["nph"])):
        s, n = g["start"][r], g["nph"][r]
        theta[s : s + n] = np.arctan2(g["sth"][r], g["cth"][r])
        phi[s : s + n] = g["phi0"][r] + 2.0 * np.pi * np.arange(n) / n
    return theta, phi


# ------------------------------------------------------------------- alm utilities
def lmax_from_nalm(nalm):
    lmax = int((np.sqrt(8.0 * nalm + 1.0) - 3.0) / 2.0 + 0.5)
    if (lmax + 1) * (lmax + 2) // 2 != nalm:
        raise ValueError("packed alm length %d is not (lmax+1)(lmax+2)/2" % nalm)
    return lmax


def alm_index(lmax, l, m):
    return m * (2 * lmax + 1 - m) // 2 + l


# ------------------------------------------------------------------- lambda_lm
def lambda_lm(lmax, m, cth, sth):
    """``lambda_lm(theta)`` for ``l = m .. lmax`` at every ring, shape (lmax-m+1, nring).

    ``lambda_lm = sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) P_l^m(cos theta)`` with the
    Condon-Shortley phase.  Seed ``lambda_mm`` in log space and run the standard
    three-term recurrence on (mantissa, binary exponent) pairs so that ``sin^m theta``
    may underflow double without harming the in-range values; the result is flushed to
    0 only where it is below the smallest normal double.
    """
    cth = np.asarray(cth, dtype=np.float64)
    sth = np.asarray(sth, dtype=np.float64)
    nr = cth.shape[0]
    out = np.zeros((lmax - m + 1, nr))
    # log|lambda_mm| = 0.5 [ log((2m+1)/(4 pi)) + sum_k log((2k-1)/(2k)) ] + m log sin
    k = np.arange(1, m + 1, dtype=np.float64)
    logn = 0.5 * (np.log((2 * m + 1) / (4.0 * np.pi)) + np.sum(np.log1p(-1.0 / (2.0 * k))))
    with np.errstate(divide="ignore"):
        loglam = logn + m * np.log(sth)
    # mantissa / exponent split (base 2)
    e = np.floor(loglam / np.log(2.0))
    e = np.where(np.isfinite(e), e, -1e9)
    mant = np.exp(loglam - e * np.log(2.0)) * (-1.0 if (m & 1) else 1.0)
    mant = np.where(sth > 0, mant, 1.0 if m == 0 else 0.0)
    if m == 0:
        mant = np.full(nr, np.sqrt(1.0 / (4.0 * np.pi)))
        e = np.zeros(nr)
    scale = e.astype(np.int64)
    p_prev = np.zeros(nr)
    p_cur = mant
    out[0] = np.ldexp(p_cur, np.clip(scale, -100000, 100000).astype(np.int32))
    a_prev = 0.0
    for l in range(m + 1, lmax + 1):
        a_l = np.sqrt((4.0 * l * l - 1.0) / (l * l - m * m))
        if l == m + 1:
            p_new = a_l * cth * p_cur
        else:
            p_new = a_l * (cth * p_cur - p_prev / a_prev)
        p_prev, p_cur, a_prev = p_cur, p_new, a_l
        big = np.abs(p_cur) > 2.0**64
        if big.any():
            p_cur = np.where(big, p_cur * 2.0**-64, p_cur)
            p_prev = np.where(big, p_prev * 2.0**-64, p_prev)
            scale = np.where(big, scale + 64, scale)
        out[l - m] = np.ldexp(p_cur, np.clip(scale, -100000, 100000).astype(np.int32))
    return out


# ------------------------------------------------------------------- ring synthesis
def _rings_from_phase(Fm, geom, ring_sel, out):
    """Phase synthesis.  ``Fm``: complex (mmax+1, nring_sel, nchan) -> ``out[chan, pix]``.

    Direct-sum definition implemented through the aliased spectrum
    ``G_k = sum_{m = k mod nph} w_m F_m e^{i m phi0}`` and ``T_j = Re sum_k G_k e^{2 pi i jk/nph}``.
    """
    mmax = Fm.shape[0] - 1
    m = np.arange(mmax + 1)
    w = np.where(m == 0, 1.0, 2.0)
    for a, r in enumerate(ring_sel):
        nph = int(geom["nph"][r])
        start = int(geom["start"][r])
        ph = w * np.exp(1j * m * geom["phi0"][r])
        G = np.zeros((nph, Fm.shape[2]), dtype=np.complex128)
        np.add.at(G, m % nph, Fm[:, a, :] * ph[:, None])
        out[:, start : start + nph] = (np.fft.ifft(G, axis=0).real * nph).T


def alm2map(alm, nside, lmax=None):
    """Scalar synthesis.  ``alm``: complex (nalm,) or (nchan, nalm) healpy-packed."""
    alm = np.asarray(alm, dtype=np.complex128)
    single = alm.ndim == 1
    alm = np.atleast_2d(alm)
    nchan, nalm = alm.shape
    if lmax is None:
        lmax = lmax_from_nalm(nalm)
    g = ring_geometry(nside)
    nring = 4 * nside - 1
    Fm = np.zeros((lmax + 1, nring, nchan), dtype=np.complex128)
    for m in range(lmax + 1):
        lam = lambda_lm(lmax, m, g["cth"], g["sth"])  # (nl, nring)
        a = alm[:, alm_index(lmax, m, m) : alm_index(lmax, lmax, m) + 1].copy()  # (nchan, nl)
        if m == 0:
            a = a.real.astype(np.complex128)
        Fm[m] = lam.T @ a.T
    out = np.empty((nchan, nside2npix(nside)))
    _rings_from_phase(Fm, g, np.arange(nring), out)
    return out[0] if single else out


# ------------------------------------------------------------------------ spin 2
def alm2map_spin2(almE, almB, nside, lmax=None):
    """(E, B) -> (Q, U) in the HEALPix convention.

    ``Q = -sum (aE X1 + i aB X2)``, ``U = -sum (aB X1 - i aE X2)`` with
    ``X1 = 2 n_l [ -((l - m^2)/s^2 + l(l-1)/2) lam_lm + (c/s^2) g_lm lam_{l-1,m} ]``,
    ``X2 = 2 n_l (m/s^2) [ -(l-1) c lam_lm + g_lm lam_{l-1,m} ]``,
    ``n_l = ((l+2)(l+1)l(l-1))^-1/2``, ``g_lm = sqrt((2l+1)/(2l-1) (l^2-m^2))``; zero for l < 2.
    """
    almE = np.atleast_2d(np.asarray(almE, dtype=np.complex128))
    almB = np.atleast_2d(np.asarray(almB, dtype=np.complex128))
    nchan, nalm = almE.shape
    if lmax is None:
        lmax = lmax_from_nalm(nalm)
    g = ring_geometry(nside)
    c, s = g["cth"], g["sth"]
    s2 = s * s
    nring = 4 * nside - 1
    FQ = np.zeros((lmax + 1, nring, nchan), dtype=np.complex128)
    FU = np.zeros_like(FQ)
    for m in range(lmax + 1):
        lam = lambda_lm(lmax, m, c, s)  # l = m..lmax
        l = np.arange(m, lmax + 1, dtype=np.float64)
        lam_prev = np.vstack([np.zeros((1, nring)), lam[:-1]])
        with np.errstate(divide="ignore", invalid="ignore"):
            nl = np.where(l >= 2, 1.0 / np.sqrt((l + 2) * (l + 1) * l * (l - 1)), 0.0)
        glm = np.sqrt((2 * l + 1) / np.maximum(2 * l - 1, 1.0) * (l * l - m * m))
        X1 = (2 * nl)[:, None] * (
            -((l - m * m)[:, None] / s2[None, :] + (l * (l - 1) / 2)[:, None]) * lam
            + (c / s2)[None, :] * glm[:, None] * lam_prev
        )
        X2 = (2 * nl)[:, None] * (m / s2)[None, :] * (
            -(l - 1)[:, None] * c[None, :] * lam + glm[:, None] * lam_prev
        )
        sl = slice(alm_index(lmax, m, m), alm_index(lmax, lmax, m) + 1)
        aE, aB = almE[:, sl].copy(), almB[:, sl].copy()
        if m == 0:
            aE = aE.real.astype(np.complex128)
            aB = aB.real.astype(np.complex128)
        FQ[m] = -(X1.T @ aE.T + 1j * (X2.T @ aB.T))
        FU[m] = -(X1.T @ aB.T - 1j * (X2.T @ aE.T))
    Q = np.empty((nchan, nside2npix(nside)))
    U = np.empty_like(Q)
    rs = np.arange(nring)
    _rings_from_phase(FQ, g, rs, Q)
    _rings_from_phase(FU, g, rs, U)
    return Q, U


def alm2map_direct(alm, nside, lmax=None):
    """O(npix * nalm) brute-force evaluation with scipy's Y_lm — validation only."""
    from scipy.special import sph_harm_y

    alm = np.asarray(alm, dtype=np.complex128)
    if lmax is None:
        lmax = lmax_from_nalm(alm.shape[-1])
    theta, phi = pix2ang_ring(nside)
    out = np.zeros(theta.shape)
    for m in range(lmax + 1):
        for l in range(m, lmax + 1):
            a = alm[alm_index(lmax, l, m)]
            if m == 0:
                out += a.real * sph_harm_y(l, 0, theta, phi).real
            else:
                out += 2.0 * (a * sph_harm_y(l, m, theta, phi)).real
    return out


# ------------------------------------------------------------------- analysis (map -> alm)
# Restates ``healpy.map2alm(map, lmax=lmax, use_weights=..., iter=...)`` as used by
# ``cora/util/hputil.py:195-234`` (``sphtrans_real``), ``:274-323`` (``sphtrans_real_pol``),
# ``:460-497`` (``sphtrans_sky``) and ``:607-619`` (``sph_ps``): the HEALPix quadrature
#
#     a_lm^(0) = (4 pi / npix) sum_rings w_ring lambda_lm(theta_ring) sum_j f(ring, j) exp(-i m phi_j)
#
# followed by ``iter`` Jacobi refinements  a += A(f - S a)  (HEALPix ``map2alm_iterative``).
# cora calls it with ``use_weights=True, iter=2`` (``hputil.py:46-47``); healpy then reads the ring
# weights from its data files (``weight_ring_n%05d.fits``), which are not available here, so the
# weights are an explicit argument (absolute, one per northern ring incl. the equator; default 1).
# Parity unpinned for the same reason as the synthesis (no healpy); pinned analytically by the
# adjoint identity against the validated synthesis (tests/test_oracle_sht.py).
def _phase_from_rings(maps, geom, ring_sel, mmax, wgt):
    """Ring analysis: ``F_m(ring) = wgt[ring] sum_j f(ring, j) exp(-i m phi_j)`` -> (mmax+1, nsel, nchan)."""
    m = np.arange(mmax + 1)
    out = np.zeros((mmax + 1, len(ring_sel), maps.shape[0]), dtype=np.complex128)
    for a, r in enumerate(ring_sel):
        nph = int(geom["nph"][r])
        start = int(geom["start"][r])
        G = np.fft.fft(maps[:, start : start + nph], axis=1)  # (nchan, nph)
        ph = wgt[r] * np.exp(-1j * m * geom["phi0"][r])
        out[:, a, :] = G[:, m % nph].T * ph[:, None]
    return out


def _full_ring_weights(nside, ring_weights):
    nring = 4 * nside - 1
    w = np.ones(nring)
    if ring_weights is not None:
        rw = np.asarray(ring_weights, dtype=np.float64)
        if rw.shape != (2 * nside,):
            raise ValueError("ring_weights must have 2*nside entries")
        i = np.arange(1, 4 * nside)
        w = rw[np.minimum(i, 4 * nside - i) - 1]
    return w * (4.0 * np.pi / nside2npix(nside))


def map2alm_adjoint(maps, nside, lmax, ring_weights=None):
    """One quadrature pass (no iteration): packed alm (nchan, nalm)."""
    maps = np.atleast_2d(np.asarray(maps, dtype=np.float64))
    g = ring_geometry(nside)
    nring = 4 * nside - 1
    F = _phase_from_rings(maps, g, np.arange(nring), lmax, _full_ring_weights(nside, ring_weights))
    nalm = (lmax + 1) * (lmax + 2) // 2
    alm = np.zeros((maps.shape[0], nalm), dtype=np.complex128)
    for m in range(lmax + 1):
        lam = lambda_lm(lmax, m, g["cth"], g["sth"])  # (nl, nring)
        alm[:, alm_index(lmax, m, m) : alm_index(lmax, lmax, m) + 1] = (lam @ F[m]).T
    return alm


def map2alm(maps, nside=None, lmax=None, iter=3, ring_weights=None):
    """``healpy.map2alm`` restatement (scalar): quadrature + ``iter`` Jacobi refinements."""
    maps = np.asarray(maps, dtype=np.float64)
    single = maps.ndim == 1
    maps = np.atleast_2d(maps)
    if nside is None:
        nside = int(round(np.sqrt(maps.shape[1] / 12.0)))
    if lmax is None:
        lmax = 3 * nside - 1
    alm = map2alm_adjoint(maps, nside, lmax, ring_weights)
    for _ in range(iter):
        alm += map2alm_adjoint(maps - alm2map(alm, nside, lmax), nside, lmax, ring_weights)
    return alm[0] if single else alm


def anafast(map1, map2=None, lmax=None, iter=3, ring_weights=None):
    """Cross/auto power spectrum ``C_l = (|a_l0|^2 + 2 sum_{m>0} Re a1 conj(a2)) / (2l+1)``
    (``hputil.sph_ps``, ``hputil.py:607-619`` / ``healpy.anafast``)."""
    map1 = np.asarray(map1, dtype=np.float64)
    nside = int(round(np.sqrt(map1.shape[-1] / 12.0)))
    if lmax is None:
        lmax = 3 * nside - 1
    a1 = map2alm(map1, nside, lmax, iter, ring_weights)
    a2 = a1 if map2 is None else map2alm(map2, nside, lmax, iter, ring_weights)
    cl = np.zeros(lmax + 1)
    for m in range(lmax + 1):
        sl = slice(alm_index(lmax, m, m), alm_index(lmax, lmax, m) + 1)
        p = (a1[..., sl] * np.conj(a2[..., sl])).real
        cl[m:] += (1.0 if m == 0 else 2.0) * p
    return cl / (2.0 * np.arange(lmax + 1) + 1.0)


# ------------------------------------------------------------------- spin-2 analysis
def _spin2_X(lmax, m, c, s):
    """X1, X2 of SURVEY App. A.9 for l = m..lmax at every ring, shapes (nl, nring)."""
    nring = c.shape[0]
    s2 = s * s
    lam = lambda_lm(lmax, m, c, s)
    l = np.arange(m, lmax + 1, dtype=np.float64)
    lam_prev = np.vstack([np.zeros((1, nring)), lam[:-1]])
    with np.errstate(divide="ignore", invalid="ignore"):
        nl = np.where(l >= 2, 1.0 / np.sqrt((l + 2) * (l + 1) * l * (l - 1)), 0.0)
    glm = np.sqrt((2 * l + 1) / np.maximum(2 * l - 1, 1.0) * (l * l - m * m))
    X1 = (2 * nl)[:, None] * (
        -((l - m * m)[:, None] / s2[None, :] + (l * (l - 1) / 2)[:, None]) * lam + (c / s2)[None, :] * glm[:, None] * lam_prev
    )
    X2 = (2 * nl)[:, None] * (m / s2)[None, :] * (-(l - 1)[:, None] * c[None, :] * lam + glm[:, None] * lam_prev)
    return X1, X2


def map2alm_spin2_adjoint(Q, U, nside, lmax, ring_weights=None):
    """One quadrature pass of the polarised analysis, the Hermitian adjoint of ``alm2map_spin2``
    times 4 pi / npix:  ``aE = -sum_r w (X1 Q_m + i X2 U_m)``, ``aB = -sum_r w (X1 U_m - i X2 Q_m)``
    (``healpy.map2alm([T, Q, U])`` at ``cora/util/hputil.py:310-312``, E and B parts)."""
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
    U = np.atleast_2d(np.asarray(U, dtype=np.float64))
    g = ring_geometry(nside)
    nring = 4 * nside - 1
    w = _full_ring_weights(nside, ring_weights)
    rs = np.arange(nring)
    FQ = _phase_from_rings(Q, g, rs, lmax, w)
    FU = _phase_from_rings(U, g, rs, lmax, w)
    nalm = (lmax + 1) * (lmax + 2) // 2
    aE = np.zeros((Q.shape[0], nalm), dtype=np.complex128)
    aB = np.zeros_like(aE)
    for m in range(lmax + 1):
        X1, X2 = _spin2_X(lmax, m, g["cth"], g["sth"])
        sl = slice(alm_index(lmax, m, m), alm_index(lmax, lmax, m) + 1)
        aE[:, sl] = -(X1 @ FQ[m] + 1j * (X2 @ FU[m])).T
        aB[:, sl] = -(X1 @ FU[m] - 1j * (X2 @ FQ[m])).T
    return aE, aB


def map2alm_spin2(Q, U, nside, lmax=None, iter=3, ring_weights=None):
    """(Q, U) -> (aE, aB): quadrature + ``iter`` Jacobi refinements, as ``healpy.map2alm`` does."""
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
    U = np.atleast_2d(np.asarray(U, dtype=np.float64))
    if lmax is None:
        lmax = 3 * nside - 1
    aE, aB = map2alm_spin2_adjoint(Q, U, nside, lmax, ring_weights)
    for _ in range(iter):
        q, u = alm2map_spin2(aE, aB, nside, lmax)
        dE, dB = map2alm_spin2_adjoint(Q - q, U - u, nside, lmax, ring_weights)
        aE += dE
        aB += dB
    return aE, aB


# ------------------------------------------------------------------- ring subsets at production sizes
# The per-m routines above loop over l in Python once per m: O(lmax^2) interpreter iterations, minutes at
# lmax = 3071.  For parity checks of the CUDA transforms at nside 256 / 512 / 1024 the same arithmetic is
# swept l-major over all m at once for a *subset of rings* (poles, cap/belt boundary, equator): lmax + 1
# interpreter iterations.  ``lambda_sweep`` performs the operations of ``lambda_lm`` (same seed, same recurrence,
# same rescaling); the two agree to a few ulp (the seed's log-sum is accumulated in a different order), which
# tests/test_oracle_sht.py checks together with the ring-subset transforms against the full ones.
def lambda_sweep(lmax, cth, sth):
    """Generator over l = 0..lmax yielding ``(l, lam, lam_prev)`` with ``lam[m, ring] = lambda_lm`` and
    ``lam_prev[m, ring] = lambda_{l-1,m}`` (0 for m >= l) for m = 0..l, at the given rings."""
    cth = np.asarray(cth, dtype=np.float64)
    sth = np.asarray(sth, dtype=np.float64)
    nr = cth.shape[0]
    ms = np.arange(lmax + 1)
    k = np.arange(1, lmax + 1, dtype=np.float64)
    csum = np.concatenate([[0.0], np.cumsum(np.log1p(-1.0 / (2.0 * k)))])
    logn = 0.5 * (np.log((2 * ms + 1) / (4.0 * np.pi)) + csum)
    p_cur = np.zeros((lmax + 1, nr))
    p_prev = np.zeros((lmax + 1, nr))
    scale = np.zeros((lmax + 1, nr), dtype=np.int64)
    a_prev = np.ones(lmax + 1)
    lam_last = np.zeros((0, nr))
    with np.errstate(divide="ignore"):
        logs = np.log(sth)
    for l in range(lmax + 1):
        if l > 0:
            m = np.arange(l, dtype=np.float64)
            a_l = np.sqrt((4.0 * l * l - 1.0) / (l * l - m * m))
            pc, pp = p_cur[:l], p_prev[:l]
            p_new = a_l[:, None] * (cth[None, :] * pc - pp / a_prev[:l, None])
            # row m = l - 1 is the first step of its recurrence: a_l * cth * p_cur exactly (p_prev = 0)
            p_new[l - 1] = a_l[l - 1] * cth * pc[l - 1]
            p_prev[:l] = pc
            p_cur[:l] = p_new
            a_prev[:l] = a_l
            big = np.abs(p_cur[:l]) > 2.0**64
            if big.any():
                p_cur[:l] = np.where(big, p_cur[:l] * 2.0**-64, p_cur[:l])
                p_prev[:l] = np.where(big, p_prev[:l] * 2.0**-64, p_prev[:l])
                scale[:l] = np.where(big, scale[:l] + 64, scale[:l])
        # seed of row m = l (lambda_mm), exactly as lambda_lm does it
        if l == 0:
            p_cur[0] = np.sqrt(1.0 / (4.0 * np.pi))
            scale[0] = 0
        else:
            loglam = logn[l] + l * logs
            e = np.floor(loglam / np.log(2.0))
            e = np.where(np.isfinite(e), e, -1e9)
            mant = np.exp(loglam - e * np.log(2.0)) * (-1.0 if (l & 1) else 1.0)
            mant = np.where(sth > 0, mant, 0.0)
            p_cur[l] = mant
            scale[l] = e.astype(np.int64)
        lam = np.ldexp(p_cur[: l + 1], np.clip(scale[: l + 1], -100000, 100000).astype(np.int32))
        prev = np.zeros_like(lam)
        prev[: lam_last.shape[0]] = lam_last
        yield l, lam, prev
        lam_last = lam


def _ring_values(Fm, geom, ring_sel):
    """Phase synthesis of selected rings -> list of arrays (nchan, nph_ring)."""
    mmax = Fm.shape[0] - 1
    m = np.arange(mmax + 1)
    w = np.where(m == 0, 1.0, 2.0)
    vals = []
    for a, r in enumerate(ring_sel):
        nph = int(geom["nph"][r])
        ph = w * np.exp(1j * m * geom["phi0"][r])
        G = np.zeros((nph, Fm.shape[2]), dtype=np.complex128)
        np.add.at(G, m % nph, Fm[:, a, :] * ph[:, None])
        vals.append((np.fft.ifft(G, axis=0).real * nph).T)
    return vals


def alm2map_rings(alm, nside, lmax, ring_sel):
    """Scalar synthesis (same definition as ``alm2map``) evaluated only on the rings ``ring_sel`` (0-based
    ring numbers, north to south).  Returns ``(vals, start)``: ``vals[k]`` is ``(nchan, nph)`` for ring
    ``ring_sel[k]``, ``start[k]`` its first RING pixel."""
    alm = np.atleast_2d(np.asarray(alm, dtype=np.complex128))
    nchan = alm.shape[0]
    g = ring_geometry(nside)
    rs = np.asarray(ring_sel, dtype=np.int64)
    Fm = np.zeros((lmax + 1, len(rs), nchan), dtype=np.complex128)
    mall = np.arange(lmax + 1)
    for l, lam, _ in lambda_sweep(lmax, g["cth"][rs], g["sth"][rs]):
        a = alm[:, alm_index(lmax, l, mall[: l + 1])].T.copy()     # (l+1, nchan)
        a[0] = a[0].real
        Fm[: l + 1] += lam[:, :, None] * a[:, None, :]
    return _ring_values(Fm, g, rs), g["start"][rs]


def alm2map_spin2_rings(almE, almB, nside, lmax, ring_sel):
    """(E, B) -> (Q, U) (same definition as ``alm2map_spin2``) on the rings ``ring_sel``."""
    almE = np.atleast_2d(np.asarray(almE, dtype=np.complex128))
    almB = np.atleast_2d(np.asarray(almB, dtype=np.complex128))
    nchan = almE.shape[0]
    g = ring_geometry(nside)
    rs = np.asarray(ring_sel, dtype=np.int64)
    c, s = g["cth"][rs], g["sth"][rs]
    s2 = s * s
    FQ = np.zeros((lmax + 1, len(rs), nchan), dtype=np.complex128)
    FU = np.zeros_like(FQ)
    mall = np.arange(lmax + 1)
    for l, lam, lam_prev in lambda_sweep(lmax, c, s):
        if l < 2:
            continue
        m = mall[: l + 1].astype(np.float64)
        nl = 1.0 / np.sqrt((l + 2.0) * (l + 1.0) * l * (l - 1.0))
        glm = np.sqrt((2.0 * l + 1) / (2.0 * l - 1) * (l * l - m * m))
        X1 = (2 * nl) * (-((l - m * m)[:, None] / s2[None, :] + (l * (l - 1) / 2.0)) * lam
                         + (c / s2)[None, :] * glm[:, None] * lam_prev)
        X2 = (2 * nl) * (m[:, None] / s2[None, :]) * (-(l - 1.0) * c[None, :] * lam + glm[:, None] * lam_prev)
        idx = alm_index(lmax, l, mall[: l + 1])
        aE, aB = almE[:, idx].T.copy(), almB[:, idx].T.copy()
        aE[0], aB[0] = aE[0].real, aB[0].real
        FQ[: l + 1] += -(X1[:, :, None] * aE[:, None, :] + 1j * (X2[:, :, None] * aB[:, None, :]))
        FU[: l + 1] += -(X1[:, :, None] * aB[:, None, :] - 1j * (X2[:, :, None] * aE[:, None, :]))
    return _ring_values(FQ, g, rs), _ring_values(FU, g, rs), g["start"][rs]


def map2alm_adjoint_ms(maps, nside, lmax, ms, ring_weights=None):
    """One quadrature pass (``map2alm_adjoint``) for the selected m only: dict m -> (nchan, lmax - m + 1)."""
    maps = np.atleast_2d(np.asarray(maps, dtype=np.float64))
    g = ring_geometry(nside)
    nring = 4 * nside - 1
    wgt = _full_ring_weights(nside, ring_weights)
    ms = [int(m) for m in ms]
    F = {m: np.zeros((nring, maps.shape[0]), dtype=np.complex128) for m in ms}
    for r in range(nring):
        nph, start = int(g["nph"][r]), int(g["start"][r])
        G = np.fft.fft(maps[:, start : start + nph], axis=1)
        for m in ms:
            F[m][r] = G[:, m % nph] * (wgt[r] * np.exp(-1j * m * g["phi0"][r]))
    return {m: (lambda_lm(lmax, m, g["cth"], g["sth"]) @ F[m]).T for m in ms}


def parity_rings(nside, n_extra=12, seed=0):
    """Ring subset for production-size parity checks: both poles, the cap/belt boundaries, the equator
    and a few random rings (0-based ring numbers)."""
    n = int(nside)
    base = [0, 1, 2, 3, n // 2, n - 3, n - 2, n - 1, n, n + 1, 2 * n - 3, 2 * n - 2, 2 * n - 1, 2 * n, 2 * n + 1,
            3 * n - 3, 3 * n - 2, 3 * n - 1, 3 * n, 3 * n + 1, 4 * n - 5, 4 * n - 4, 4 * n - 3, 4 * n - 2]
    rng = np.random.default_rng(seed)
    extra = rng.integers(0, 4 * n - 1, size=n_extra).tolist()
    return np.array(sorted({r for r in base + extra if 0 <= r < 4 * n - 1}), dtype=np.int64)
