"""Oracle restatement of ``cora/signal/corrfunc.py:265-400`` (``legendre_array``, ``corr_to_clarray``): the
xi(r) -> C_l(chi, chi') front end that feeds ``mkfullsky`` from the LSS pipeline (test infrastructure, see
package doc).  Single rank (the reference's ``mpiarray`` transposes are no-ops on one rank).

Parity: PINNED for everything but ``cosine_rule`` by ``tests/golden/corr_to_clarray.npz`` -- the reference's own
``corr_to_clarray`` run in the build container (``tests/golden/make_golden.py``) with this module's
``cosine_rule`` standing in for ``caput.astro.coordinates.spherical.cosine_rule``.  caput (un-pinned git HEAD,
``pyproject.toml:26``) is not under /root/reference and not installable: ``cosine_rule`` restates its documented
behaviour -- the separation of two points at radii x1, x2 whose directions make an angle with cosine mu, in the
cancellation-free form -- and is **parity unpinned**."""

import numpy as np
import scipy.special as ss


def cosine_rule(mu, x1, x2):
    """r[i, a, b] = sqrt(x1_a^2 + x2_b^2 - 2 x1_a x2_b mu_i), evaluated as sqrt((x1 - x2)^2 + 2 x1 x2 (1 - mu))
    (``caput.astro.coordinates.spherical.cosine_rule`` as called at ``corrfunc.py:369``)."""
    mu = np.asarray(mu)[:, np.newaxis, np.newaxis]
    a = np.asarray(x1)[np.newaxis, :, np.newaxis]
    b = np.asarray(x2)[np.newaxis, np.newaxis, :]
    return np.sqrt((a - b) ** 2 + 2.0 * a * b * (1.0 - mu))


def legendre_array(lmax, mu):
    """P_l(mu_i), ``float64[lmax + 1, len(mu)]`` (``corrfunc.py:265-287``: scipy ``lpn`` per node)."""
    lm = np.zeros((lmax + 1, len(mu)), dtype=np.float64)
    lpn = getattr(ss, "lpn", None)     # removed from scipy >= 1.17; legendre_p_all is its replacement
    for i, v in enumerate(mu):
        lm[:, i] = lpn(lmax, v)[0] if lpn is not None else np.asarray(ss.legendre_p_all(lmax, v))[0]
    return lm


def radial_nodes(xarray, xromb, xwidth=None):
    """(xa, x_w, xint): the radial Gauss-Legendre samples of every bin and their normalised weights
    (``corrfunc.py:340-359``).  The half width of bin i is |x_i - x_{i-1}| / 2 (first bin: that of the second)."""
    xarray = np.asarray(xarray, dtype=np.float64)
    if xromb <= 0:
        return xarray, np.ones(1), 1
    if xwidth is None:
        xhalf = np.empty(xarray.shape)
        xhalf[0] = np.abs(xarray[1] - xarray[0]) / 2.0
        xhalf[1:] = np.abs(xarray[1:] - xarray[:-1]) / 2.0
    else:
        xhalf = np.ones(xarray.shape) * xwidth / 2.0
    xint = 2**xromb + 1
    x_r, x_w, x_wsum = ss.roots_legendre(xint, mu=True)
    x_w = x_w / x_wsum
    xa = (xarray[:, np.newaxis] + xhalf[:, np.newaxis] * x_r).flatten()
    return xa, x_w, xint


def corr_to_clarray(corr, lmax, xarray, xromb=3, xwidth=None, q=2, chunksize=50):
    """C_l(x, x') from a real-space correlation function (``corrfunc.py:290-400``): Gauss-Legendre in
    mu = cos(theta) with ``M = q lmax`` nodes (``:332-333``), radial bin average by Gauss-Legendre (``:340-381``),
    contraction with ``P_l(mu_i) w_i 4 pi / sum(w)`` (``:386-395``).  ``float64[lmax + 1, nx, nx]``."""
    xarray = np.asarray(xarray, dtype=np.float64)
    M = q * lmax
    mu, w, wsum = ss.roots_legendre(M, mu=True)
    xa, x_w, xint = radial_nodes(xarray, xromb, xwidth)
    xlen = xarray.size
    corr_array = np.zeros((M, xlen, xlen))
    for msec in np.array_split(np.arange(M), M // chunksize):
        rc = cosine_rule(mu[msec], xa, xa)
        corr1 = corr(rc)
        if xromb > 0:
            corr1 = corr1.reshape(-1, xint)
            corr1 = np.matmul(corr1, x_w).reshape(-1, xlen, xint, xlen)
            corr1 = np.matmul(corr1.transpose(0, 1, 3, 2), x_w)
        corr_array[msec] = corr1
    lm = legendre_array(lmax, mu)
    lm *= w[np.newaxis] * 4.0 * np.pi / wsum
    return np.dot(lm, corr_array.reshape(M, -1)).reshape(lmax + 1, xlen, xlen)
