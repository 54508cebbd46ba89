"""Oracle restatement of ``cora/core/skysim.py`` ``clarray`` and ``mkfullsky``
(test infrastructure, see package doc).  Single rank; the MPI partition is restated in
``partition`` for the multi-GPU tests."""

import numpy as np
import scipy.integrate as si

from . import hputil, nputil


def romberg_weights(zromb):
    """Weights w with ``si.romb(y, dx) == dx * w . y`` for ``2**zromb + 1`` samples."""
    n = 2**zromb + 1
    return np.array([si.romb(np.eye(n)[i], dx=1.0) for i in range(n)])


def clarray(aps, lmax, zarray, zromb=3, zwidth=None):
    """Channel-averaged ``C_l(z, z')`` table (``skysim.py:10-69``).

    ``zromb == 0``: point evaluation (``:33-38``).  Otherwise each channel is sampled at
    ``2**zromb + 1`` points across its width and Romberg-integrated in both arguments
    (``:41-67``), l processed in ``lmax // 5`` sections.
    """
    zarray = np.asarray(zarray)
    if zromb == 0:
        return aps(np.arange(lmax + 1)[:, None, None], zarray[None, :, None], zarray[None, None, :])
    zsort = np.sort(zarray)
    zhalf = np.abs(zsort[1] - zsort[0]) / 2.0 if zwidth is None else zwidth / 2.0
    zlen = zarray.size
    zint = 2**zromb + 1
    zspace = 2.0 * zhalf / 2**zromb
    za = (zarray[:, None] + np.linspace(-zhalf, zhalf, zint)[None, :]).flatten()
    cla = np.zeros((lmax + 1, zlen, zlen), dtype=np.float64)
    for lsec in np.array_split(np.arange(lmax + 1), lmax // 5):
        clt = aps(lsec[:, None, None], za[None, :, None], za[None, None, :])
        clt = clt.reshape(-1, zlen, zint, zlen, zint)
        clt = si.romb(clt, dx=zspace, axis=4)
        clt = si.romb(clt, dx=zspace, axis=2)
        cla[lsec] = clt / (2 * zhalf) ** 2
    return cla


def partition(n, size, rank):
    """caput.mpiarray contiguous block split: first ``n % size`` ranks get one extra."""
    base, rem = divmod(n, size)
    num = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, start + num


def mkfullsky(corr, nside, alms=False, rng=None, record=None):
    """Correlated full-sky Gaussian maps (``skysim.py:72-136``), single rank.

    Per l: ``corr[l] + 1e-14 max(diag) I`` (``:116-117``) -> ``matrix_root_manynull(truncate=False)``
    (``:119``) -> ``complex_std_normal((numz, l+1))`` (``:120``) -> ``alm[:, 0, l, :l+1] = M g``
    (``:121``).  ``record`` (a dict) captures the roots and draws for injected-draw parity.
    """
    numz = corr.shape[1]
    maxl = corr.shape[0] - 1
    if corr.shape[2] != numz:
        raise Exception("Correlation matrix is incorrect shape.")
    alm_array = np.zeros((numz, 1, maxl + 1, maxl + 1), dtype=np.complex128)
    for l in range(maxl + 1):
        cmax = corr[l].diagonal().max() * 1e-14
        corrm = corr[l] + np.identity(numz) * cmax
        trans = nputil.matrix_root_manynull(corrm, truncate=False)
        gaussvars = nputil.complex_std_normal((numz, l + 1), rng=rng)
        if record is not None:
            record.setdefault("roots", []).append(trans)
            record.setdefault("gauss", []).append(gaussvars)
        alm_array[:, 0, l, : (l + 1)] = np.dot(trans, gaussvars)
    if alms:
        return alm_array
    return hputil.sphtrans_inv_sky(alm_array, nside)[:, 0]


def mkconstrained(corr, constraints, nside):
    """Restatement of ``cora/core/skysim.py:139-201`` with the oracle transforms in place of healpy
    (``map2alm`` defaults: no ring weights, ``iter=3``)."""
    import scipy.linalg as la

    from . import sht

    numz = corr.shape[1]
    maxl = corr.shape[0] - 1
    nmodes = len(constraints)
    f_ind = [c[0] for c in constraints]
    if corr.shape[2] != numz:
        raise Exception("Correlation matrix is incorrect shape.")
    nalm = (maxl + 1) * (maxl + 2) // 2
    larr = np.concatenate([np.arange(m, maxl + 1) for m in range(maxl + 1)])   # healpy.Alm.getlm order
    trans = np.zeros((corr.shape[0], nmodes, numz))
    tmat = np.zeros((corr.shape[0], nmodes, nmodes))
    cmap = np.zeros((nalm, nmodes), dtype=np.complex128)
    cv = np.zeros((numz, nalm), dtype=np.complex128)
    for i in range(maxl + 1):
        trans[i] = la.eigh(corr[i])[1][:, -nmodes:].T
        tmat[i] = trans[i][:, f_ind]
    for i, cons in enumerate(constraints):
        cmap[:, i] = sht.map2alm(cons[1], nside, maxl, iter=3)
    for i, l in enumerate(larr):
        if l == 0:
            cv[:, i] = 0.0
        else:
            cv[:, i] = np.dot(trans[l].T, la.solve(tmat[l].T, cmap[i]))
    return sht.alm2map(cv, nside, maxl)
