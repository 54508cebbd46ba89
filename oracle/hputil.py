"""Oracle restatement of the alm packing and inverse-transform wrappers of
``cora/util/hputil.py`` (test infrastructure, see package doc).  The ``healpy.alm2map``
calls are served by ``oracle.sht`` (parity unpinned there)."""

import numpy as np

from . import sht


def unpack_alm(alm, lmax):
    """healpy-packed 1-D -> square [l, m] (``hputil.py:93-121``, ``fullm=False``)."""
    almarray = np.zeros((lmax + 1, lmax + 1), dtype=alm.dtype)
    (almarray.T)[np.triu_indices(lmax + 1)] = alm
    return almarray


def pack_alm(almarray, lmax=None):
    """square [l, m] -> healpy-packed m-major 1-D (``hputil.py:124-152``).

    ``idx(l, m) = m (2 lmax + 1 - m)/2 + l``  <=>  ``almarray.T[triu_indices]``.
    """
    if not lmax:
        lmax = almarray.shape[0] - 1
    return (almarray.T)[np.triu_indices(lmax + 1)]


def sphtrans_inv_real(alm, nside):
    """``hputil.py:369-391``."""
    if alm.shape[1] != alm.shape[0]:
        raise Exception("a_lm array wrong shape.")
    return sht.alm2map(pack_alm(alm), nside)


def sphtrans_inv_real_pol(alm, nside):
    """``hputil.py:394-432``: T scalar, (E,B)->(Q,U) spin-2, V (if present) scalar."""
    npol = alm.shape[0]
    if alm.shape[1] != alm.shape[2] or not (npol == 3 or npol == 4):
        raise Exception("a_lm array wrong shape.")
    maps = np.zeros((npol, 12 * nside**2), dtype=np.float64)
    maps[0] = sht.alm2map(pack_alm(alm[0]), nside)
    q, u = sht.alm2map_spin2(pack_alm(alm[1]), pack_alm(alm[2]), nside)
    maps[1], maps[2] = q[0], u[0]
    if npol == 4:
        maps[3] = sht.alm2map(pack_alm(alm[3]), nside)
    return maps


def sphtrans_inv_sky(alm, nside):
    """``hputil.py:500-531``: loop over frequency; polarised iff ``npol >= 3``."""
    nfreq, npol = alm.shape[0], alm.shape[1]
    pol = npol >= 3
    sky = np.zeros((nfreq, npol, 12 * nside**2), dtype=np.float64)
    for i in range(nfreq):
        if pol:
            sky[i] = sphtrans_inv_real_pol(alm[i], nside)
        else:
            sky[i, 0] = sphtrans_inv_real(alm[i, 0], nside)
    return sky
