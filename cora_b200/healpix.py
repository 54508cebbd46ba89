"""HEALPix pixel utilities on the host (numpy): the healpy calls that ``ConstrainedGalaxy`` and the coordinate rotation
make around the hot path (``cora/foreground/galaxy.py:43-55,133-207``, ``cora/util/hputil.py:534-604``):
``reorder`` (RING <-> NESTED), ``ud_grade``, ``pix2ang`` / ``ang2vec``, ``get_interp_weights`` / ``get_interp_val``
(bilinear interpolation on the rings) and the Galactic <-> celestial rotation.

healpy itself (``healpy>=1.17``, ``pyproject.toml:30``) is a third-party dependency that is neither under
/root/reference nor installable here: everything below restates the published HEALPix definitions (Gorski et al. 2005:
the 12 base faces, the (ix, iy) bit-interleaved NESTED index inside a face, the RING geometry of SURVEY App. A.8) and is
validated geometrically in ``tests/test_healpix.py`` (NESTED parents enclose their children, interpolation reproduces
pixel centres and smooth functions, rotations are orthogonal and map the Galactic poles to their J2000 positions):
**parity with healpy unpinned**.
"""

import numpy as np

_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4])
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7])


def nside2npix(nside):
    return 12 * int(nside) ** 2


def npix2nside(npix):
    nside = int(round(np.sqrt(npix / 12.0)))
    if 12 * nside * nside != npix:
        raise ValueError("not a HEALPix map size: %d" % npix)
    return nside


def _spread_bits(v):
    """Insert a zero bit between the bits of v (v < 2^16 per call chunk; handles up to 2^29 by two passes)."""
    v = np.asarray(v, dtype=np.int64)
    out = np.zeros_like(v)
    for b in range(30):
        out |= ((v >> b) & 1) << (2 * b)
    return out


def _compress_bits(v):
    """Inverse of ``_spread_bits``: keep the even bits."""
    v = np.asarray(v, dtype=np.int64)
    out = np.zeros_like(v)
    for b in range(30):
        out |= ((v >> (2 * b)) & 1) << b
    return out


def _ring2xyf(nside, pix):
    """RING pixel index -> (ix, iy, face): position inside one of the 12 base faces."""
    pix = np.asarray(pix, dtype=np.int64)
    ncap = 2 * nside * (nside - 1)
    npix = 12 * nside * nside
    nl2 = 2 * nside
    iring = np.empty_like(pix)
    iphi = np.empty_like(pix)
    kshift = np.zeros_like(pix)
    nr = np.empty_like(pix)
    face = np.empty_like(pix)

    north = pix < ncap
    south = pix >= npix - ncap
    belt = ~(north | south)

    p = pix[north]
    ir = (1 + np.floor(np.sqrt(1 + 2 * p.astype(np.float64))).astype(np.int64)) >> 1
    ir = np.where(2 * ir * (ir - 1) > p, ir - 1, ir)            # guard the float square root
    ir = np.where(2 * (ir + 1) * ir <= p, ir + 1, ir)
    ip = p + 1 - 2 * ir * (ir - 1)
    iring[north], iphi[north], nr[north] = ir, ip, ir
    face[north] = (ip - 1) // ir

    p = pix[belt] - ncap
    tmp = p // (4 * nside)
    ir = tmp + nside
    ip = p - tmp * 4 * nside + 1
    iring[belt], iphi[belt], nr[belt] = ir, ip, nside
    kshift[belt] = (ir + nside) & 1
    ire = ir - nside + 1
    irm = nl2 + 2 - ire
    ifm = (ip - ire // 2 + nside - 1) // nside
    ifp = (ip - irm // 2 + nside - 1) // nside
    face[belt] = np.where(ifp == ifm, ifp | 4, np.where(ifp < ifm, ifp, ifm + 8))

    p = npix - pix[south]
    ir = (1 + np.floor(np.sqrt(2 * p.astype(np.float64) - 1)).astype(np.int64)) >> 1
    ir = np.where(2 * ir * (ir - 1) >= p, ir - 1, ir)
    ir = np.where(2 * (ir + 1) * ir < p, ir + 1, ir)
    ip = 4 * ir + 1 - (p - 2 * ir * (ir - 1))
    iring[south], iphi[south], nr[south] = 2 * nl2 - ir, ip, ir
    face[south] = 8 + (ip - 1) // ir

    irt = iring - _JRLL[face] * nside + 1
    ipt = 2 * iphi - _JPLL[face] * nr - kshift - 1
    ipt = np.where(ipt >= nl2, ipt - 8 * nside, ipt)
    ix = (ipt - irt) >> 1
    iy = (-(ipt + irt)) >> 1
    return ix, iy, face


def _xyf2ring(nside, ix, iy, face):
    """(ix, iy, face) -> RING pixel index."""
    nl4 = 4 * nside
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    jr = _JRLL[face] * nside - ix - iy - 1
    north = jr < nside
    south = jr > 3 * nside
    nr = np.where(north, jr, np.where(south, nl4 - jr, nside))
    n_before = np.where(north, 2 * nr * (nr - 1), np.where(south, npix - 2 * (nr + 1) * nr, ncap + (jr - nside) * nl4))
    kshift = np.where(north | south, 0, (jr - nside) & 1)
    jp = (_JPLL[face] * nr + ix - iy + 1 + kshift) // 2
    jp = np.where(jp > nl4, jp - nl4, jp)
    jp = np.where(jp < 1, jp + nl4, jp)
    return n_before + jp - 1


def ring2nest(nside, ipix):
    """``healpy.ring2nest``."""
    ix, iy, face = _ring2xyf(int(nside), ipix)
    return face * int(nside) ** 2 + _spread_bits(ix) + (_spread_bits(iy) << 1)


def nest2ring(nside, ipix):
    """``healpy.nest2ring``."""
    nside = int(nside)
    ipix = np.asarray(ipix, dtype=np.int64)
    face = ipix // (nside * nside)
    p = ipix % (nside * nside)
    return _xyf2ring(nside, _compress_bits(p), _compress_bits(p >> 1), face)


def reorder(hpmap, r2n=False, n2r=False):
    """``healpy.reorder``: RING -> NESTED (``r2n``) or NESTED -> RING (``n2r``) along the last axis."""
    if r2n == n2r:
        raise ValueError("exactly one of r2n / n2r")
    hpmap = np.asarray(hpmap)
    nside = npix2nside(hpmap.shape[-1])
    idx = np.arange(hpmap.shape[-1], dtype=np.int64)
    if r2n:      # out[nest] = in[ring(nest)]
        return hpmap[..., nest2ring(nside, idx)]
    return hpmap[..., ring2nest(nside, idx)]


def ud_grade(hpmap, nside_out):
    """``healpy.ud_grade`` for RING maps with its defaults (``power=None``): degrading averages the children of every
    coarse pixel, upgrading copies the parent's value into its children.  Works along the last axis."""
    hpmap = np.asarray(hpmap, dtype=np.float64)
    nside_in = npix2nside(hpmap.shape[-1])
    nside_out = int(nside_out)
    if nside_out == nside_in:
        return hpmap.copy()
    nest = reorder(hpmap, r2n=True)
    if nside_out < nside_in:
        ratio = (nside_in // nside_out) ** 2
        out = nest.reshape(nest.shape[:-1] + (-1, ratio)).mean(axis=-1)
    else:
        ratio = (nside_out // nside_in) ** 2
        out = np.repeat(nest, ratio, axis=-1)
    return reorder(out, n2r=True)


def ring_info(nside):
    """Per-ring geometry of the RING scheme: cos(theta), pixels per ring, first pixel, phi of the
    first pixel centre; arrays of length ``4 nside - 1``."""
    nside = int(nside)
    i = np.arange(1, 4 * nside, dtype=np.int64)
    ip = np.where(i > 2 * nside, 4 * nside - i, i)               # mirror ring
    cap = ip < nside
    nph = np.where(cap, 4 * ip, 4 * nside)
    z = np.where(cap, 1.0 - ip.astype(np.float64) ** 2 / (3.0 * nside * nside), 4.0 / 3.0 - 2.0 * ip / (3.0 * nside))
    z = np.where(i > 2 * nside, -z, z)
    shifted = np.where(cap, 1, ((ip - nside) & 1) == 0)
    phi0 = np.where(shifted == 1, np.pi / nph, 0.0)
    start = np.concatenate([[0], np.cumsum(nph)[:-1]])
    return z, nph, start, phi0


def pix2ang(nside, ipix=None):
    """``healpy.pix2ang`` (RING): (theta, phi) of the pixel centres (all pixels by default)."""
    z, nph, start, phi0 = ring_info(nside)
    if ipix is None:
        ipix = np.arange(nside2npix(nside), dtype=np.int64)
    ipix = np.asarray(ipix, dtype=np.int64)
    ring = np.searchsorted(start, ipix, side="right") - 1
    j = ipix - start[ring]
    return np.arccos(z[ring]), phi0[ring] + 2.0 * np.pi * j / nph[ring]


def ang2vec(theta, phi):
    st = np.sin(theta)
    return np.stack([st * np.cos(phi), st * np.sin(phi), np.cos(theta)], axis=-1)


def vec2ang(v):
    v = np.asarray(v, dtype=np.float64)
    theta = np.arctan2(np.hypot(v[..., 0], v[..., 1]), v[..., 2])
    phi = np.arctan2(v[..., 1], v[..., 0])
    return theta, np.where(phi < 0, phi + 2.0 * np.pi, phi)


def get_interp_weights(nside, theta, phi):
    """``healpy.get_interp_weights`` (RING): the four pixels around every direction -- two neighbours in phi on the ring
    above and two on the ring below -- and their bilinear weights, ``(pix[4, n], wgt[4, n])``.  Directions beyond the
    first / last ring interpolate between that ring's two neighbours and the polar cap's four-pixel mean, as healpy
    does."""
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    phi = np.atleast_1d(np.asarray(phi, dtype=np.float64)) % (2.0 * np.pi)
    nside = int(nside)
    z, nph, start, phi0 = ring_info(nside)
    th_ring = np.arccos(z)
    nring = 4 * nside - 1
    zz = np.cos(theta)
    # ring above (ir1: largest ring index with theta_ring <= theta; -1 above the first ring) and below
    ir1 = np.searchsorted(th_ring, theta, side="right") - 1
    ir2 = ir1 + 1

    def ring_pair(ir):
        irc = np.clip(ir, 0, nring - 1)
        n = nph[irc]
        dphi = 2.0 * np.pi / n
        t = (phi - phi0[irc]) / dphi
        i1 = np.floor(t).astype(np.int64)
        w = t - i1
        i2 = i1 + 1
        i1 = np.where(i1 < 0, i1 + n, i1) % n
        i2 = i2 % n
        return start[irc] + i1, start[irc] + i2, w

    a1, a2, wa = ring_pair(ir1)
    b1, b2, wb = ring_pair(ir2)
    pix = np.stack([a1, a2, b1, b2])
    th1 = th_ring[np.clip(ir1, 0, nring - 1)]
    th2 = th_ring[np.clip(ir2, 0, nring - 1)]
    inside = (ir1 >= 0) & (ir2 <= nring - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        wth = np.where(inside, (theta - th1) / (th2 - th1), 0.0)
    wgt = np.stack([(1 - wa) * (1 - wth), wa * (1 - wth), (1 - wb) * wth, wb * wth])
    # polar caps: between the pole (mean of the four pixels of the extreme ring) and that ring
    north = ir1 < 0
    if north.any():
        wt = theta[north] / th_ring[0]
        p4 = np.arange(4, dtype=np.int64)
        pix[:, north] = np.stack([b1[north], b2[north], (b1[north] + 2) % 4, (b2[north] + 2) % 4])
        f = (1.0 - wt) * 0.25
        wgt[:, north] = np.stack([(1 - wb[north]) * wt + f, wb[north] * wt + f, f, f])
        del p4
    south = ir2 > nring - 1
    if south.any():
        npix = 12 * nside * nside
        wt = (np.pi - theta[south]) / (np.pi - th_ring[-1])
        base = npix - 4
        pix[:, south] = np.stack([a1[south], a2[south], base + (a1[south] - base + 2) % 4, base + (a2[south] - base + 2) % 4])
        f = (1.0 - wt) * 0.25
        wgt[:, south] = np.stack([(1 - wa[south]) * wt + f, wa[south] * wt + f, f, f])
    del zz
    return pix, wgt


def get_interp_val(hpmap, theta, phi):
    """``healpy.get_interp_val`` (RING): bilinear interpolation of a map (or of maps along the leading axes)."""
    hpmap = np.asarray(hpmap)
    pix, wgt = get_interp_weights(npix2nside(hpmap.shape[-1]), theta, phi)
    return np.sum(hpmap[..., pix] * wgt, axis=-2)


# ---- coordinate systems -----------------------------------------------------------------------------------------
# Galactic <-> celestial (equatorial, J2000): the IAU definition of the Galactic system carried to J2000
# (north Galactic pole at RA 192.85948 deg, Dec +27.12825 deg; Galactic longitude of the north celestial pole
# 122.93192 deg).  Rows of G2C^T are the Galactic axes in equatorial coordinates.
_NGP_RA, _NGP_DEC, _L_NCP = np.radians(192.85948), np.radians(27.12825), np.radians(122.93192)


def _gal_to_equ_matrix():
    pole = ang2vec(np.pi / 2 - _NGP_DEC, _NGP_RA)                      # Galactic z axis in equatorial coordinates
    ncp = np.array([0.0, 0.0, 1.0])
    # Galactic x axis: in the plane, at longitude 0; the NCP lies at longitude _L_NCP
    e_l = ncp - pole * np.dot(ncp, pole)
    e_l /= np.linalg.norm(e_l)                                          # direction of longitude _L_NCP in the plane
    e_m = np.cross(pole, e_l)                                           # longitude _L_NCP + 90 deg
    x = np.cos(_L_NCP) * e_l - np.sin(_L_NCP) * e_m
    y = np.sin(_L_NCP) * e_l + np.cos(_L_NCP) * e_m
    return np.stack([x, y, pole], axis=1)                               # columns = Galactic axes -> v_equ = M v_gal


def rotation_matrix(coord_from, coord_to):
    """3 x 3 matrix taking unit vectors of system ``coord_from`` to ``coord_to`` ('G' Galactic, 'C' celestial)."""
    m = _gal_to_equ_matrix()
    if coord_from == coord_to:
        return np.identity(3)
    if (coord_from, coord_to) == ("G", "C"):
        return m
    if (coord_from, coord_to) == ("C", "G"):
        return m.T
    raise Exception("Co-ordinate system invalid.")          # (ecliptic 'E' is not needed by the path)


def rotate_angles(theta, phi, coord_from, coord_to):
    """``healpy.Rotator(coord=[from, to])(theta, phi)``."""
    v = ang2vec(np.asarray(theta, dtype=np.float64), np.asarray(phi, dtype=np.float64))
    return vec2ang(v @ rotation_matrix(coord_from, coord_to).T)
