"""Host HEALPix utilities (cora_b200/healpix.py) that stand in for the healpy calls around the path
(reorder / ud_grade / get_interp_val / Rotator).  healpy is absent: the checks are the worked examples printed in
healpy's own docstrings (quoted from its documentation), permutation / geometry invariants and analytic cases."""

import numpy as np
import pytest

from cora_b200 import healpix as hpx


def test_ring_nest_known_answers_from_healpy_docs():
    # healpy.ring2nest / nest2ring docstring examples
    assert hpx.ring2nest(16, np.array([1504]))[0] == 1130
    assert hpx.nest2ring(16, np.array([1130]))[0] == 1504
    np.testing.assert_array_equal(hpx.ring2nest(2, np.arange(10)), [3, 7, 11, 15, 2, 1, 6, 5, 10, 9])
    np.testing.assert_array_equal(hpx.nest2ring(2, np.arange(10)), [13, 5, 4, 0, 15, 7, 6, 1, 17, 9])
    np.testing.assert_array_equal(hpx.ring2nest(1, np.arange(12)), np.arange(12))


@pytest.mark.parametrize("nside", [1, 2, 4, 8, 16, 64])
def test_ring_nest_are_inverse_permutations(nside):
    npix = 12 * nside * nside
    idx = np.arange(npix)
    nest = hpx.ring2nest(nside, idx)
    assert np.array_equal(np.sort(nest), idx)
    assert np.array_equal(hpx.nest2ring(nside, nest), idx)
    m = np.random.default_rng(nside).standard_normal(npix)
    np.testing.assert_array_equal(hpx.reorder(hpx.reorder(m, r2n=True), n2r=True), m)


@pytest.mark.parametrize("nside", [2, 8, 32])
def test_nested_children_sit_inside_their_parent(nside):
    """NESTED pixels 4p .. 4p+3 at nside are the children of pixel p at nside / 2: their centres lie within the
    parent's angular size of the parent's centre, and their mean direction is the parent's centre to second order."""
    th, ph = hpx.pix2ang(nside)
    v = hpx.ang2vec(th, ph)[hpx.nest2ring(nside, np.arange(12 * nside * nside))]          # child centres in NESTED order
    thp, php = hpx.pix2ang(nside // 2)
    vp = hpx.ang2vec(thp, php)[hpx.nest2ring(nside // 2, np.arange(3 * nside * nside))]   # parents in NESTED order
    kids = v.reshape(-1, 4, 3)
    size = np.sqrt(4 * np.pi / (3 * nside * nside))          # parent pixel scale
    d = np.arccos(np.clip(np.einsum("pkc,pc->pk", kids, vp), -1, 1))
    assert d.max() < 0.85 * size
    mean = kids.mean(axis=1)
    mean /= np.linalg.norm(mean, axis=1)[:, None]
    assert np.arccos(np.clip(np.einsum("pc,pc->p", mean, vp), -1, 1)).max() < 0.2 * size


def test_pix2ang_matches_the_sht_plan_geometry():
    from oracle import sht as osht

    for nside in (1, 4, 16):
        th, ph = hpx.pix2ang(nside)
        tho, pho = osht.pix2ang_ring(nside)
        np.testing.assert_allclose(th, tho, rtol=0, atol=1e-14)
        np.testing.assert_allclose(ph, pho, rtol=0, atol=1e-14)


def test_ud_grade():
    rng = np.random.default_rng(3)
    m = rng.standard_normal(12 * 16 * 16)
    lo = hpx.ud_grade(m, 4)
    assert lo.shape == (12 * 16,)
    np.testing.assert_allclose(lo.mean(), m.mean(), atol=1e-14)          # averaging preserves the mean
    up = hpx.ud_grade(lo, 16)
    np.testing.assert_allclose(hpx.ud_grade(up, 4), lo, atol=1e-14)      # upgrade copies, degrade averages
    np.testing.assert_array_equal(hpx.ud_grade(np.full(48, 2.5), 8), np.full(768, 2.5))
    # a smooth function: the degraded map is the function at the coarse centres up to the pixel size squared
    th, ph = hpx.pix2ang(32)
    f = np.cos(th) + 0.3 * np.sin(th) * np.cos(ph)
    thc, phc = hpx.pix2ang(8)
    np.testing.assert_allclose(hpx.ud_grade(f, 8), np.cos(thc) + 0.3 * np.sin(thc) * np.cos(phc), atol=6e-3)
    stack = hpx.ud_grade(np.stack([f, 2 * f]), 8)
    np.testing.assert_allclose(stack[1], 2 * stack[0], atol=1e-14)


def test_get_interp_val_known_answers_from_healpy_docs():
    m = np.arange(12.0)
    assert hpx.get_interp_val(m, np.pi / 2, 0)[0] == pytest.approx(4.0)
    assert hpx.get_interp_val(m, np.pi / 2, np.pi / 2)[0] == pytest.approx(5.0)
    assert hpx.get_interp_val(m, np.pi / 2, np.pi / 2 + 2 * np.pi)[0] == pytest.approx(5.0)
    got = hpx.get_interp_val(m, np.linspace(0, np.pi, 10), 0)
    np.testing.assert_allclose(got, [1.5, 1.5, 1.5, 2.20618428, 3.40206143, 5.31546486, 7.94639458, 9.5, 9.5, 9.5], atol=1e-7)


@pytest.mark.parametrize("nside", [4, 16])
def test_get_interp_val_geometry(nside):
    th, ph = hpx.pix2ang(nside)
    m = np.random.default_rng(nside).standard_normal(th.size)
    np.testing.assert_allclose(hpx.get_interp_val(m, th, ph), m, atol=1e-12)        # exact at the pixel centres
    rng = np.random.default_rng(1)
    t, p = np.arccos(rng.uniform(-1, 1, 4000)), rng.uniform(0, 2 * np.pi, 4000)
    pix, wgt = hpx.get_interp_weights(nside, t, p)
    np.testing.assert_allclose(wgt.sum(axis=0), 1.0, atol=1e-12)
    assert wgt.min() >= -1e-12 and pix.min() >= 0 and pix.max() < 12 * nside * nside
    f = lambda a, b: np.cos(a) + 0.4 * np.sin(a) * np.sin(b)
    err = np.abs(hpx.get_interp_val(f(th, ph), t, p) - f(t, p)).max()
    assert err < 2.5 / nside**2 + 0.02 / nside, err


def test_galactic_celestial_rotation():
    m = hpx.rotation_matrix("G", "C")
    np.testing.assert_allclose(m @ m.T, np.identity(3), atol=1e-14)
    assert np.linalg.det(m) == pytest.approx(1.0)
    # north Galactic pole -> RA 192.85948, Dec 27.12825 (J2000); Galactic centre -> RA 266.405, Dec -28.936
    th, ph = hpx.rotate_angles(0.0, 0.0, "G", "C")
    assert np.degrees(ph) == pytest.approx(192.85948, abs=1e-6) and 90 - np.degrees(th) == pytest.approx(27.12825, abs=1e-6)
    th, ph = hpx.rotate_angles(np.pi / 2, 0.0, "G", "C")
    assert np.degrees(ph) == pytest.approx(266.405, abs=2e-3) and 90 - np.degrees(th) == pytest.approx(-28.936, abs=2e-3)
    # and back
    t0, p0 = np.array([0.3, 1.2, 2.9]), np.array([0.1, 3.0, 6.0])
    t1, p1 = hpx.rotate_angles(*hpx.rotate_angles(t0, p0, "G", "C"), "C", "G")
    np.testing.assert_allclose(t1, t0, atol=1e-13)
    np.testing.assert_allclose(p1, p0, atol=1e-13)
    with pytest.raises(Exception, match="Co-ordinate system invalid"):
        hpx.rotation_matrix("G", "X")


def test_coord_rotation_of_a_dipole():
    from cora_b200 import healpix, hputil      # (coord_x2y is host code: no GPU needed)

    nside = 32
    ang = hputil.ang_positions(nside)
    v = healpix.ang2vec(ang[:, 0], ang[:, 1])
    a = np.array([0.3, -0.5, 0.8])
    gal = np.stack([v @ a, 2.0 + (v @ a) ** 2])
    cel = hputil.coord_g2c(gal.copy())
    M = healpix.rotation_matrix("G", "C")
    want = v @ (M @ a)                       # f_out(v_c) = a . (M^T v_c)
    # (bilinear interpolation: second order in the pixel size, coarser next to the poles where a ring has 4 .. 8 pixels)
    err = np.abs(cel[0] - want)
    assert err.max() < 5e-3 and np.mean(err > 1e-3) < 0.005
    back = hputil.coord_c2g(cel.copy())
    assert np.max(np.abs(back[0] - gal[0])) < 1e-2
    with pytest.raises(Exception, match="Co-ordinate system invalid"):
        hputil.coord_x2y(gal, "G", "Q")


def test_pix2ang_known_answers_from_healpy_docs():
    # healpy.pix2ang / pix2vec docstring examples (nside 16, RING)
    th, ph = hpx.pix2ang(16, np.array([1440, 427, 1520, 0, 3068]))
    np.testing.assert_allclose(th, [1.52911759, 0.78550497, 1.57079633, 0.05103658, 3.09055608], atol=5e-9)
    np.testing.assert_allclose(ph, [0.0, 0.78539816, 1.61988371, 0.78539816, 0.78539816], atol=5e-9)
    v = hpx.ang2vec(*hpx.pix2ang(16, np.array([1504])))[0]
    np.testing.assert_allclose(v, [0.99879545620517241, 0.049067674327418015, 0.0], atol=1e-15)
