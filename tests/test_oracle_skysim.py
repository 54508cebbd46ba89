"""Pin the root / draw / apply / packing oracle against fixtures produced by the real
reference (tests/golden/make_golden.py)."""

import numpy as np
import pytest

from conftest import golden
from oracle import hputil, nputil, skysim


def test_root_cases():
    g = golden("root_cases.npz")
    r = nputil.matrix_root_manynull(g["spd"].copy(), truncate=False)
    np.testing.assert_allclose(r, g["root_spd"], rtol=1e-13, atol=1e-15)
    assert np.all(np.triu(r, 1) == 0)  # Cholesky branch: lower triangular
    # eigh branch (rank 4 of 12): eigenvector signs are not unique -> compare M M^T, rank, layout
    low = g["lowrank"]
    r = nputil.matrix_root_manynull(low.copy(), truncate=False)
    assert r.shape == (12, 12)
    assert np.count_nonzero(np.abs(r).sum(axis=0)) == int(g["num_pos"]) == 4
    assert np.all(r[:, :8] == 0)  # clipped columns first (ascending eigenvalues), kept as zeros
    np.testing.assert_allclose(r @ r.T, low, atol=1e-13 * np.abs(low).max())
    np.testing.assert_allclose(np.abs(r), np.abs(g["root_lowrank"]), atol=1e-12)
    rt, npos = nputil.matrix_root_manynull(low.copy())
    assert npos == 4 and rt.shape == (1, 12, 4)  # reference shape quirk, see oracle/nputil.py
    np.testing.assert_allclose(np.abs(rt), np.abs(g["root_lowrank_trunc"]), atol=1e-12)


def test_complex_std_normal_stream_order():
    g = golden("root_cases.npz")
    v = nputil.complex_std_normal((5, 7), rng=np.random.default_rng(11))
    np.testing.assert_array_equal(v, g["cstd_seed11"])
    # real block first, then imaginary block
    rng = np.random.default_rng(11)
    re = rng.standard_normal((5, 7))
    im = rng.standard_normal((5, 7))
    np.testing.assert_array_equal(v, (re + 1j * im) / 2**0.5)


def test_alm_packing():
    g = golden("alm_pack.npz")
    np.testing.assert_array_equal(hputil.pack_alm(g["square"]), g["packed"])
    np.testing.assert_array_equal(hputil.unpack_alm(g["packed"], 5), g["unpacked"])
    lmax = 5
    for l in range(lmax + 1):
        for m in range(l + 1):
            assert g["packed"][m * (2 * lmax + 1 - m) // 2 + l] == g["square"][l, m]


def _check_mkfullsky(name):
    g = golden(name)
    rec = {}
    alm = skysim.mkfullsky(g["cl"], int(g["nside"]), alms=True, rng=np.random.default_rng(0), record=rec)
    assert alm.shape == g["alm"].shape
    scale = np.abs(g["alm"]).max()
    assert np.max(np.abs(alm - g["alm"])) / scale < 1e-12
    L = g["cl"].shape[0]
    for l in range(L):
        np.testing.assert_array_equal(rec["gauss"][l], g["gauss"][l][:, : l + 1])
        np.testing.assert_allclose(rec["roots"][l], g["roots"][l], rtol=1e-12, atol=1e-14 * np.abs(g["roots"][l]).max())
    # zeros above the diagonal, m <= l only
    for l in range(L):
        assert np.all(alm[:, 0, l, l + 1 :] == 0)
    return g, rec


def test_mkfullsky_alms_sck():
    g, rec = _check_mkfullsky("mkfullsky_sck.npz")
    # l = 0 of the SCK table is the zero matrix -> Cholesky fails -> eigh -> root = 0
    assert np.all(rec["roots"][0] == 0)


def test_mkfullsky_alms_21cm():
    _check_mkfullsky("mkfullsky_21cm.npz")


def test_mkfullsky_alms_polarised_block():
    g = golden("mkfullsky_pol.npz")
    alm = skysim.mkfullsky(g["cl"], 4, alms=True, rng=np.random.default_rng(3))
    scale = np.abs(g["alm"]).max()
    assert np.max(np.abs(alm - g["alm"])) / scale < 1e-12
    nf = int(g["nfreq"])
    # Cholesky path with few channels: V rows = sqrt(cmax) g  (SURVEY App. C.6): tiny, non-zero
    v = alm[3 * nf :, 0, 5, :6]
    t = alm[:nf, 0, 5, :6]
    assert 0 < np.abs(v).max() < 1e-5 * np.abs(t).max()


def test_mkfullsky_map_shape_and_m0_quirk():
    g = golden("mkfullsky_sck.npz")
    nside = int(g["nside"])
    sky = skysim.mkfullsky(g["cl"], nside, rng=np.random.default_rng(0))
    assert sky.shape == (g["cl"].shape[1], 12 * nside**2)
    # same alm with Im(a_l0) removed gives the identical map
    alm = g["alm"].copy()
    alm[:, :, :, 0] = alm[:, :, :, 0].real
    sky2 = hputil.sphtrans_inv_sky(alm, nside)[:, 0]
    np.testing.assert_allclose(sky, sky2, atol=1e-13 * np.abs(sky).max())


def test_mkfullsky_bad_shape():
    import pytest

    with pytest.raises(Exception, match="incorrect shape"):
        skysim.mkfullsky(np.zeros((4, 3, 2)), 2)


def test_partition_matches_caput_rule():
    # first n % P ranks get one extra item, contiguous blocks
    for n, P in ((10, 4), (8, 8), (7, 2), (3, 4)):
        blocks = [skysim.partition(n, P, r) for r in range(P)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        sizes = [b - a for a, b in blocks]
        assert sizes == [n // P + (1 if r < n % P else 0) for r in range(P)]
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(P - 1))


def test_oracle_mkconstrained_reproduces_constraints():
    """Restatement of skysim.py:139-201: the constrained frequency slices come back as the
    (monopole-free, band-limited) constraint maps."""
    from oracle import sht
    from oracle import spectra as osp2

    nside, nz = 8, 5
    lmax = 3 * nside - 1
    freq = np.linspace(800.0, 400.0, nz, endpoint=False)
    cl = skysim.clarray(osp2.full_sky_synchrotron().angular_powerspectrum, lmax, freq)
    cl[0] = cl[1]
    rng = np.random.default_rng(3)
    maps = []
    for _ in range(2):
        lb = lmax // 2
        nalm = (lb + 1) * (lb + 2) // 2
        a = rng.standard_normal(nalm) + 1j * rng.standard_normal(nalm)
        a[: lb + 1] = a[: lb + 1].real
        maps.append(sht.alm2map(a, nside, lb))
    out = skysim.mkconstrained(cl, [[0, maps[0]], [3, maps[1]]], nside)
    assert out.shape == (nz, 12 * nside**2)
    for idx, mp in ((0, maps[0]), (3, maps[1])):
        assert np.abs(out[idx] - (mp - mp.mean())).max() / np.abs(mp).max() < 2e-2
    with pytest.raises(Exception, match="incorrect shape"):
        skysim.mkconstrained(np.zeros((4, 3, 2)), [[0, maps[0]]], nside)
