"""Host-side logic of the drop-in layer that needs no device: the channel axis of cora-makesky
(makesky.py:44-92), the Romberg weights clarray folds into the fill kernels (skysim.py:62-67), alm
packing (hputil.py:93-193) and the shard plan's bookkeeping."""

import numpy as np
import pytest
import scipy.integrate as si


def test_freqstate_modes():
    from cora_b200 import makesky

    fs = makesky.FreqState()
    assert fs.freq == (800.0, 400.0, 1025) and fs.freq_mode == "centre"
    fs.freq = (800.0, 400.0, 4)
    np.testing.assert_allclose(fs.frequencies, [800.0, 700.0, 600.0, 500.0])
    assert fs.freq_width == 100.0
    fs.freq_mode = "centre_nyquist"
    np.testing.assert_allclose(fs.frequencies, np.linspace(800.0, 400.0, 4))
    assert fs.freq_width == pytest.approx(400.0 / 3)
    fs.freq_mode = "edge"
    np.testing.assert_allclose(fs.frequencies, [750.0, 650.0, 550.0, 450.0])
    assert fs.freq_width == -100.0          # the reference keeps the sign in "edge" mode (makesky.py:82)
    fs.freq_mode = "centre"
    fs.channel_bin = 2
    np.testing.assert_allclose(fs.frequencies, [750.0, 550.0])
    assert fs.freq_width == 200.0
    fs.channel_bin = 1
    fs.channel_range = (1, 3)
    np.testing.assert_allclose(fs.frequencies, [700.0, 600.0])
    fs.channel_list = [3, 0]                # the list wins over the range
    np.testing.assert_allclose(fs.frequencies, [500.0, 800.0])


@pytest.mark.parametrize("zromb", [0, 1, 2, 3, 4])
def test_romberg_weights_equal_scipy_romb(zromb):
    from cora_b200 import skysim

    w = skysim.romberg_weights(zromb)
    if zromb == 0:       # point evaluation (skysim.py:33-38)
        assert w.shape == (1,) and w[0] == 1.0
        return
    n = 2**zromb + 1
    assert w.shape == (n,) and abs(w.sum() - 1.0) < 1e-15
    rng = np.random.default_rng(zromb)
    y = rng.standard_normal(n)
    h = 0.37
    dx = 2 * h / (n - 1)
    assert abs(si.romb(y, dx=dx) / (2 * h) - w @ y) < 1e-14
    if zromb == 3:   # (4/14175) [1085, 5120, 1760, 5120, 2180, 5120, 1760, 5120, 1085] dx, normalised by 8 dx
        ref = (4.0 / 14175.0) * np.array([1085, 5120, 1760, 5120, 2180, 5120, 1760, 5120, 1085]) / 8.0
        np.testing.assert_allclose(w, ref, rtol=1e-14)


def test_sample_frequencies_layout():
    from cora_b200 import skysim

    z = np.array([800.0, 700.0, 600.0])
    za, zint = skysim._sample_frequencies(z, 3, None)
    assert zint == 9 and za.shape == (27,)
    np.testing.assert_allclose(za.reshape(3, 9)[:, 4], z)                 # centre sample
    np.testing.assert_allclose(za.reshape(3, 9)[:, 0], z - 50.0)          # half the channel spacing either side
    za0, zint0 = skysim._sample_frequencies(z, 0, None)
    assert zint0 == 1 and np.array_equal(za0, z)
    za_w, _ = skysim._sample_frequencies(z, 1, 20.0)
    np.testing.assert_allclose(za_w.reshape(3, 3), z[:, None] + np.array([-10.0, 0.0, 10.0]))


def test_alm_packing_roundtrip_and_full_half():
    from cora_b200 import hputil

    rng = np.random.default_rng(0)
    lmax = 6
    L = lmax + 1
    a = np.tril(rng.standard_normal((L, L)) + 1j * rng.standard_normal((L, L)))
    packed = hputil.pack_alm(a)
    assert packed.shape == (L * (L + 1) // 2,)
    for l in range(L):
        for m in range(l + 1):
            assert packed[m * (2 * lmax + 1 - m) // 2 + l] == a[l, m]     # hputil.py:124-152
    np.testing.assert_array_equal(hputil.unpack_alm(packed, lmax), a)
    full = hputil._make_full_alm(a)
    assert full.shape == (L, 2 * L - 1)
    for m in range(1, L):
        np.testing.assert_allclose(full[:, -m], (-1) ** m * np.conj(a[:, m]))
    np.testing.assert_allclose(hputil._make_half_alm(full), a)
    cen = hputil._make_full_alm(a, centered=True)
    np.testing.assert_allclose(cen[:, L - 1 :], a)
    np.testing.assert_allclose(hputil.unpack_alm(packed, lmax, fullm=True), full)


def test_shard_plan_bookkeeping():
    from cora_b200 import dist as cdist

    plan = cdist.ShardPlan(lmax=10, nz=7, size=3, partition="interleaved")
    assert [list(x) for x in plan.l_lists] == [[0, 3, 6, 9], [1, 4, 7, 10], [2, 5, 8]]
    assert list(plan.cb) == [3, 2, 2] and list(plan.chan_lo) == [0, 3, 5]
    assert int(plan.rows.sum()) == 11 * 12 // 2 == plan.nalm_total()
    for r in range(3):
        assert sum(plan.send_splits(r)) == plan.rows[r] * 7
        base, width = plan.nu_tables(r)
        assert base.shape == (7,) and list(width) == [3, 3, 3, 2, 2, 2, 2]
    # what rank s receives from r is what r sends to s
    for r in range(3):
        for s in range(3):
            assert plan.send_splits(r)[s] == plan.recv_splits(s)[r]
    with pytest.raises(ValueError):
        cdist.ShardPlan(4, 3, 2, partition="nope")
    assert cdist.block_partition(10, 4, 0) == (0, 3) and cdist.block_partition(10, 4, 3) == (8, 10)


def test_makesky_cli_parses(tmp_path, monkeypatch):
    """The CLI front end builds the FreqState and dispatches (the generators themselves need a GPU)."""
    from cora_b200 import makesky

    seen = {}

    def fake_21cm(fs, nside, pol, eor, oversample):
        seen.update(freq=fs.frequencies.copy(), nside=nside, pol=pol, eor=eor, oversample=oversample)
        return np.zeros((len(fs.frequencies), 12 * nside * nside))

    monkeypatch.setattr(makesky, "make_21cm", fake_21cm)
    out = tmp_path / "m.npy"
    makesky.main(["21cm", "--nside", "4", "--freq", "800", "600", "2", "--pol", "none", "--oversample", "2", str(out)])
    assert seen["nside"] == 4 and seen["pol"] == "none" and seen["oversample"] == 2 and not seen["eor"]
    np.testing.assert_allclose(seen["freq"], [800.0, 700.0])
    assert np.load(out).shape == (2, 192)



def test_fused_fill_only_for_the_package_own_spectrum():
    """clarray's fused-kernel dispatch: a subclass overriding the spectrum (or a piece of it) must be
    evaluated through the callable, like the reference's clarray always does."""
    from cora_b200 import corr21cm, galaxy, skysim

    fs = galaxy.FullSkySynchrotron.__new__(galaxy.FullSkySynchrotron)
    assert skysim._fused_fill(fs.angular_powerspectrum, fs) is not None

    class Tilted(galaxy.FullSkySynchrotron):
        def angular_ps(self, l):
            return 2.0 * super().angular_ps(l)

    t = Tilted.__new__(Tilted)
    assert skysim._fused_fill(t.angular_powerspectrum, t) is None

    class Other(galaxy.FullSkySynchrotron):
        def angular_powerspectrum(self, l, nu1, nu2):
            return 0.0 * l

    o = Other.__new__(Other)
    assert skysim._fused_fill(o.angular_powerspectrum, o) is None
    c = corr21cm.EoR21cm.__new__(corr21cm.EoR21cm)          # overrides T_b / bias only: still the fused kernel
    assert skysim._fused_fill(c.angular_powerspectrum, c) is not None

    class My21(corr21cm.Corr21cm):
        def angular_powerspectrum(self, l, nu1, nu2, redshift=False):
            return 1.0

    m = My21.__new__(My21)
    assert skysim._fused_fill(m.angular_powerspectrum, m) is None
    assert skysim._fused_fill(lambda l, a, b: 1.0, None) is None


def test_map_writer_layout_roundtrip(tmp_path):
    """The sharded map writer (makesky.py:412-450 write_map's layout: [freq, pol, pixel] + index maps): two ranks'
    blocks, unpolarised data padded to Stokes I/Q/U/V like the reference does."""
    from cora_b200 import mapio
    from cora_b200.dist import block_partition

    freq = np.linspace(800.0, 400.0, 5, endpoint=False)
    npix = 48
    rng = np.random.default_rng(0)
    full = rng.standard_normal((5, npix))
    out = str(tmp_path / "map")
    for r in (1, 0):
        lo, hi = block_partition(5, 2, r)
        mapio.write_map(out, full[lo:hi], freq, fwidth=80.0, include_pol=True, freq_start=lo, rank=r, size=2)
    got, hdr = mapio.read_map(out)
    assert got.shape == (5, 4, npix) and hdr["pol"] == ["I", "Q", "U", "V"] and hdr["axis"] == ["freq", "pol", "pixel"]
    np.testing.assert_array_equal(got[:, 0], full)
    assert not got[:, 1:].any()
    assert hdr["freq"]["centre"] == freq.tolist() and hdr["freq"]["width"] == [80.0] * 5
    assert [s["freq_start"] for s in hdr["shards"]] == [0, 3] and hdr["attrs"]["__memh5_distributed_file"] is True
    # polarised block, no padding; intensity only
    pol = rng.standard_normal((5, 4, npix))
    mapio.write_map(str(tmp_path / "p"), pol, freq)
    np.testing.assert_array_equal(mapio.read_map(str(tmp_path / "p"))[0], pol)
    mapio.write_map(str(tmp_path / "i"), full, freq, include_pol=False)
    g, h = mapio.read_map(str(tmp_path / "i"))
    assert g.shape == (5, 1, npix) and h["pol"] == ["I"] and h["freq"]["width"][0] == 80.0


def test_tabulated_correlation_host_semantics():
    """TabulatedCorrelation called on the host is numpy.interp (clamped at both ends), linear in r or in ln r."""
    from cora_b200 import corrfunc

    r = np.array([0.5, 1.0, 2.0, 4.0])
    xi = np.array([3.0, 2.0, 0.5, 0.25])
    lin = corrfunc.TabulatedCorrelation(r, xi)
    np.testing.assert_allclose(lin(np.array([0.0, 0.5, 0.75, 3.0, 9.0])), [3.0, 3.0, 2.5, 0.375, 0.25])
    log = corrfunc.TabulatedCorrelation(r, xi, kind="log")
    np.testing.assert_allclose(log(np.array([0.0, np.sqrt(2.0), 8.0])), [3.0, 1.25, 0.25])
    assert lin(np.ones((2, 3, 4))).shape == (2, 3, 4)
    with pytest.raises(ValueError):
        corrfunc.TabulatedCorrelation(r[::-1], xi)
    with pytest.raises(ValueError):
        corrfunc.TabulatedCorrelation(r, xi, kind="cubic")
    # cosine rule: the cancellation-free form equals the textbook one
    mu, x = np.array([-0.3, 0.999999]), np.array([10.0, 10.5, 300.0])
    got = corrfunc.cosine_rule(mu, x, x)
    want = np.sqrt(x[None, :, None] ** 2 + x[None, None, :] ** 2 - 2 * x[None, :, None] * x[None, None, :] * mu[:, None, None])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-6)
    assert np.all(np.diagonal(got, axis1=1, axis2=2)[1] < 0.5)
