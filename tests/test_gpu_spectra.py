"""GPU parity: C_l fill kernels (csrc/cl.cu) vs the reference's known-answer values, the
reference-generated fixtures and the oracle."""

import numpy as np
import pytest

from conftest import golden
from oracle import skysim as osk
from oracle import spectra as osp

pytestmark = pytest.mark.gpu


def _normwise(a, ref):
    d = np.sqrt(np.abs(np.einsum("lii->li", ref)))
    scale = d[:, :, None] * d[:, None, :] + 1e-300
    return np.max(np.abs(a - ref) / scale)


def test_sck_reference_goldens():
    # /root/reference/tests/test_corr.py:34-57, through the CUDA point-wise kernel
    from cora_b200 import galaxy

    cr = galaxy.FullSkySynchrotron()
    aps1 = cr.angular_powerspectrum(np.arange(1000), 800.0, 800.0)
    assert len(aps1) == 1000
    assert np.allclose(aps1.sum(), 75.47681191093129, rtol=1e-7)
    fa = np.linspace(400.0, 800.0, 64)
    aps2 = cr.angular_powerspectrum(np.arange(1000)[:, None, None], fa[None, :, None], fa[None, None, :])
    assert aps2.shape == (1000, 64, 64)
    assert np.allclose(aps2[400, 40, 40], 9.690708728692975e-06, rtol=1e-7)
    assert np.allclose(aps2[200, 10, 40], 0.00017630767166797886, rtol=1e-7)
    assert aps1[0] == 0.0


def test_sck_clarray_vs_fixture_and_oracle():
    from cora_b200 import galaxy, skysim

    g = golden("cl_sck.npz")
    freq, lmax = g["freq"], int(g["lmax"])
    cl = skysim.clarray(galaxy.FullSkySynchrotron().angular_powerspectrum, lmax, freq)
    np.testing.assert_allclose(cl, g["cl"], rtol=1e-12, atol=0)
    clp = skysim.clarray(galaxy.FullSkyPolarisedSynchrotron().angular_powerspectrum, lmax, freq)
    np.testing.assert_allclose(clp, g["cl_pol"], rtol=1e-12, atol=0)
    cl0 = skysim.clarray(galaxy.FullSkySynchrotron().angular_powerspectrum, lmax, freq, zromb=0)
    np.testing.assert_allclose(cl0, g["cl_zromb0"], rtol=1e-12, atol=0)
    # config 1 shape against the oracle: nside 64 -> lmax 192, 32 channels
    f32 = np.linspace(800.0, 400.0, 32, endpoint=False)
    a = skysim.clarray(galaxy.FullSkySynchrotron().angular_powerspectrum, 192, f32)
    b = osk.clarray(osp.full_sky_synchrotron().angular_powerspectrum, 192, f32)
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=0)
    # zwidth override and another Romberg order
    a = skysim.clarray(galaxy.FullSkySynchrotron().angular_powerspectrum, 20, f32[:5], zromb=2, zwidth=3.0)
    b = osk.clarray(osp.full_sky_synchrotron().angular_powerspectrum, 20, f32[:5], zromb=2, zwidth=3.0)
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=0)


def test_clarray_generic_callable_and_errors():
    from cora_b200 import skysim

    def aps(l, z1, z2):
        return (1.0 + l) ** -2.0 * np.exp(-0.5 * (z1 - z2) ** 2) * (1 + 0.1 * z1 * z2)

    z = np.linspace(1.0, 2.0, 7)
    a = skysim.clarray(aps, 23, z, zromb=3)
    b = osk.clarray(aps, 23, z, zromb=3)
    np.testing.assert_allclose(a, b, rtol=1e-13)
    a0 = skysim.clarray(aps, 23, z, zromb=0)
    np.testing.assert_allclose(a0, osk.clarray(aps, 23, z, zromb=0), rtol=1e-15)
    with pytest.raises(ValueError):
        skysim.clarray(aps, 4, z)  # lmax < 5 breaks the reference's array_split too


@pytest.fixture(scope="module")
def gpu_corr21cm():
    from cora_b200 import corr21cm

    return corr21cm.Corr21cm()


def test_21cm_table_vs_fixture(gpu_corr21cm):
    g = golden("cl_21cm.npz")
    rows, cols = g["tab_rows"], g["tab_cols"]
    xs, ys = np.meshgrid(rows, cols, indexing="ij")
    ent = gpu_corr21cm.table_entries(xs.ravel(), ys.ravel()).reshape(len(rows), len(cols), 3)
    # row scale = the y = 0 entry of dd (largest of the row)
    scale = np.abs(gpu_corr21cm.table_entries(rows, np.zeros_like(rows))[:, 0])[:, None]
    for k, name in enumerate(("tab_dd", "tab_dv", "tab_vv")):
        assert np.max(np.abs(ent[:, :, k] - g[name]) / scale) < 1e-13, name


def test_21cm_reference_goldens_planck2013():
    # /root/reference/tests/test_corr.py:7-31 (goldens were generated under Planck 2013)
    from cora_b200 import corr21cm
    from cora_b200.cosmology import Cosmology

    cr = corr21cm.Corr21cm(cosmology=Cosmology(omega_b=0.0483, omega_c=0.2589, omega_l=0.6914, H0=67.77))
    aps1 = cr.angular_powerspectrum(np.arange(1000), 800.0, 800.0)
    assert len(aps1) == 1000
    assert np.allclose(aps1.sum(), 1.5963772205823096e-09, rtol=1e-7)
    fa = np.linspace(400.0, 800.0, 64)
    aps2 = cr.angular_powerspectrum(np.arange(1000)[:, None, None], fa[None, :, None], fa[None, None, :])
    assert aps2.shape == (1000, 64, 64)
    assert np.allclose(aps2[400, 40, 40], 8.986790805379046e-13, rtol=1e-7)
    assert np.allclose(aps2[200, 10, 40], 1.1939298801340165e-18, rtol=1e-7)
    g = golden("cl_21cm.npz")
    np.testing.assert_allclose(aps1, g["p13_aps1"], rtol=1e-11)
    d = np.sqrt(np.abs(np.einsum("lii->li", aps2)))
    scale = (d[:, :, None] * d[:, None, :])[::37, ::7, ::5]
    assert np.max(np.abs(aps2[::37, ::7, ::5] - g["p13_aps2_sub"]) / scale) < 1e-11


def test_21cm_clarray_vs_fixture(gpu_corr21cm):
    from cora_b200 import skysim

    g = golden("cl_21cm.npz")
    freq, lmax = g["freq"], int(g["lmax"])
    np.testing.assert_allclose(gpu_corr21cm.angular_powerspectrum(np.arange(1000), 800.0, 800.0), g["p18_aps1"], rtol=1e-11)
    for key, zromb in (("p18_cl", 3), ("p18_cl_romb1", 1)):
        cl = skysim.clarray(gpu_corr21cm.angular_powerspectrum, lmax, freq, zromb=zromb)
        assert _normwise(cl, g[key]) < 1e-11


def test_21cm_clarray_vs_oracle_midsize(gpu_corr21cm, oracle_corr21cm):
    """nside 32 -> lmax 95, 24 channels of the 400-800 MHz band, and a narrow-channel case."""
    from cora_b200 import skysim

    for freq in (np.linspace(800.0, 400.0, 24, endpoint=False), np.linspace(700.0, 690.0, 16, endpoint=False)):
        a = skysim.clarray(gpu_corr21cm.angular_powerspectrum, 95, freq)
        b = osk.clarray(oracle_corr21cm.angular_powerspectrum, 95, freq)
        assert _normwise(a, b) < 1e-11
        assert np.array_equal(a, np.transpose(a, (0, 2, 1)))  # exactly symmetric by construction


C2_ROWS = np.array([0, 1, 2, 5, 100, 383, 766, 767])


def _rows_aps(aps, rows):
    """Wrap a spectrum so that ``clarray(..., lmax = len(rows) - 1, ...)`` evaluates the l's in ``rows``
    (per-l results do not depend on the l-sectioning)."""
    rows = np.asarray(rows)

    def wrapped(l, z1, z2):
        return aps(rows[np.asarray(l)].astype(np.float64) if np.ndim(l) else float(rows[int(l)]), z1, z2)

    return wrapped


def test_21cm_clarray_config2_axis_vs_oracle_rows(gpu_corr21cm, oracle_corr21cm):
    """Config 2's exact channel axis (256 channels 800 -> 400 MHz, lmax 767): the fused fill kernel against the
    oracle's clarray on 8 l rows spread over the range."""
    from cora_b200 import skysim

    freq = np.linspace(800.0, 400.0, 256, endpoint=False)
    a = skysim.clarray(gpu_corr21cm.angular_powerspectrum, 767, freq)[C2_ROWS]
    b = osk.clarray(_rows_aps(oracle_corr21cm.angular_powerspectrum, C2_ROWS), len(C2_ROWS) - 1, freq)
    assert _normwise(a, b) < 1e-11
    # the committed reference-generated rows (tests/golden/make_golden.py c2rows): every 16th channel row
    g = golden("cl_21cm_c2rows.npz")
    assert np.array_equal(g["rows"], C2_ROWS)
    d = np.sqrt(np.abs(np.einsum("lii->li", a)))
    scale = d[:, g["chan_rows"], None] * d[:, None, :] + 1e-300
    assert np.max(np.abs(a[:, g["chan_rows"], :] - g["cl_rows"]) / scale) < 1e-11


def test_21cm_fill_row_weight_kernel_vs_per_sample_kernel(gpu_corr21cm):
    """The row-weight fill kernel (one x-interpolation per l over row-weighted band sums + exact boundary
    corrections) against the per-sample-pair kernel of round 1 (CORA_B200_FILL_V1=1 in a subprocess, same device
    table): agreement to a few ulp of the row scale on config-2-like and narrow-channel axes, with interleaved l."""
    import os
    import subprocess
    import sys
    import tempfile

    from cora_b200 import skysim

    cases = [(np.linspace(800.0, 400.0, 64, endpoint=False), 383), (np.linspace(650.0, 600.0, 48, endpoint=False), 1535),
             (np.linspace(800.0, 400.0, 16, endpoint=False), 47)]
    code = """
import sys, numpy as np
sys.path.insert(0, %r)
from cora_b200 import corr21cm, skysim
c = corr21cm.Corr21cm()
cases = [(np.linspace(800.0, 400.0, 64, endpoint=False), 383), (np.linspace(650.0, 600.0, 48, endpoint=False), 1535),
         (np.linspace(800.0, 400.0, 16, endpoint=False), 47)]
np.savez(sys.argv[1], *[skysim.clarray(c.angular_powerspectrum, lmax, f) for f, lmax in cases])
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "v1.npz")
        subprocess.run([sys.executable, "-c", code, path], check=True, env=dict(os.environ, CORA_B200_FILL_VARIANT="1"))
        old = np.load(path)
        os.environ["CORA_B200_FILL_VARIANT"] = "0"      # the host would pick per geometry: force the row-weight kernel
        try:
            for k, (freq, lmax) in enumerate(cases):
                new = skysim.clarray(gpu_corr21cm.angular_powerspectrum, lmax, freq)
                assert _normwise(new, old["arr_%d" % k]) < 2e-14
                assert np.array_equal(new, np.transpose(new, (0, 2, 1)))
        finally:
            del os.environ["CORA_B200_FILL_VARIANT"]
