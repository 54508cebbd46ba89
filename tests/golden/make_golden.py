#!/usr/bin/env python
"""Generate the committed golden fixtures by running the REAL reference.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

What it does (SURVEY 8c recipe):
  1. copies /root/reference/cora + setup.py to a scratch dir under /tmp, strips ``-flto``
     (the stock link step dies on it) and builds the Cython extensions with gcc;
  2. puts tiny shims for the two un-installable dependencies on sys.path:
     ``caput.astro.constants`` (c, nu21, ...), single-rank ``caput.mpiarray``, and an
     import-only ``healpy`` stub (the SHT is NOT exercised: healpy is absent, parity at
     that boundary is unpinned);
  3. imports the unmodified reference and records inputs/outputs of the hot-path
     functions as small ``.npz`` files next to this script.

Nothing from the reference is copied into the repository; only numbers it computed.
"""

import os
import shutil
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SCRATCH = "/tmp/cora_ref_build"


def build_reference():
    if not os.path.isdir(os.path.join(SCRATCH, "cora")):
        os.makedirs(SCRATCH, exist_ok=True)
        shutil.copytree(os.path.join(REF, "cora"), os.path.join(SCRATCH, "cora"))
        shutil.copy(os.path.join(REF, "setup.py"), SCRATCH)
        sp = os.path.join(SCRATCH, "setup.py")
        os.chmod(sp, 0o644)
        txt = open(sp).read().replace('["-flto", *FAST_MATH_ARGS]', "[*FAST_MATH_ARGS]")
        open(sp, "w").write(txt)
    import glob

    if not glob.glob(os.path.join(SCRATCH, "cora", "util", "bilinearmap*.so")):
        env = dict(os.environ, CC="gcc")
        subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=SCRATCH, env=env)


def install_shims():
    caput = types.ModuleType("caput")
    astro = types.ModuleType("caput.astro")
    const = types.ModuleType("caput.astro.constants")
    const.c = 2.99792458e8
    const.nu21 = 1420.40575177
    const.mega_parsec = 3.0856775814913673e22
    const.mega_year = 3.15576e13
    const.degree = np.pi / 180.0
    const.G_n = 6.6743e-11
    const.a_rad = 7.565723e-16
    astro.constants = const
    caput.astro = astro

    mpiarray = types.ModuleType("caput.mpiarray")

    class MPIArray(np.ndarray):
        """single-rank stand-in: enumerate / local_array / allgather / redistribute / wrap."""

        @property
        def local_array(self):
            return self.view(np.ndarray)

        @property
        def global_shape(self):
            return self.shape

        def enumerate(self, axis):
            return [(i, i) for i in range(self.shape[axis])]

        def allgather(self):
            return self.view(np.ndarray)

        def redistribute(self, axis):
            return self

        @classmethod
        def wrap(cls, arr, axis):
            return arr.view(cls)

        @property
        def local_offset(self):
            return (0,) * self.ndim

        def reshape(self, *shape):        # caput allows None = "keep this (distributed) axis"
            if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
                shape = tuple(shape[0])
            shape = tuple(self.shape[i] if v is None else v for i, v in enumerate(shape))
            return np.ndarray.reshape(self, shape)

    def zeros(shape, dtype=np.float64, axis=0):
        return np.zeros(shape, dtype=dtype).view(MPIArray)

    mpiarray.MPIArray = MPIArray
    mpiarray.zeros = zeros
    caput.mpiarray = mpiarray

    healpy = types.ModuleType("healpy")  # import-only stub; alm2map is never called here

    # corrfunc.py imports hankl / hankel / pyfftlog at module level (not used by corr_to_clarray) and takes
    # cosine_rule from caput.astro.coordinates: import-only stubs + the oracle's restatement of cosine_rule
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import corrfunc as ocf

    coords = types.ModuleType("caput.astro.coordinates")
    sph = types.ModuleType("caput.astro.coordinates.spherical")
    sph.cosine_rule = ocf.cosine_rule
    coords.spherical = sph
    astro.coordinates = coords
    for name in ("hankl", "hankel", "pyfftlog"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules.update({"caput.astro.coordinates": coords, "caput.astro.coordinates.spherical": sph})

    sys.modules.update(
        {"caput": caput, "caput.astro": astro, "caput.astro.constants": const,
         "caput.mpiarray": mpiarray, "healpy": healpy}
    )
    sys.path.insert(0, SCRATCH)


def main():
    build_reference()
    install_shims()

    from cora.core import skysim
    from cora.foreground import galaxy
    from cora.signal import corr21cm
    from cora.util import cosmology, nputil, hputil

    # ------------------------------------------------------------------ C_l, foreground
    fsyn = galaxy.FullSkySynchrotron()
    fpol = galaxy.FullSkyPolarisedSynchrotron()
    nside, nfreq = 8, 6
    freq = np.linspace(800.0, 400.0, nfreq, endpoint=False)
    lmax_fg = 3 * nside
    cl_fg = skysim.clarray(fsyn.angular_powerspectrum, lmax_fg, freq)
    cl_fgpol = skysim.clarray(fpol.angular_powerspectrum, lmax_fg, freq)
    cl_fg_z0 = skysim.clarray(fsyn.angular_powerspectrum, lmax_fg, freq, zromb=0)
    np.savez(os.path.join(HERE, "cl_sck.npz"), freq=freq, lmax=lmax_fg, cl=cl_fg, cl_pol=cl_fgpol,
             cl_zromb0=cl_fg_z0)

    # ------------------------------------------------------------------ C_l, 21cm
    out21 = {}
    for tag, cosmo in (("p18", cosmology.Cosmology()),
                       ("p13", cosmology.Cosmology(omega_b=0.0483, omega_c=0.2589, omega_l=0.6914, H0=67.77))):
        cr = corr21cm.Corr21cm()
        cr.cosmology = cosmo
        aps1 = cr.angular_powerspectrum(np.arange(1000), 800.0, 800.0)
        fa = np.linspace(400.0, 800.0, 64)
        aps2 = cr.angular_powerspectrum(np.arange(1000)[:, None, None], fa[None, :, None], fa[None, None, :])
        out21[tag + "_aps1"] = aps1
        out21[tag + "_aps2_l400_40_40"] = aps2[400, 40, 40]
        out21[tag + "_aps2_l200_10_40"] = aps2[200, 10, 40]
        out21[tag + "_aps2_sub"] = aps2[::37, ::7, ::5].copy()
        lmax21 = 3 * nside - 1
        out21[tag + "_cl"] = skysim.clarray(cr.angular_powerspectrum, lmax21, freq, zromb=3)
        out21[tag + "_cl_romb1"] = skysim.clarray(cr.angular_powerspectrum, lmax21, freq, zromb=1)
        if tag == "p18":
            # table samples + sampled vectors to pin the table-build and host vectors
            rows = np.array([0, 1, 7, 100, 250, 333, 498, 499])
            cols = np.array([0, 1, 2, 3, 17, 100, 1000, 4097, 16384, 30000, 32766, 32767])
            out21["tab_rows"], out21["tab_cols"] = rows, cols
            out21["tab_dd"] = cr._aps_dd[np.ix_(rows, cols)]
            out21["tab_dv"] = cr._aps_dv[np.ix_(rows, cols)]
            out21["tab_vv"] = cr._aps_vv[np.ix_(rows, cols)]
            z = 1420.40575177 / np.linspace(400.0, 800.0, 11) - 1.0
            out21["vec_z"] = z
            out21["vec_chi"] = cosmo.comoving_distance(z)
            out21["vec_f"] = cr.growth_rate(z)
            out21["vec_D"] = cr.growth_factor(z) / cr.growth_factor(cr.ps_redshift)
            out21["vec_pf"] = cr.prefactor(z)
            kk = np.array([1e-5, 1e-4, 3.3e-3, 0.1, 1.0, 11.0, 11.9, 44.7])
            out21["ps_k"] = kk
            out21["ps_vv"] = cr.ps_vv(kk)
    np.savez(os.path.join(HERE, "cl_21cm.npz"), freq=freq, lmax=3 * nside - 1, **out21)

    # ------------------------------------------------------------------ mkfullsky(alms=True)
    rec = {"roots": [], "gauss": []}
    _root, _draw = nputil.matrix_root_manynull, nputil.complex_std_normal

    def root_spy(mat, **kw):
        r = _root(mat, **kw)
        rec["roots"].append(np.array(r))
        return r

    def draw_spy(shape, rng=None):
        g = _draw(shape, rng=rng)
        rec["gauss"].append(np.array(g))
        return g

    nputil.matrix_root_manynull, nputil.complex_std_normal = root_spy, draw_spy
    try:
        alm_fg = skysim.mkfullsky(cl_fg, nside, alms=True, rng=np.random.default_rng(0))
        roots_fg, gauss_fg = rec["roots"], rec["gauss"]
        rec["roots"], rec["gauss"] = [], []
        alm_21 = skysim.mkfullsky(out21["p18_cl"], nside, alms=True, rng=np.random.default_rng(0))
        roots_21, gauss_21 = rec["roots"], rec["gauss"]
    finally:
        nputil.matrix_root_manynull, nputil.complex_std_normal = _root, _draw

    def pad_gauss(gl, L):
        out = np.zeros((L, gl[0].shape[0], L), dtype=np.complex128)
        for l, g in enumerate(gl):
            out[l, :, : l + 1] = g
        return out

    np.savez(os.path.join(HERE, "mkfullsky_sck.npz"), cl=cl_fg, nside=nside, alm=np.asarray(alm_fg),
             roots=np.array(roots_fg), gauss=pad_gauss(gauss_fg, lmax_fg + 1))
    np.savez(os.path.join(HERE, "mkfullsky_21cm.npz"), cl=out21["p18_cl"], nside=nside, alm=np.asarray(alm_21),
             roots=np.array(roots_21), gauss=pad_gauss(gauss_21, 3 * nside))

    # polarised block matrix of makesky.gaussianfg (makesky.py:368-387), 3 channels
    nf = 3
    fq = freq[:nf]
    lm = 12
    cv = np.zeros((lm + 1, 4, nf, 4, nf))
    cv[:, 0, :, 0, :] = skysim.clarray(fsyn.angular_powerspectrum, lm, fq)
    cv[:, 1, :, 1, :] = skysim.clarray(fpol.angular_powerspectrum, lm, fq)
    cv[:, 2, :, 2, :] = cv[:, 1, :, 1, :]
    cv = cv.reshape(lm + 1, 4 * nf, 4 * nf)
    alm_pol = skysim.mkfullsky(cv, 4, alms=True, rng=np.random.default_rng(3))
    np.savez(os.path.join(HERE, "mkfullsky_pol.npz"), cl=cv, alm=np.asarray(alm_pol), nfreq=nf)

    # ------------------------------------------------------------------ root corner cases
    rng = np.random.default_rng(5)
    a = rng.standard_normal((12, 4))
    lowrank = a @ a.T  # rank 4 -> Cholesky fails -> eigh branch
    r_low = nputil.matrix_root_manynull(lowrank.copy(), truncate=False)
    r_low_t, npos = nputil.matrix_root_manynull(lowrank.copy())
    spd = lowrank + 12 * np.eye(12)
    r_spd = nputil.matrix_root_manynull(spd.copy(), truncate=False)
    g = nputil.complex_std_normal((5, 7), rng=np.random.default_rng(11))
    np.savez(os.path.join(HERE, "root_cases.npz"), lowrank=lowrank, root_lowrank=r_low, root_lowrank_trunc=r_low_t,
             num_pos=npos, spd=spd, root_spd=r_spd, cstd_seed11=g)

    # ------------------------------------------------------------------ alm packing
    L = 6
    sq = np.tril(rng.standard_normal((L, L)) + 1j * rng.standard_normal((L, L)))
    packed = hputil.pack_alm(sq)
    np.savez(os.path.join(HERE, "alm_pack.npz"), square=sq, packed=packed, unpacked=hputil.unpack_alm(packed, L - 1))

    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("  %-24s %8d B" % (f, os.path.getsize(os.path.join(HERE, f))))


def extra_c2rows():
    """Config 2's channel axis (256 channels, lmax 767) at 8 l rows through the reference's own ``clarray`` and
    ``Corr21cm.angular_powerspectrum``: the spectrum is wrapped so that l index i evaluates l = rows[i] (clarray's
    per-l results do not depend on its l-sectioning).  Stored: every 16th channel row (+ the last)."""
    build_reference()
    install_shims()
    from cora.core import skysim
    from cora.signal import corr21cm

    rows = np.array([0, 1, 2, 5, 100, 383, 766, 767])
    cr = corr21cm.Corr21cm()
    freq = np.linspace(800.0, 400.0, 256, endpoint=False)

    def aps(l, z1, z2):
        return cr.angular_powerspectrum(rows[np.asarray(l)].astype(np.float64), z1, z2)

    cl = skysim.clarray(aps, len(rows) - 1, freq)
    chan_rows = np.unique(np.append(np.arange(0, 256, 16), 255))
    np.savez(os.path.join(HERE, "cl_21cm_c2rows.npz"), rows=rows, freq=freq, chan_rows=chan_rows,
             cl_rows=cl[:, chan_rows, :])
    print("cl_21cm_c2rows.npz written")


def extra_root_large():
    """The eigen branch where it is live: the 1024-channel foreground covariance.
    Through the reference's own ``clarray`` (l rows by the wrapper trick), ``mkfullsky``'s jitter and
    ``nputil.matrix_root_manynull``: the unpolarised T matrix at four l, and the polarised
    blockdiag(T, E, B, V) matrix ((4 x 1024)^2) at l = 5.  Stored: num_pos, the retained eigenvalues
    (squared column norms of the root), the retained root columns (T) and M M^T on a sampled index set."""
    build_reference()
    install_shims()
    from cora.core import skysim
    from cora.foreground import galaxy
    from cora.util import nputil

    rows = np.array([1, 5, 100, 700, 1535, 1536])
    nz = 1024
    freq = np.linspace(800.0, 400.0, nz, endpoint=False)

    def rows_cl(model):
        def aps(l, z1, z2):
            return model.angular_powerspectrum(rows[np.asarray(l)].astype(np.float64), z1, z2)
        return skysim.clarray(aps, len(rows) - 1, freq)

    clT = rows_cl(galaxy.FullSkySynchrotron())
    clP = rows_cl(galaxy.FullSkyPolarisedSynchrotron())
    out = {"rows": rows, "freq": freq}
    sel = np.unique(np.linspace(0, nz - 1, 96).astype(int))
    out["sel"] = sel
    for i in (1, 2, 4):           # l = 5, 100, 1535
        cm = clT[i] + np.identity(nz) * np.max(np.diag(clT[i])) * 1.0e-14          # skysim.py:116-117
        root = nputil.matrix_root_manynull(cm.copy(), truncate=False)
        _, npos = nputil.matrix_root_manynull(cm.copy(), truncate=True)
        lam = (root**2).sum(axis=0)
        tag = "T%d_" % rows[i]
        out[tag + "num_pos"] = npos
        out[tag + "evals"] = lam[lam > 0]
        out[tag + "cols"] = root[:, lam > 0]
        out[tag + "mmt_sel"] = (root @ root.T)[np.ix_(sel, sel)]
        out[tag + "cl_sel"] = clT[i][np.ix_(sel, sel)]
    # polarised block matrix at l = 5 (makesky.py:368-382)
    i = 1
    cv = np.zeros((4, nz, 4, nz))
    cv[0, :, 0, :] = clT[i]
    cv[1, :, 1, :] = clP[i]
    cv[2, :, 2, :] = clP[i]
    cv = cv.reshape(4 * nz, 4 * nz)
    cm = cv + np.identity(4 * nz) * np.max(np.diag(cv)) * 1.0e-14
    root = nputil.matrix_root_manynull(cm, truncate=False)
    lam = (root**2).sum(axis=0)
    keep = lam > 0
    out["pol5_num_pos"] = int(keep.sum())
    out["pol5_evals"] = lam[keep]
    blk = np.array([[np.abs(root[b * nz:(b + 1) * nz, c]).max() > 0 for b in range(4)] for c in np.where(keep)[0]])
    out["pol5_cols_per_block"] = blk.sum(axis=0)
    psel = np.concatenate([b * nz + sel[::3] for b in range(4)])
    out["pol5_sel"] = psel
    out["pol5_mmt_sel"] = (root[psel] @ root[psel].T)
    out["pol5_v_rows_max"] = np.abs(root[3 * nz:]).max()
    np.savez(os.path.join(HERE, "root_large.npz"), **out)
    print("root_large.npz written:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def corr_test_function(r):
    """A smooth correlation function with a finite r -> 0 limit (stands in for the LSS xi(r) interpolators)."""
    return np.exp(-((r / 60.0) ** 2)) / (1.0 + (r / 15.0) ** 2) + 0.05 * np.cos(r / 35.0) * np.exp(-r / 400.0)


def extra_corr_to_clarray():
    """``cora.signal.corrfunc.corr_to_clarray`` and ``legendre_array`` (the unmodified reference functions) on a
    small radial grid, with and without the radial bin quadrature."""
    build_reference()
    install_shims()
    import scipy.special as ss

    if not hasattr(ss, "lpn"):
        # the reference calls scipy.special.lpn (corrfunc.py:285), removed from this image's scipy (1.18):
        # legendre_p_all is its replacement (same recurrence, same values)
        ss.lpn = lambda n, z: (np.asarray(ss.legendre_p_all(n, z))[0], None)
    from cora.signal import corrfunc

    xarray = np.linspace(2900.0, 3400.0, 6)
    lmax = 40
    out = {"xarray": xarray, "lmax": lmax}
    out["cl_romb2_q2"] = np.asarray(corrfunc.corr_to_clarray(corr_test_function, lmax, xarray, xromb=2, q=2, chunksize=16))
    out["cl_romb0_q3"] = np.asarray(corrfunc.corr_to_clarray(corr_test_function, lmax, xarray, xromb=0, q=3, chunksize=50))
    out["cl_romb1_w40"] = np.asarray(corrfunc.corr_to_clarray(corr_test_function, lmax, xarray, xromb=1, xwidth=40.0, q=2, chunksize=20))
    mu = np.array([-0.99, -0.5, 0.0, 0.3, 0.77, 0.999])
    out["leg_mu"] = mu
    out["leg"] = corrfunc.legendre_array(60, mu)
    np.savez(os.path.join(HERE, "corr_to_clarray.npz"), **out)
    print("corr_to_clarray.npz written:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        main()
    elif what == "c2rows":
        extra_c2rows()
    elif what == "root_large":
        extra_root_large()
    elif what == "corr_to_clarray":
        extra_corr_to_clarray()
    else:
        raise SystemExit("usage: make_golden.py [all|c2rows|root_large|corr_to_clarray]")
