"""Galactic synchrotron: the full-sky spectra used by ``cora-makesky gaussianfg`` (parameter sets of
``cora/foreground/galaxy.py:20-40``) and the constrained realisation ``ConstrainedGalaxy`` (``:43-345``)."""

from . import gaussianfg, maps


class FullSkySynchrotron(gaussianfg.Synchrotron):
    """Amplitudes matched to La Porta et al. 2008 for |b| > 5 deg."""

    A = 6.6e-3
    beta = 2.8
    nu_0 = 408.0
    l_0 = 100.0


class FullSkyPolarisedSynchrotron(gaussianfg.Synchrotron):
    """Polarised synchrotron: polarisation fraction 0.5, short frequency correlation length."""

    A = 1.65e-3
    beta = 2.8
    nu_0 = 408.0
    l_0 = 100.0
    zeta = 0.04


# ------------------------------------------------------------------ constrained Galactic synchrotron
def map_variance(input_map, nside):
    """Variance of ``input_map`` inside every pixel of the coarser ``nside`` map, RING in and out
    (``cora/foreground/galaxy.py:43-55``: group the NESTED children of each coarse pixel)."""
    import numpy as np

    from . import healpix

    input_map = np.asarray(input_map, dtype=np.float64)
    inp_nside = healpix.npix2nside(input_map.shape[-1])
    map_nest = healpix.reorder(input_map, r2n=True).reshape(-1, (inp_nside // nside) ** 2)
    return healpix.reorder(map_nest.var(axis=1), n2r=True)


def chunk_var(a):
    """Variance of a (complex) array about its mean, accumulated in chunks (``galaxy.py:58-83``)."""
    import numpy as np

    nchunks = min(30, a.size)
    mean = a.mean()
    t = 0.0
    for sec in np.array_split(a.ravel(), nchunks):
        t += np.sum(np.abs(sec - mean) ** 2)
    return t / a.size


class ConstrainedGalaxy(maps.Sky3d):
    """Realistic simulations of the Galactic synchrotron sky (``cora/foreground/galaxy.py:86-345``): a Gaussian
    realisation with the ``FullSkySynchrotron`` spectrum, constrained on large scales to the Haslam 408 MHz map
    extrapolated with a spectral-index map, with fluctuations scaled by the local Haslam variance.

    The reference loads ``skydata.npz`` (Haslam map, the ``gsm`` / ``md`` / ``gd`` spectral-index maps, a Faraday
    rotation map) from its data directory; that file is not part of the reference checkout, so the maps are passed in:
    ``ConstrainedGalaxy(data)`` with ``data`` a dict (or ``.npz`` path) holding ``haslam``, ``spectral_gsm`` /
    ``spectral_md`` / ``spectral_gd`` and ``faraday`` as RING maps in Galactic coordinates.  healpy's ``ud_grade``,
    ``smoothing``, ``reorder``, ``Rotator`` and ``get_interp_val`` are the ones of ``cora_b200.healpix`` /
    ``hputil.smoothing``; the Gaussian field, the constrained realisation and every harmonic transform run through the
    GPU path (``skysim.mkfullsky``, ``skysim.mkconstrained``)."""

    spectral_map = "md"
    _dphi = 1.0
    _maxphi = 500.0
    _amp_nside = 512

    def __init__(self, data):
        import numpy as np

        from . import healpix, hputil

        self._load_data(data)
        vm = map_variance(hputil.smoothing(self._haslam, sigma=np.radians(0.5)), 16)
        self._amp_map = hputil.smoothing(healpix.ud_grade(vm**0.5, self._amp_nside), sigma=np.radians(2.0))

    def _load_data(self, data):
        import numpy as np

        f = np.load(data) if isinstance(data, str) else data
        self._haslam = np.asarray(f["haslam"], dtype=np.float64)
        self._sp_ind = {k: np.asarray(f["spectral_" + k], dtype=np.float64) for k in ("gsm", "md", "gd") if ("spectral_" + k) in f}
        self._faraday = np.asarray(f["faraday"], dtype=np.float64) if "faraday" in f else None

    def getsky(self, debug=False, celestial=True):
        """A realisation of the unpolarised sky, ``float64[freq, pixel]`` (``galaxy.py:133-207``)."""
        import numpy as np

        from . import healpix, hputil, skysim

        haslam = healpix.ud_grade(self._haslam, self.nside)
        syn = FullSkySynchrotron()
        lmax = 3 * self.nside - 1
        efreq = np.concatenate((np.array([408.0, 1420.0]), np.asarray(self.nu_pixels, dtype=np.float64)))

        # map of random fluctuations, including the two constraint frequencies
        cla = skysim.clarray(syn.angular_powerspectrum, lmax, efreq, zromb=0)
        fg = skysim.mkfullsky(cla, self.nside)

        # the smoothed fluctuations on each scale, and a multifrequency map constrained to look like them
        sub408 = hputil.smoothing(fg[0], fwhm=np.radians(1.0))
        sub1420 = hputil.smoothing(fg[1], fwhm=np.radians(5.8))
        if self.spectral_map == "gsm":
            fgs = skysim.mkconstrained(cla, [(0, sub408), (1, sub1420)], self.nside)
        else:
            fgs = skysim.mkconstrained(cla, [(0, sub408)], self.nside)

        sc = healpix.ud_grade(self._sp_ind[self.spectral_map], self.nside)
        am = healpix.ud_grade(self._amp_map, self.nside)

        # bump up the variance of the fluctuations according to the variance map
        vm = hputil.smoothing(fg[0], sigma=np.radians(0.5))
        vm = hputil.smoothing(map_variance(vm, 16) ** 0.5, sigma=np.radians(2.0))
        mv = vm.mean()

        fgt = (am / mv) * (fg - fgs)
        if not debug:
            del fg, fgs

        # the smooth, large scale emission from Haslam + spectral map
        fgsmooth = haslam[np.newaxis, :] * ((efreq / 408.0)[:, np.newaxis] ** sc)

        # rescale so that the output is always positive
        fgt /= fgsmooth
        fgt = np.where(fgt < 0, np.tanh(fgt), fgt)
        fgt += 1
        fgt *= fgsmooth
        fgt = fgt[2:]

        if celestial:
            fgt = hputil.coord_g2c(fgt)
        if debug:
            return fgt, fg, fgs, fgsmooth, am, mv
        return fgt

    def getpolsky(self, debug=False, celestial=True):
        """A realisation of the polarised sky, ``float64[freq, pol, pixel]`` (``galaxy.py:209-345``): random maps in
        the Fourier conjugate of Faraday depth, weighted by the local Faraday-depth width, transformed to frequency and
        scaled by the unpolarised realisation."""
        import numpy as np

        from . import healpix, hputil

        sigma_phi = healpix.ud_grade(hputil.smoothing(np.abs(self._faraday), fwhm=np.radians(10.0)), self.nside)
        xiphi = 1.0
        lmax = 3 * self.nside - 1
        la = np.arange(lmax + 1)

        def angular(l):
            l = np.array(l, dtype=np.float64)
            l[np.where(l == 0)] = 1.0e16
            return (l / 100.0) ** -2.8

        dphi, maxphi = self._dphi, self._maxphi
        nphi = 2 * int(maxphi / dphi)
        phifreq = np.fft.fftfreq(nphi, d=(1.0 / (dphi * nphi)))
        ps_weight = (angular(la[:, np.newaxis]) / 2.0) ** 0.5

        # random maps in the Fourier conjugate of phi
        map2 = np.zeros((12 * self.nside**2, nphi), dtype=np.complex128)
        for i in range(nphi):
            w = np.random.standard_normal((lmax + 1, 2 * lmax + 1, 2)).view(np.complex128)[..., 0]
            w *= ps_weight
            map2[:, i] = hputil.sphtrans_inv_complex(w, self.nside)

        # weight the conj-phi direction to give the phi correlation structure, and transform back into phi
        pcfreq = np.fft.fftfreq(nphi, d=dphi)
        map2 *= np.exp(-2 * (np.pi * xiphi * pcfreq[np.newaxis, :]) ** 2)
        map2 = np.fft.ifft(map2, axis=1)
        map2 /= 2.0 * chunk_var(map2) ** 0.5

        w = np.exp(-0.25 * (phifreq[np.newaxis, :] / sigma_phi[:, np.newaxis]) ** 2)
        w /= w.sum(axis=1)[:, np.newaxis]
        map2 *= w
        if not debug:
            del w

        def ptrans(phi, freq, dfreq):
            dx = dfreq / freq
            alpha = 2.0 * phi * 3e2**2 / freq**2
            return np.exp(1.0j * alpha) * np.sinc(alpha * dx / np.pi)

        fa = np.asarray(self.nu_pixels, dtype=np.float64)
        df = np.median(np.diff(fa))
        pta = ptrans(phifreq[:, np.newaxis], fa[np.newaxis, :], df) / dphi
        map4 = np.dot(map2, pta)
        if not debug:
            del map2

        map4a = np.abs(map4)
        map4 = map4 * np.tanh(map4a) / map4a
        del map4a

        map5 = np.zeros((self.nu_num, 4, 12 * self.nside**2), dtype=np.float64)
        map5[:, 0] = self.getsky(celestial=False)
        map5[:, 1] = map4.real.T
        map5[:, 2] = map4.imag.T
        map5[:, 1:3] *= map5[:, 0, np.newaxis, :]
        if not debug:
            del map4
        if celestial:
            map5 = hputil.coord_g2c(map5)
        if debug:
            return map2, map4, w, sigma_phi, pta, map5
        return map5
