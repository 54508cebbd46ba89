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
import numpy as np

_DEG = np.pi / 180.0
_VAR_NSIDE = 16            # the local variance is taken inside the pixels of an nside-16 map (galaxy.py:101,183)


def map_variance(input_map, nside):
    """Variance of ``input_map`` inside every pixel of the coarser ``nside`` map, RING in and out
    (``cora/foreground/galaxy.py:43-55``): in NESTED order the children of a coarse pixel are consecutive."""
    from . import healpix

    fine = np.asarray(input_map, dtype=np.float64)
    per_parent = (healpix.npix2nside(fine.shape[-1]) // int(nside)) ** 2
    grouped = healpix.reorder(fine, r2n=True).reshape(-1, per_parent)
    return healpix.reorder(grouped.var(axis=1), n2r=True)


def chunk_var(a):
    """Mean squared deviation of a (complex) array from its mean, summed in at most 30 pieces so that no
    array-sized temporary is made (``galaxy.py:58-83``)."""
    centre = a.mean()
    pieces = np.array_split(a.ravel(), min(30, a.size))
    return sum(float(np.sum(np.abs(piece - centre) ** 2)) for piece in pieces) / a.size


def _local_rms_map(hpmap):
    """The smoothed local r.m.s. of a map: 0.5 deg smoothing, r.m.s. inside nside-16 pixels, 2 deg smoothing of that
    coarse map (the recipe used both for the Haslam amplitude map, ``galaxy.py:101-104``, and for the simulated
    fluctuations, ``:182-184``).  Returned at nside 16."""
    from . import hputil

    rms = map_variance(hputil.smoothing(hpmap, sigma=0.5 * _DEG), _VAR_NSIDE) ** 0.5
    return rms


def _compress_negative(ratio):
    """x for x >= 0, tanh(x) below: fluctuations relative to the template can then never reach -1 (``galaxy.py:196-203``)."""
    return np.where(ratio < 0, np.tanh(ratio), ratio)


class ConstrainedGalaxy(maps.Sky3d):
    """Realistic simulations of the Galactic synchrotron sky (``cora/foreground/galaxy.py:86-345``): a Gaussian
    realisation with the ``FullSkySynchrotron`` spectrum whose large scales are replaced by the Haslam 408 MHz map
    extrapolated with a spectral-index map, and whose small-scale fluctuations follow the local Haslam variance.

    The reference loads ``skydata.npz`` (Haslam map, the ``gsm`` / ``md`` / ``gd`` spectral-index maps, a Faraday
    rotation map) from its data directory; that file is not part of the reference checkout, so the maps are passed in:
    ``ConstrainedGalaxy(data)`` with ``data`` a dict (or ``.npz`` path) holding ``haslam``, ``spectral_gsm`` /
    ``spectral_md`` / ``spectral_gd`` and ``faraday`` as RING maps in Galactic coordinates.  healpy's ``ud_grade``,
    ``smoothing``, ``reorder``, ``Rotator`` and ``get_interp_val`` are the ones of ``cora_b200.healpix`` /
    ``hputil.smoothing``; the Gaussian field, the constrained realisation and every harmonic transform run through the
    GPU path (``skysim.mkfullsky``, ``skysim.mkconstrained``, ``hputil.smoothing``, ``hputil.sphtrans_inv_complex``).

    Attributes: ``spectral_map`` -- which spectral-index map extrapolates the Haslam map ('md' Miville-Deschenes
    et al. 2008, the default; 'gsm'; 'gd' Giardino et al. 2002).  With 'gsm' the realisation is constrained at 408 and
    1420 MHz, otherwise at 408 MHz only."""

    spectral_map = "md"
    _dphi = 1.0            # Faraday-depth grid step and half range (rad / m^2)
    _maxphi = 500.0
    _amp_nside = 512       # resolution the amplitude map is kept at
    _constraint_freq = (408.0, 1420.0)
    _constraint_fwhm_deg = (1.0, 5.8)

    def __init__(self, data):
        from . import healpix, hputil

        table = np.load(data) if isinstance(data, str) else data
        self._haslam = np.asarray(table["haslam"], dtype=np.float64)
        self._sp_ind = {name: np.asarray(table["spectral_" + name], dtype=np.float64)
                        for name in ("gsm", "md", "gd") if ("spectral_" + name) in table}
        self._faraday = np.asarray(table["faraday"], dtype=np.float64) if "faraday" in table else None
        # amplitude of the small-scale fluctuations: smoothed local r.m.s. of the Haslam map (galaxy.py:101-104)
        self._amp_map = hputil.smoothing(healpix.ud_grade(_local_rms_map(self._haslam), self._amp_nside), sigma=2.0 * _DEG)

    # ---- unpolarised -------------------------------------------------------------------------------
    def _template(self, freqs):
        """Smooth large-scale emission: Haslam map times (nu / 408 MHz)^(spectral index per pixel) (galaxy.py:193)."""
        from . import healpix

        base = healpix.ud_grade(self._haslam, self.nside)
        index = healpix.ud_grade(self._sp_ind[self.spectral_map], self.nside)
        return base[np.newaxis, :] * (np.asarray(freqs)[:, np.newaxis] / self._constraint_freq[0]) ** index

    def getsky(self, debug=False, celestial=True):
        """A realisation of the unpolarised sky, ``float64[freq, pixel]`` (``galaxy.py:133-207``).
        ``debug``: also return the intermediate products (fluctuations, constrained part, template, amplitude map,
        mean fluctuation r.m.s.), as the reference does."""
        from . import healpix, hputil, skysim

        nconstr = len(self._constraint_freq)
        freqs = np.concatenate((np.array(self._constraint_freq), np.asarray(self.nu_pixels, dtype=np.float64)))
        lmax = 3 * self.nside - 1

        # Gaussian fluctuations at the constraint frequencies and the requested ones (point evaluation of the spectrum)
        cov = skysim.clarray(FullSkySynchrotron().angular_powerspectrum, lmax, freqs, zromb=0)
        fluct = skysim.mkfullsky(cov, self.nside)

        # their large scales -- what the data constrain -- re-expressed on every frequency through the covariance
        smooth = [hputil.smoothing(fluct[i], fwhm=w * _DEG) for i, w in enumerate(self._constraint_fwhm_deg)]
        used = range(nconstr) if self.spectral_map == "gsm" else range(1)
        large = skysim.mkconstrained(cov, [(i, smooth[i]) for i in used], self.nside)

        # small-scale part, normalised by its own mean local r.m.s. and scaled by the Haslam amplitude map
        amp = healpix.ud_grade(self._amp_map, self.nside)
        mean_rms = hputil.smoothing(_local_rms_map(fluct[0]), sigma=2.0 * _DEG).mean()
        small = (amp / mean_rms) * (fluct - large)
        if not debug:
            del fluct, large

        # relative to the data-driven template, compressed so that the sky stays positive (galaxy.py:196-203)
        template = self._template(freqs)
        small /= template
        sky = (_compress_negative(small) + 1) * template
        sky = sky[nconstr:]
        if celestial:
            sky = hputil.coord_g2c(sky)
        return (sky, fluct, large, template, amp, mean_rms) if debug else sky

    # ---- polarised ---------------------------------------------------------------------------------
    def _depth_grid(self):
        """(number of Faraday-depth cells, the depths in FFT order, their conjugate variable)."""
        ncell = 2 * int(self._maxphi / self._dphi)
        depth = np.fft.fftfreq(ncell, d=1.0 / (self._dphi * ncell))
        conj = np.fft.fftfreq(ncell, d=self._dphi)
        return ncell, depth, conj

    def _depth_to_frequency(self, depth, freqs):
        """Response of a channel of width ``median(diff(freqs))`` at ``freqs`` (MHz) to unit polarised emission at
        Faraday depth ``depth``: rotation e^{2 i phi lambda^2} averaged over the channel (galaxy.py:298-311)."""
        width = np.median(np.diff(freqs))
        angle = 2.0 * depth[:, np.newaxis] * 3e2**2 / freqs[np.newaxis, :] ** 2
        return np.exp(1.0j * angle) * np.sinc(angle * (width / freqs[np.newaxis, :]) / np.pi) / self._dphi

    def getpolsky(self, debug=False, celestial=True):
        """A realisation of the polarised sky, ``float64[freq, pol, pixel]`` (``galaxy.py:209-345``): independent
        Gaussian maps (spectrum l^-2.8) per cell of the variable conjugate to Faraday depth, correlated along depth
        (Gaussian, length 1 rad/m^2), confined in depth by the local width of the Faraday map, carried to frequency,
        saturated (|P| -> tanh |P|) and multiplied by the unpolarised realisation."""
        from . import healpix, hputil

        npix = 12 * self.nside**2
        lmax = 3 * self.nside - 1
        ncell, depth, conj = self._depth_grid()
        corr_length = 1.0

        # local width of the Faraday-depth distribution (galaxy.py:229-232)
        width_phi = healpix.ud_grade(hputil.smoothing(np.abs(self._faraday), fwhm=10.0 * _DEG), self.nside)

        # one random field per conjugate-depth cell: complex a_lm (all m) of variance C_l / 2 per component, C_l = (l/100)^-2.8
        ell = np.arange(lmax + 1, dtype=np.float64)
        ell[0] = 1.0e16                                                   # no monopole (galaxy.py:243-245)
        amplitude = (0.5 * (ell / 100.0) ** -2.8) ** 0.5
        cube = np.empty((npix, ncell), dtype=np.complex128)
        for cell in range(ncell):
            draws = np.random.standard_normal((lmax + 1, 2 * lmax + 1, 2)).view(np.complex128)[..., 0]
            cube[:, cell] = hputil.sphtrans_inv_complex(draws * amplitude[:, np.newaxis], self.nside)

        # correlate along depth, go to depth space, normalise to unit half-variance (galaxy.py:263-280)
        cube *= np.exp(-2.0 * (np.pi * corr_length * conj[np.newaxis, :]) ** 2)
        cube = np.fft.ifft(cube, axis=1)
        cube /= 2.0 * chunk_var(cube) ** 0.5

        # confine to the local Faraday-depth width; explicit normalisation (the grid is coarse where the width is small)
        window = np.exp(-0.25 * (depth[np.newaxis, :] / width_phi[:, np.newaxis]) ** 2)
        window /= window.sum(axis=1, keepdims=True)
        cube *= window

        # to frequency, saturate, scale by the total intensity
        freqs = np.asarray(self.nu_pixels, dtype=np.float64)
        response = self._depth_to_frequency(depth, freqs)
        pol = np.dot(cube, response)                                      # [pixel, freq], Q + iU as a fraction
        modulus = np.abs(pol)
        pol = pol * np.tanh(modulus) / modulus

        stokes = np.zeros((self.nu_num, 4, npix), dtype=np.float64)
        stokes[:, 0] = self.getsky(celestial=False)
        stokes[:, 1] = pol.real.T * stokes[:, 0]
        stokes[:, 2] = pol.imag.T * stokes[:, 0]
        if celestial:
            stokes = hputil.coord_g2c(stokes)
        return (cube, pol, window, width_phi, response, stokes) if debug else stokes
