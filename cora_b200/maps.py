"""Sky-map base classes: frequency axis + the callers of the hot path.

Mirrors what the path needs of ``cora/core/maps.py``: the ``frequencies`` / ``nu_pixels`` axis
(``:140-170``) and ``Sky3d.getsky / getpolsky / getalms`` (``:203-252``).
"""

import numpy as np

from . import skysim


class Map3d(object):
    """Frequency axis of a 3-D map (``maps.py:95-170``; angular-patch geometry is flat-sky only
    and out of scope)."""

    nu_lower = 500.0
    nu_upper = 900.0
    _nu_num = 128
    _frequencies = None

    @property
    def nu_num(self):
        return len(self.frequencies)

    @nu_num.setter
    def nu_num(self, num):
        self._nu_num = num

    @property
    def frequencies(self):
        """List of frequencies in the map (channel centres, MHz)."""
        if self._frequencies is not None:
            return self._frequencies
        return self.nu_lower + (np.arange(self._nu_num) + 0.5) * ((self.nu_upper - self.nu_lower) / self._nu_num)

    @frequencies.setter
    def frequencies(self, freq):
        self._frequencies = freq

    nu_pixels = frequencies


class Sky3d(Map3d):
    """Base class for maps of the spherical sky at multiple frequencies (``maps.py:203-252``).

    Attributes
    ----------
    oversample : int
        Romberg order for the integration over each finite-width channel.
    """

    oversample = 3
    nside = 128

    def angular_powerspectrum(self, l, nu1, nu2):
        raise Exception("Not implemented in base class.")

    def mean_nu(self, freq):
        return np.zeros_like(freq)

    def getsky(self):
        """Create a map of the unpolarised sky, ``float64[nfreq, npix]``."""
        lmax = 3 * self.nside - 1
        cla = skysim.clarray(self.angular_powerspectrum, lmax, self.nu_pixels, zromb=self.oversample, device_out=True,
                             _lower_only=True)      # mkfullsky's root reads the lower triangle only (LAPACK-style)
        mean = np.asarray(self.mean_nu(np.asarray(self.nu_pixels, dtype=np.float64)), dtype=np.float64)
        if not mean.any():
            return skysim.mkfullsky(cla, self.nside)
        from . import _dev

        sky = skysim.mkfullsky(cla, self.nside, device_out=True)
        sky += _dev.to_device(mean, sky.dtype)[:, None]  # the reference's host-side broadcast add (maps.py:235)
        return _dev.to_host(sky)

    def getpolsky(self):
        """Stokes I, Q, U, V maps ``[nfreq, 4, npix]`` (Q = U = V = 0 for an unpolarised model)."""
        sky_I = self.getsky()
        sky_IQU = np.zeros((sky_I.shape[0], 4, sky_I.shape[1]), dtype=sky_I.dtype)
        sky_IQU[:, 0] = sky_I
        return sky_IQU

    def getalms(self, lmax):
        cla = skysim.clarray(self.angular_powerspectrum, lmax, self.nu_pixels, device_out=True)
        return skysim.mkfullsky(cla, self.nside, alms=True)
