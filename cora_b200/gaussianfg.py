"""Santos-Cooray-Knox style Gaussian foregrounds: covariance separable in angle and frequency.

Mirrors the spectrum half of ``cora/foreground/gaussianfg.py`` (``ForegroundMap`` ``:20-41``,
``ForegroundSCK`` ``:87-130``, parameter sets ``:188-213``); the flat-sky field generator
(``getfield``, ``angular_correlation``) is out of scope.
"""

import numpy as np

from . import _dev, _lib, maps


class ForegroundMap(maps.Sky3d):
    r"""Foregrounds with :math:`C_l(\nu,\nu') = A_l B(\nu, \nu')`."""

    def angular_ps(self, l):
        pass

    def frequency_covariance(self, nu1, nu2):
        pass

    def angular_powerspectrum(self, l, nu1, nu2):
        return self.angular_ps(l) * self.frequency_covariance(nu1, nu2)


class ForegroundSCK(ForegroundMap):
    r"""SCK power law: :math:`A (l/l_0)^{-\beta} (\nu_1\nu_2/\nu_0^2)^{-\alpha}
    \exp(-\ln^2(\nu_1/\nu_2) / 2\zeta^2)`.  Set ``A``, ``alpha``, ``beta``, ``zeta``."""

    nu_0 = 130.0
    l_0 = 1000.0

    def _params(self):
        return (float(self.A), float(self.beta), float(self.l_0), float(self.alpha), float(self.nu_0), float(self.zeta))

    def angular_powerspectrum(self, l, nu1, nu2):
        """C_l(nu1, nu2) for broadcastable arguments; ``l == 0`` gives exactly 0
        (``gaussianfg.py:107-130``).  Unlike the reference the ``l`` array is not mutated."""
        t = _dev.torch()
        la, n1, n2 = np.broadcast_arrays(np.asarray(l, dtype=np.float64), np.asarray(nu1, dtype=np.float64),
                                         np.asarray(nu2, dtype=np.float64))
        shape = la.shape
        n = la.size
        if n == 0:
            return np.zeros(shape)
        dl, d1, d2 = (_dev.to_device(np.ascontiguousarray(a).ravel(), t.float64) for a in (la, n1, n2))
        out = _dev.empty((n,), t.float64)
        _lib.call("cora_b200_aps_sck_points", *self._params(), _lib.ptr(dl), _lib.ptr(d1), _lib.ptr(d2), n, _lib.ptr(out),
                  _lib.stream_ptr())
        res = out.cpu().numpy().reshape(shape)
        return res if res.ndim else float(res)

    def angular_ps(self, larray):
        return self.angular_powerspectrum(larray, self.nu_0, self.nu_0)

    def frequency_covariance(self, nu1, nu2):
        # A_l at l = l_0 is A: divide it out
        return self.angular_powerspectrum(self.l_0, nu1, nu2) / self.A

    # fused clarray path (skysim.clarray looks this up)
    def _b200_fill_inputs(self, nu_samples, w):
        """Device-resident inputs of the fill kernel (sample frequencies, Romberg weights)."""
        t = _dev.torch()
        return _dev.to_device(nu_samples, t.float64), _dev.to_device(w, t.float64)

    # skysim.clarray takes the fused kernel only while these methods are the ones defined here
    _b200_fill_methods = ("angular_powerspectrum", "angular_ps", "frequency_covariance", "_b200_fill", "_params")

    def _b200_fill(self, inputs, l0, l_step, nl, nz, zint, out, stream=None, lower_only=False):
        ns, wd = inputs   # (lower_only is a 21cm-kernel option; the closed-form fill always writes full matrices)
        _lib.call("cora_b200_cl_fill_sck", *self._params(), _lib.ptr(ns), _lib.ptr(wd), int(l0), int(l_step), int(nl),
                  int(nz), int(zint), _lib.ptr(out), _lib.stream_ptr(stream))


ForegroundSCK._b200_fill_origin = ForegroundSCK


class Synchrotron(ForegroundSCK):
    A = 7.00e-4
    alpha = 2.80
    beta = 2.4
    zeta = 4.0


class ExtraGalacticFreeFree(ForegroundSCK):
    A = 1.40e-8
    alpha = 2.10
    beta = 1.0
    zeta = 35.0


class GalacticFreeFree(ForegroundSCK):
    A = 8.80e-8
    alpha = 2.15
    beta = 3.0
    zeta = 35.0


class PointSources(ForegroundSCK):
    A = 5.70e-5
    alpha = 2.07
    beta = 1.1
    zeta = 1.0
