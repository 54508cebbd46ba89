"""Host-side cosmology for the 21cm spectrum (mirrors the part of ``cora/util/cosmology.py``
the C_l chain uses: density parameters ``:63-96``, ``H(z)`` ``:156-188``, comoving distance by
``odeint`` ``:190-210,404-430``).  O(nfreq * 9) evaluations per run: stays on the host and
feeds the C_l kernel its per-sample vectors."""

from dataclasses import asdict, dataclass

import numpy as np
from scipy import integrate as si

# caput.astro.constants equivalents; mega_parsec cancels in the comoving distance
C_LIGHT = 2.99792458e8
NU21 = 1420.40575177
MEGA_PARSEC = 3.0856775814913673e22


@dataclass
class Cosmology(object):
    """Planck-2018 defaults, as the reference (``cosmology.py:63-80``)."""

    units: str = "cosmo"
    omega_b: float = 0.04897
    omega_c: float = 0.26067
    omega_l: float = 0.69036
    omega_g: float = 0.0
    omega_n: float = 0.0
    H0: float = 67.66
    w_0: float = -1.0
    w_a: float = 0.0

    @property
    def omega_m(self):
        return self.omega_b + self.omega_c

    @property
    def omega_r(self):
        return self.omega_g + self.omega_n

    @property
    def omega_k(self):
        return 1.0 - (self.omega_l + self.omega_b + self.omega_c + self.omega_g + self.omega_n)

    def to_dict(self):
        return asdict(self)

    def H(self, z=0.0):
        """Hubble parameter in SI units (s^-1)."""
        H = self.H0 * (
            self.omega_r * (1 + z) ** 4
            + self.omega_m * (1 + z) ** 3
            + self.omega_k * (1 + z) ** 2
            + self.omega_l * (1 + z) ** (3 * (1 + self.w_0 + self.w_a)) * np.exp(-3 * self.w_a * z / (1 + z))
        ) ** 0.5
        return H * 1000.0 / MEGA_PARSEC

    @property
    def _unit_distance(self):
        if self.units == "astro":
            return MEGA_PARSEC
        if self.units == "cosmo":
            return MEGA_PARSEC / (self.H0 / 100.0)
        if self.units == "si":
            return 1.0
        raise RuntimeError("Units not known")

    def comoving_distance(self, z):
        """Comoving distance to redshift(s) z, one ODE solve over the sorted redshifts."""
        return _integrate_from_zero(lambda z1: C_LIGHT / self.H(z1), z) / self._unit_distance


def _integrate_from_zero(f, z):
    if not isinstance(z, np.ndarray):
        return _integrate_from_zero(f, np.array([z], dtype=np.float64))[0]
    order = np.argsort(z, axis=None)
    grid = np.insert(z.ravel()[order], 0, 0)
    out = np.zeros_like(z)
    out.ravel()[order] = si.odeint(lambda y, zz: f(zz), 0.0, grid)[1:, 0]
    return out
