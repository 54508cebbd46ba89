"""Oracle restatement of the angular power spectra C_l(nu, nu') on the hot path.

TEST INFRASTRUCTURE ONLY (see package doc).  numpy/scipy only.

* SCK foreground:  ``cora/foreground/gaussianfg.py:107-130`` with the parameter sets
  ``cora/foreground/galaxy.py:20-40`` + ``gaussianfg.py:188-192``.
* 21cm:  ``cora/signal/corr.py:891-982`` (flat-sky DCT table + bilinear lookup),
  ``cora/signal/corr21cm.py:37-208`` (T_b, growth Pade, bias, nu->z),
  ``cora/util/cosmology.py:156-210,404-430`` (H(z), comoving distance by odeint),
  ``cora/util/cubicspline.pyx:124-231,254-288`` (natural log-log cubic spline),
  ``cora/util/bilinearmap.pyx:14-59`` (clipped bilinear interpolation).
"""

import math
import os

import numpy as np
import scipy.fftpack
import scipy.integrate

# caput.astro.constants values that survive in the C_l chain.  Validated by
# the reference's own goldens (tests/test_corr.py) in tests/test_oracle_spectra.py.
C_LIGHT = 2.99792458e8
NU21 = 1420.40575177
# Cancels analytically in the comoving distance (H carries 1/Mpc, the unit carries Mpc);
# kept, with the operation order of cosmology.py:188,208,307, because odeint's adaptive
# steps react to last-bit changes of the integrand at the 1e-8 level (its rtol).
MEGA_PARSEC = 3.0856775814913673e22

PS_FILE = os.path.join(os.path.dirname(__file__), "..", "cora_b200", "data", "ps_z1.5.dat")


# ----------------------------------------------------------------------------- SCK
class SCK:
    """``A (l/l0)^-beta [(nu1/nu0)^-2a (nu2/nu0)^-2a]^1/2 exp(-ln^2(nu1/nu2)/2 zeta^2)``.

    ``gaussianfg.py:107-130``; ``l == 0`` evaluates to exactly 0 (``gaussianfg.py:108-115``).
    """

    def __init__(self, A, alpha, beta, zeta, nu_0, l_0):
        self.A, self.alpha, self.beta, self.zeta = A, alpha, beta, zeta
        self.nu_0, self.l_0 = nu_0, l_0

    def angular_ps(self, larray):
        larray = np.array(larray, dtype=np.float64)  # copy: no in-place mutation of the caller
        zero = larray == 0
        larray[zero] = 1.0
        ps = self.A * (larray / self.l_0) ** (-self.beta)
        ps[zero] = 0.0
        return ps

    def frequency_variance(self, nu):
        return (nu / self.nu_0) ** (-2 * self.alpha)

    def frequency_covariance(self, nu1, nu2):
        corr = np.exp(-0.5 * (np.log(nu1 / nu2) / self.zeta) ** 2)
        return (self.frequency_variance(nu1) * self.frequency_variance(nu2)) ** 0.5 * corr

    def angular_powerspectrum(self, l, nu1, nu2):
        return self.angular_ps(l) * self.frequency_covariance(nu1, nu2)


def full_sky_synchrotron():
    """``galaxy.py:20-27`` over ``gaussianfg.Synchrotron`` (alpha 2.80, zeta 4.0)."""
    return SCK(A=6.6e-3, alpha=2.80, beta=2.8, zeta=4.0, nu_0=408.0, l_0=100.0)


def full_sky_polarised_synchrotron():
    """``galaxy.py:30-40``."""
    return SCK(A=1.65e-3, alpha=2.80, beta=2.8, zeta=0.04, nu_0=408.0, l_0=100.0)


# ----------------------------------------------------------------------- cosmology
class Cosmology:
    """Density parameters + comoving distance in h^-1 Mpc (``cosmology.py:63-210``)."""

    def __init__(self, omega_b=0.04897, omega_c=0.26067, omega_l=0.69036, omega_g=0.0,
                 omega_n=0.0, H0=67.66, w_0=-1.0, w_a=0.0):
        self.omega_b, self.omega_c, self.omega_l = omega_b, omega_c, omega_l
        self.omega_g, self.omega_n, self.H0, self.w_0, self.w_a = omega_g, omega_n, H0, w_0, w_a

    @classmethod
    def planck2013(cls):
        """The cosmology under which ``tests/test_corr.py:18,28-29`` were generated."""
        return cls(omega_b=0.0483, omega_c=0.2589, omega_l=0.6914, H0=67.77)

    @property
    def omega_m(self):
        return self.omega_b + self.omega_c

    @property
    def omega_r(self):
        return self.omega_g + self.omega_n

    @property
    def omega_k(self):
        return 1.0 - (self.omega_l + self.omega_b + self.omega_c + self.omega_g + self.omega_n)

    def E(self, z):
        return (
            self.omega_r * (1 + z) ** 4
            + self.omega_m * (1 + z) ** 3
            + self.omega_k * (1 + z) ** 2
            + self.omega_l * (1 + z) ** (3 * (1 + self.w_0 + self.w_a))
            * np.exp(-3 * self.w_a * z / (1 + z))
        ) ** 0.5

    def comoving_distance(self, z):
        """``int_0^z c / H(z') dz'`` in Mpc/h, by ``odeint`` on the sorted redshift vector.

        ``cosmology.py:190-210`` + ``_intf_0_z`` (``:404-430``).  ``mega_parsec`` cancels
        between ``H`` (``:188``) and ``_unit_distance`` (``:307``).
        """
        scalar = not isinstance(z, np.ndarray)
        zv = np.atleast_1d(np.asarray(z, dtype=np.float64))
        order = np.argsort(zv, axis=None)
        za = np.insert(zv.ravel()[order], 0, 0.0)

        def yp(y, zz):
            return C_LIGHT / (self.H0 * self.E(zz) * 1000.0 / MEGA_PARSEC)

        out = np.zeros_like(zv)
        out.ravel()[order] = scipy.integrate.odeint(yp, 0.0, za)[1:, 0]
        out = out / (MEGA_PARSEC / (self.H0 / 100.0))
        return out[0] if scalar else out


# -------------------------------------------------------------------- cubic spline
class LogSpline:
    """Natural cubic spline in (ln k, ln P) with linear extrapolation.

    ``cubicspline.pyx:179-231`` (second derivatives by tridiagonal LU),
    ``:124-175`` (bisection + NR evaluation; end-point slope -/+ h*y2/6 outside),
    ``:254-288`` (log wrapper).
    """

    def __init__(self, data):
        d = np.log(np.asarray(data, dtype=np.float64))
        self.x, self.y = d[:, 0].copy(), d[:, 1].copy()
        n = len(self.x)
        x, y = self.x, self.y
        m = n - 2
        f = (y[2:] - y[1:-1]) / (x[2:] - x[1:-1]) - (y[1:-1] - y[:-2]) / (x[1:-1] - x[:-2])
        al = (x[2:] - x[:-2]) / 3
        bt = np.zeros(m)
        gm = np.zeros(m)
        bt[1:] = (x[2:-1] - x[1:-2]) / 6
        gm[:-1] = (x[2:-1] - x[1:-2]) / 6
        lo = np.zeros(m)
        mu = np.zeros(m)
        zz = np.zeros(m)
        lo[0] = al[0]
        mu[0] = gm[0] / al[0]
        for i in range(1, m):
            lo[i] = al[i] - bt[i] * mu[i - 1]
            mu[i] = gm[i] / lo[i]
        zz[0] = f[0] / lo[0]
        for i in range(1, m):
            zz[i] = (f[i] - bt[i] * zz[i - 1]) / lo[i]
        for i in range(m - 2, -1, -1):
            zz[i] = zz[i] - mu[i] * zz[i + 1]
        self.y2 = np.zeros(n)
        self.y2[1:-1] = zz

    @classmethod
    def fromfile(cls, fname):
        return cls(np.loadtxt(fname, usecols=[0, 1]))

    def log_value(self, lx):
        """Spline value (in log space) at log-abscissa ``lx``."""
        x, y, y2 = self.x, self.y, self.y2
        lx = np.asarray(lx, dtype=np.float64)
        n = len(x)
        kl = np.clip(np.searchsorted(x, lx, side="right") - 1, 0, n - 2)
        kh = kl + 1
        h = x[kh] - x[kl]
        a = (x[kh] - lx) / h
        b = (lx - x[kl]) / h
        c = (a**3 - a) * h**2 / 6
        d = (b**3 - b) * h**2 / 6
        v = a * y[kl] + b * y[kh] + c * y2[kl] + d * y2[kh]
        # below the first knot
        h0 = x[1] - x[0]
        lo_v = ((y[1] - y[0]) / h0 - h0 * y2[1] / 6) * (lx - x[0]) + y[0]
        # at or above the last knot
        h1 = x[n - 1] - x[n - 2]
        hi_v = ((y[n - 1] - y[n - 2]) / h1 + h1 * y2[n - 2] / 6) * (lx - x[n - 1]) + y[n - 1]
        v = np.where(lx < x[0], lo_v, v)
        return np.where(lx >= x[n - 1], hi_v, v)

    def __call__(self, k):
        return np.exp(self.log_value(np.log(k)))


# ------------------------------------------------------------------------ bilinear
def bilinear_interp(arr, x, y):
    """Clipped bilinear interpolation (``bilinearmap.pyx:14-59``).

    Clip x to [0, nx - 1e-5], y to [0, ny - 1e-5], truncate to unsigned int, 4-point
    blend.  (If the clip hits the top edge the reference reads one element past the row
    / table; the hot path never gets there: x < 499 and y < 32767 for all supported
    inputs, which the host code asserts.)
    """
    nx, ny = arr.shape
    xx = np.clip(x, 0.0, nx - 1e-5)
    yy = np.clip(y, 0.0, ny - 1e-5)
    x0 = xx.astype(np.uint32).astype(np.int64)
    y0 = yy.astype(np.uint32).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    wa = (x1 - xx) * (y1 - yy)
    wb = (x1 - xx) * (yy - y0)
    wc = (xx - x0) * (y1 - yy)
    wd = (xx - x0) * (yy - y0)
    return wa * arr[x0, y0] + wb * arr[x0, y1] + wc * arr[x1, y0] + wd * arr[x1, y1]


# ---------------------------------------------------------------------------- 21cm
KPERP_MIN, KPERP_MAX, NKPERP = 1e-4, 40.0, 500
KPAR_MAX, NKPAR = 20.0, 32768


class Corr21cm:
    """21cm brightness-temperature C_l(nu, nu') in the flat-sky DCT-table approximation."""

    kstar = 5.0
    ps_redshift = 1.5
    bias = 1.0
    omega_HI = 6.2e-4

    def __init__(self, cosmology=None, ps_file=PS_FILE):
        self.cosmology = cosmology if cosmology is not None else Cosmology()
        self.spline = LogSpline.fromfile(ps_file)
        self._tables = None

    # corr21cm.py:25-29
    def ps_vv(self, k):
        return np.exp(-0.5 * k**2 / self.kstar**2) * self.spline(k)

    # corr21cm.py:37-62, 87
    def T_b(self, z):
        c = self.cosmology
        return (
            3.9e-4
            * ((c.omega_m + c.omega_l * (1 + z) ** -3) / 0.29) ** -0.5
            * ((1.0 + z) / 2.5) ** 0.5
            * (self.omega_HI / 1e-3)
        )

    # corr21cm.py:109-138
    def growth_factor(self, z):
        x = ((1.0 / self.cosmology.omega_m) - 1.0) / (1.0 + z) ** 3
        num = 1.0 + 1.175 * x + 0.3064 * x**2 + 0.005355 * x**3
        den = 1.0 + 1.857 * x + 1.021 * x**2 + 0.1530 * x**3
        return (1.0 + x) ** 0.5 / (1.0 + z) * num / den

    # corr21cm.py:140-175
    def growth_rate(self, z):
        x = ((1.0 / self.cosmology.omega_m) - 1.0) / (1.0 + z) ** 3
        dnum = 3.0 * x * (1.175 + 0.6127 * x + 0.01607 * x**2)
        dden = 3.0 * x * (1.857 + 2.042 * x + 0.4590 * x**2)
        num = 1.0 + 1.175 * x + 0.3064 * x**2 + 0.005355 * x**3
        den = 1.0 + 1.857 * x + 1.021 * x**2 + 0.1530 * x**3
        return 1.0 + 1.5 * x / (1.0 + x) + dnum / num - dden / den

    def tables(self):
        """The three DCT-I tables dd, dv, vv of shape (500, 32768) (``corr.py:915-942``)."""
        if self._tables is None:
            kperp = np.logspace(np.log10(KPERP_MIN), np.log10(KPERP_MAX), NKPERP)[:, None]
            kpar = np.linspace(0, KPAR_MAX, NKPAR)[None, :]
            k = (kpar**2 + kperp**2) ** 0.5
            mu2 = kpar**2 / k**2
            dd = self.ps_vv(k)  # _freq_window = 0 -> sinc^2 factor is 1 (corr.py:889,928-932)
            dv = dd * mu2
            vv = dd * mu2**2
            norm = KPAR_MAX / (2 * NKPAR)
            self._tables = tuple(scipy.fftpack.dct(t, type=1) * norm for t in (dd, dv, vv))
        return self._tables

    def sample_vectors(self, z):
        """Per-redshift vectors (chi, b, f, pf, D) entering ``corr.py:944-951``."""
        z = np.asarray(z, dtype=np.float64)
        chi = self.cosmology.comoving_distance(z)
        b = np.ones_like(z) * self.bias
        f = self.growth_rate(z)
        pf = self.T_b(z)
        D = self.growth_factor(z) / self.growth_factor(self.ps_redshift)
        return chi, b, f, pf, D

    def angular_powerspectrum_z(self, la, za1, za2):
        """``corr.py:944-982``; arguments broadcast like the reference's."""
        dd, dv, vv = self.tables()
        la = np.asarray(la, dtype=np.float64)
        za1 = np.asarray(za1, dtype=np.float64)
        za2 = np.asarray(za2, dtype=np.float64)
        xa1, b1, f1, pf1, D1 = self.sample_vectors(za1)
        xa2, b2, f2, pf2, D2 = self.sample_vectors(za2)
        xc = 0.5 * (xa1 + xa2)
        rpar = np.abs(xa2 - xa1)
        la = np.where(la == 0.0, 1e-10, la)
        x = (np.log10(la) - np.log10(xc * KPERP_MIN)) / np.log10(KPERP_MAX / KPERP_MIN) * (NKPERP - 1)
        y = rpar / (math.pi / KPAR_MAX)
        x, y = np.broadcast_arrays(x, y)
        psdd = bilinear_interp(dd, x, y)
        psdv = bilinear_interp(dv, x, y)
        psvv = bilinear_interp(vv, x, y)
        return (D1 * D2 * pf1 * pf2 / (xc**2 * np.pi)) * (
            (b1 * b2) * psdd + (f1 * b2 + f2 * b1) * psdv + (f1 * f2) * psvv
        )

    def angular_powerspectrum(self, l, nu1, nu2, redshift=False):
        """``corr21cm.py:183-208``: frequencies in MHz unless ``redshift``."""
        if not redshift:
            nu1 = NU21 / np.asarray(nu1, dtype=np.float64) - 1.0
            nu2 = NU21 / np.asarray(nu2, dtype=np.float64) - 1.0
        return self.angular_powerspectrum_z(l, nu1, nu2)
