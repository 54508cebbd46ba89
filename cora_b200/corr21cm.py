"""21cm brightness-temperature angular power spectrum (flat-sky DCT-table approximation).

Mirrors ``cora/signal/corr21cm.py:9-239`` (``Corr21cm``: T_b, growth Pade forms, bias,
frequency -> redshift) on top of ``RedshiftCorrelation.angular_powerspectrum_fft``
(``cora/signal/corr.py:891-982``).  The one-off 500 x 32768 P(k_perp, k_par) tables and their
DCT-I are built on the GPU and stay device-resident (the analogue of the reference's
``_aps_cache``); the table lookup + Romberg channel average is one fused kernel.
"""

import os

import numpy as np

from . import _dev, _lib, maps
from .cosmology import NU21, Cosmology

_PS_FILE = os.path.join(os.path.dirname(__file__), "data", "ps_z1.5.dat")

NKPERP, NKPAR = 500, 32768


def _natural_spline_y2(x, y):
    """Second derivatives of the natural cubic spline through (x, y): Thomas solve of the
    tridiagonal system (what ``cubicspline.pyx:179-231`` sets up)."""
    n = len(x)
    h = np.diff(x)
    rhs = np.diff(y)[1:] / h[1:] - np.diff(y)[:-1] / h[:-1]
    diag = (x[2:] - x[:-2]) / 3.0
    off = h[1:-1] / 6.0
    m = n - 2
    cp = np.zeros(m)
    dp = np.zeros(m)
    cp[0] = (off[0] / diag[0]) if m > 1 else 0.0
    dp[0] = rhs[0] / diag[0]
    for i in range(1, m):
        den = diag[i] - off[i - 1] * cp[i - 1]
        if i < m - 1:
            cp[i] = off[i] / den
        dp[i] = (rhs[i] - off[i - 1] * dp[i - 1]) / den
    sol = np.zeros(m)
    sol[-1] = dp[-1]
    for i in range(m - 2, -1, -1):
        sol[i] = dp[i] - cp[i] * sol[i + 1]
    y2 = np.zeros(n)
    y2[1:-1] = sol
    return y2


class Corr21cm(maps.Sky3d):
    r"""Correlation function of HI brightness temperature fluctuations (``corr21cm.py:9-35``).

    Default power spectrum: ``ps_z1.5.dat`` (k, P(k) at z = 1.5) through a natural cubic spline
    in log-log with a Gaussian cut ``exp(-k^2 / 2 k*^2)``, ``k* = 5``.
    """

    add_mean = False
    _kstar = 5.0
    ps_redshift = 1.5
    _bias = 1.0

    def __init__(self, ps_file=None, redshift=1.5, cosmology=None, **kwargs):
        data = np.loadtxt(ps_file or _PS_FILE, usecols=[0, 1])
        if np.any(data <= 0):
            raise ValueError("Data must be non-negative.")
        self._lnk = np.ascontiguousarray(np.log(data[:, 0]))
        self._lnp = np.ascontiguousarray(np.log(data[:, 1]))
        self._y2 = _natural_spline_y2(self._lnk, self._lnp)
        self.ps_redshift = redshift
        self.cosmology = cosmology if cosmology is not None else Cosmology()
        self._tab = None

    # ---- scalar model functions (host, closed form) --------------------------------
    def omega_HI(self, z):
        return 6.2e-4

    def T_b(self, z):
        """Mean 21cm brightness temperature in K (``corr21cm.py:37-62``)."""
        c = self.cosmology
        return (3.9e-4 * ((c.omega_m + c.omega_l * (1 + z) ** -3) / 0.29) ** -0.5 * ((1.0 + z) / 2.5) ** 0.5
                * (self.omega_HI(z) / 1e-3))

    def mean(self, z):
        return self.T_b(z) if self.add_mean else np.zeros_like(z)

    def prefactor(self, z):
        return self.T_b(z)

    def _pade(self, z):
        x = ((1.0 / self.cosmology.omega_m) - 1.0) / (1.0 + z) ** 3
        num = 1.0 + 1.175 * x + 0.3064 * x**2 + 0.005355 * x**3
        den = 1.0 + 1.857 * x + 1.021 * x**2 + 0.1530 * x**3
        return x, num, den

    def growth_factor(self, z):
        """Pade approximation of the growth factor (``corr21cm.py:109-138``, arXiv:1012.2671)."""
        x, num, den = self._pade(z)
        return (1.0 + x) ** 0.5 / (1.0 + z) * num / den

    def growth_rate(self, z):
        """Growth rate from the derivative of the Pade form (``corr21cm.py:140-175``)."""
        x, num, den = self._pade(z)
        dnum = 3.0 * x * (1.175 + 0.6127 * x + 0.01607 * x**2)
        dden = 3.0 * x * (1.857 + 2.042 * x + 0.4590 * x**2)
        return 1.0 + 1.5 * x / (1.0 + x) + dnum / num - dden / den

    def bias_z(self, z):
        return np.ones_like(z) * self._bias

    def mean_nu(self, freq):
        return self.mean(NU21 / freq - 1.0)

    # ---- device table ----------------------------------------------------------------
    def table(self):
        """Device-resident DCT tables, built once (``corr.py:915-942``), layout
        planar ``tab[{dd, dv, vv}][y][x]``."""
        if self._tab is None:
            t = _dev.torch()
            lib = _lib.load()
            tab = _dev.empty((lib.cora_b200_ps_table_21cm_bytes() // 8,), t.float64)
            nbytes = lib.cora_b200_ps_table_21cm_workspace_bytes()
            ws = _dev.workspace(nbytes)
            _lib.call("cora_b200_ps_table_21cm", _lib.ptr(self._lnk), _lib.ptr(self._lnp), _lib.ptr(self._y2),
                      len(self._lnk), float(self._kstar), _lib.ptr(tab), _lib.ptr(ws), int(nbytes), _lib.stream_ptr())
            t.cuda.current_stream().synchronize()
            self._tab = tab
        return self._tab

    def table_entries(self, xs, ys):
        """Entries ``[dd, dv, vv]`` at reference-style indices ``[x][y]`` (for tests / cache export)."""
        t = _dev.torch()
        xs = _dev.to_device(np.asarray(xs, dtype=np.int32), t.int32)
        ys = _dev.to_device(np.asarray(ys, dtype=np.int32), t.int32)
        out = _dev.empty((xs.numel(), 3), t.float64)
        _lib.call("cora_b200_ps_table_21cm_gather", _lib.ptr(self.table()), _lib.ptr(xs), _lib.ptr(ys), int(xs.numel()),
                  _lib.ptr(out), _lib.stream_ptr())
        return out.cpu().numpy()

    def _sample_vectors(self, z):
        """Rows chi, b, f, pf, D for redshifts ``z`` (``corr.py:944-951``), host side."""
        z = np.asarray(z, dtype=np.float64)
        chi = self.cosmology.comoving_distance(z)
        return np.stack([chi, self.bias_z(z), self.growth_rate(z), self.prefactor(z),
                         self.growth_factor(z) / self.growth_factor(self.ps_redshift)])

    # ---- spectra -----------------------------------------------------------------------
    def angular_powerspectrum(self, l, nu1, nu2, redshift=False):
        """C_l(nu1, nu2); frequencies in MHz unless ``redshift`` (``corr21cm.py:183-208``)."""
        t = _dev.torch()
        if not redshift:
            z1 = NU21 / np.asarray(nu1, dtype=np.float64) - 1.0
            z2 = NU21 / np.asarray(nu2, dtype=np.float64) - 1.0
        else:
            z1, z2 = np.asarray(nu1, dtype=np.float64), np.asarray(nu2, dtype=np.float64)
        # per-argument vectors are computed on the un-broadcast arrays (one ODE solve each, like the
        # reference) and then broadcast
        la = np.asarray(l, dtype=np.float64)
        shape = np.broadcast_shapes(la.shape, z1.shape, z2.shape)
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape)

        def vectors(z):
            v = self._sample_vectors(z.ravel()).reshape((5,) + (1,) * (len(shape) - z.ndim) + z.shape)
            return np.ascontiguousarray(np.broadcast_to(v, (5,) + tuple(shape))).reshape(5, n)

        lb = np.ascontiguousarray(np.broadcast_to(la, shape)).ravel()
        v1b, v2b = vectors(z1), vectors(z2)
        out = _dev.empty((n,), t.float64)
        dl, d1, d2 = _dev.to_device(lb, t.float64), _dev.to_device(v1b, t.float64), _dev.to_device(v2b, t.float64)
        _lib.call("cora_b200_aps_21cm_points", _lib.ptr(self.table()), _lib.ptr(dl), _lib.ptr(d1), _lib.ptr(d2), n,
                  _lib.ptr(out), _lib.stream_ptr())
        res = out.cpu().numpy().reshape(shape)
        return res if res.ndim else float(res)

    # fused clarray path (skysim.clarray looks this up); samples are frequencies (skysim.py:41-49)
    def _b200_fill_inputs(self, nu_samples, w):
        """Device-resident inputs of the fill kernel: the DCT table (built once), the per-sample
        vectors chi, b, f, pf, D (host cosmology, ``corr.py:944-951``) and the Romberg weights."""
        t = _dev.torch()
        z = NU21 / np.asarray(nu_samples, dtype=np.float64) - 1.0
        hv = self._sample_vectors(z)
        vec = _dev.to_device(hv, t.float64)  # [5, nz*zint]
        # kernel choice: the row-weight kernel reads every distinct table row of a channel pair's y window once; its
        # window is ~2 x (comoving width of a channel) x 20/pi rows tall.  Tall windows (wide channels) favour the
        # per-sample-pair kernel (measured: 256 channels over 400-800 MHz, ~100-140 rows: 4.3 vs 5.0 ms; 1024 channels,
        # ~30 rows: 104 vs 44 ms).
        zint = len(w)
        chi = hv[0].reshape(-1, zint)
        rows = 2.0 * float(np.max(np.abs(chi[:, -1] - chi[:, 0]))) * (20.0 / np.pi) if zint > 1 else 0.0
        variant = 1 if rows > 80.0 else 0
        forced = os.environ.get("CORA_B200_FILL_VARIANT")          # "0" / "1": A/B runs and tests
        if forced in ("0", "1"):
            variant = int(forced)
        return self.table(), vec, _dev.to_device(w, t.float64), variant

    # skysim.clarray takes the fused kernel only while these methods are the ones defined here (T_b, bias_z,
    # growth_* and the cosmology enter through _sample_vectors and may be overridden freely, e.g. EoR21cm)
    _b200_fill_methods = ("angular_powerspectrum", "_b200_fill", "_b200_fill_inputs", "_sample_vectors")
    _b200_fill_lower = True      # the fill kernel writes lower triangles; skysim.clarray mirrors them

    def _b200_fill(self, inputs, l0, l_step, nl, nz, zint, out, stream=None, lower_only=False):
        """``lower_only``: write the entries ``(i, j <= i)`` only -- all the root stage reads; the upper triangle of
        ``out`` is left as it was."""
        tab, vec, wd, variant = inputs
        _lib.call("cora_b200_cl_fill_21cm", _lib.ptr(tab), _lib.ptr(vec[0]), _lib.ptr(vec[1]), _lib.ptr(vec[2]),
                  _lib.ptr(vec[3]), _lib.ptr(vec[4]), _lib.ptr(wd), int(l0), int(l_step), int(nl), int(nz), int(zint),
                  _lib.ptr(out), int(bool(lower_only)), int(variant), _lib.stream_ptr(stream))


    def _b200_fill_tiles(self, inputs, nl, nz, zint, tile0, ntiles, out_ptrs, l_owner, l_row, stream=None, tile_step=1):
        """Fill sharded over channel-pair tiles (tile0, tile0 + tile_step, ...: ntiles of them) with the rows scattered
        to the GPUs owning each l (multi-GPU path; lower triangle only)."""
        tab, vec, wd, variant = inputs
        _lib.call("cora_b200_cl_fill_21cm_tiles", _lib.ptr(tab), _lib.ptr(vec[0]), _lib.ptr(vec[1]), _lib.ptr(vec[2]),
                  _lib.ptr(vec[3]), _lib.ptr(vec[4]), _lib.ptr(wd), int(nl), int(nz), int(zint), int(tile0), int(ntiles), int(tile_step), int(variant),
                  _lib.ptr(out_ptrs), _lib.ptr(l_owner), _lib.ptr(l_row), _lib.stream_ptr(stream))


Corr21cm._b200_fill_origin = Corr21cm


class EoR21cm(Corr21cm):
    """Parameters more suitable for the reionisation epoch (``corr21cm.py:333-385``)."""

    _bias = 3.0

    def T_b(self, z):
        """Eq. (4) of Santos, Ferramacho & Silva 2009, in K (``corr21cm.py:334-361``)."""
        c = self.cosmology
        h = c.H0 / 100.0
        return 23e-3 * (c.omega_b * h**2 / 0.02) * (0.15 / (c.omega_m * h**2) * ((1.0 + z) / 10)) ** 0.5 * (h / 0.7) ** -1

    def omega_HI(self, z):
        return 5e-3

    def x_h(self, z):
        return 0.25
