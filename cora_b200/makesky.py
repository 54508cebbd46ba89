"""Drivers of record for the path: ``cora-makesky 21cm`` and ``cora-makesky gaussianfg``.

Mirrors ``cora/scripts/makesky.py``: ``FreqState`` (``:44-92``), the ``21cm`` command
(``:313-344``) and the ``gaussianfg`` command (``:347-390``).  The click CLI and the HDF5 writer
(``write_map``, ``:412-450``; h5py is not available here) are outside the hot path; the functions
below return the ``(nfreq, npol, npix)`` arrays the commands would write.

    python -m cora_b200.makesky 21cm --nside 64 --freq 800 400 32 --pol none out.npy
"""

import numpy as np


def _axis_centre(start, stop, num):
    """``num`` channel centres from ``start`` towards ``stop``, the last one a channel short of ``stop``."""
    return np.linspace(start, stop, num, endpoint=False), abs(stop - start) / num


def _axis_centre_nyquist(start, stop, num):
    """Centres including both ends (the Nyquist channel is kept)."""
    return np.linspace(start, stop, num, endpoint=True), abs((stop - start) / (num - 1))


def _axis_edge(start, stop, num):
    """``start`` / ``stop`` are the band edges; the width keeps the sign of the axis direction."""
    width = (stop - start) / num
    return start + width * (np.arange(num) + 0.5), width


_AXIS_MODES = {"centre": _axis_centre, "centre_nyquist": _axis_centre_nyquist}


class FreqState(object):
    """The frequency specification of ``cora-makesky`` and the channel axis it stands for
    (same attributes, defaults and results as ``cora/scripts/makesky.py:44-92``).

    ``freq = (start, stop, number)`` in MHz with ``freq_mode`` one of ``centre`` (default; the CHIME band
    800 -> 400 MHz in 1025 channels unless set), ``centre_nyquist`` or -- any other value -- ``edge``;
    ``channel_bin`` averages groups of neighbouring channels, ``channel_list`` (which takes precedence) or
    ``channel_range = (first, last + 1)`` select a subset afterwards."""

    def __init__(self):
        self.freq = (800.0, 400.0, 1025)
        self.freq_mode = "centre"
        self.channel_bin = 1
        self.channel_list = None
        self.channel_range = None

    def _calculate(self):
        """-> (channel centres, channel width) in MHz."""
        centres, width = _AXIS_MODES.get(self.freq_mode, _axis_edge)(*self.freq)
        nbin = self.channel_bin
        if nbin > 1:
            centres = centres.reshape(-1, nbin).mean(axis=1)
            width = width * nbin
        if self.channel_list is not None:
            centres = centres[self.channel_list]
        elif self.channel_range is not None:
            first, end = self.channel_range[0], self.channel_range[1]
            centres = centres[first:end]
        return centres, width

    frequencies = property(lambda self: self._calculate()[0], doc="The frequency centres in MHz.")
    freq_width = property(lambda self: self._calculate()[1], doc="The frequency width in MHz.")


def make_21cm(fstate, nside, pol="full", eor=False, oversample=None):
    """Gaussian simulation of the unresolved 21cm background (``makesky.py:325-344``).

    Returns ``float64[nfreq, npix]`` for ``pol == "none"``/``"zero"`` or ``[nfreq, 4, npix]``
    (Q = U = V = 0) for ``pol == "full"``."""
    from . import corr21cm

    cr = corr21cm.EoR21cm() if eor else corr21cm.Corr21cm()
    cr.nside = nside
    cr.frequencies = fstate.frequencies
    cr.oversample = oversample if oversample is not None else 3
    return cr.getpolsky() if pol == "full" else cr.getsky()


def make_gaussianfg(fstate, nside, pol="full", rng=None, seed=None, blocked=None):
    """Full-sky Gaussian random field for synchrotron emission (``makesky.py:349-390``).

    The polarised case builds the block-diagonal ``(npol*nfreq)^2`` covariance exactly as the
    reference does (T = FullSkySynchrotron, E = B = FullSkyPolarisedSynchrotron, V = 0,
    ``lmax = 3*nside``) so that the global jitter / eigenvalue clip of the root acts on the whole
    matrix.  Returns ``float64[nfreq, npol, npix]``.

    ``blocked``: keep the T / E / B blocks apart instead of forming the dense ``(4 nfreq)^2``
    matrices (``dist.ShardedPolSky``; same draws, global jitter, per-block Cholesky/eigen decision,
    V = 0).  Default: only when the dense array would not fit comfortably (> 8 GB) -- the dense
    path is the one with the reference's exact regulariser semantics."""
    from . import _dev, galaxy, hputil, skysim

    if blocked is None:
        blocked = pol == "full" and rng is None and 8.0 * (3 * nside + 1) * (4 * len(fstate.frequencies)) ** 2 > 8e9
    if blocked:
        if pol != "full" or rng is not None:
            raise ValueError("blocked=True needs pol='full' and the device generator (rng=None)")
        from . import dist as cdist

        sh = cdist.ShardedPolSky(nside, fstate.frequencies, rank=0, size=1)
        sky = sh.step(seed=int(np.random.randint(0, 2**31 - 1)) if seed is None else seed)
        return _dev.to_host(sky)

    t = _dev.torch()
    fsyn = galaxy.FullSkySynchrotron()
    fpol = galaxy.FullSkyPolarisedSynchrotron()
    fsyn.frequencies = fstate.frequencies
    nfreq = len(fsyn.frequencies)
    lmax = 3 * nside
    npol = 4 if pol == "full" else 1

    cv_fg = _dev.zeros((lmax + 1, npol, nfreq, npol, nfreq), t.float64)
    cv_fg[:, 0, :, 0, :] = skysim.clarray(fsyn.angular_powerspectrum, lmax, fsyn.nu_pixels, device_out=True)
    if pol == "full":
        cpol = skysim.clarray(fpol.angular_powerspectrum, lmax, fsyn.nu_pixels, device_out=True)
        cv_fg[:, 1, :, 1, :] = cpol
        cv_fg[:, 2, :, 2, :] = cpol
    cv_fg = cv_fg.reshape(lmax + 1, npol * nfreq, npol * nfreq)

    alms = skysim.mkfullsky(cv_fg, nside, alms=True, rng=rng, seed=seed, device_out=True)
    alms = alms.reshape(npol, nfreq, lmax + 1, lmax + 1).permute(1, 0, 2, 3).contiguous()
    return hputil.sphtrans_inv_sky(alms, nside)


def main(argv=None):
    import argparse

    ap = argparse.ArgumentParser(prog="cora_b200.makesky", description=__doc__.split("\n")[0])
    ap.add_argument("command", choices=["21cm", "gaussianfg"])
    ap.add_argument("--nside", type=int, default=128)
    ap.add_argument("--freq", type=float, nargs=3, default=(800.0, 400.0, 1024), metavar=("START", "STOP", "NUM"))
    ap.add_argument("--freq-mode", choices=["centre", "centre_nyquist", "edge"], default="centre")
    ap.add_argument("--pol", choices=["full", "zero", "none"], default="full")
    ap.add_argument("--eor", action="store_true")
    ap.add_argument("--oversample", type=int, default=None)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("filename")
    a = ap.parse_args(argv)
    fs = FreqState()
    fs.freq = (a.freq[0], a.freq[1], int(a.freq[2]))
    fs.freq_mode = a.freq_mode
    if a.seed is not None:
        np.random.seed(a.seed)
    if a.command == "21cm":
        m = make_21cm(fs, a.nside, a.pol, a.eor, a.oversample)
    else:
        m = make_gaussianfg(fs, a.nside, a.pol, seed=a.seed)
    np.save(a.filename, m)


if __name__ == "__main__":
    main()
