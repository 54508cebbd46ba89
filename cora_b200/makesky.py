"""Drivers of record for the path: ``cora-makesky 21cm`` and ``cora-makesky gaussianfg``.

Mirrors ``cora/scripts/makesky.py``: ``FreqState`` (``:44-92``), the ``21cm`` command
(``:313-344``) and the ``gaussianfg`` command (``:347-390``).  The functions below return the
``(nfreq, npol, npix)`` arrays the commands write; ``write_map`` (``:412-450``) is ``cora_b200.mapio``
(frequency-sharded ``.npy`` + JSON index maps; ``tools/map_to_hdf5.py`` converts to the reference's HDF5 layout).

    python -m cora_b200.makesky 21cm --nside 64 --freq 800 400 32 --pol none out.npy
    torchrun --nproc-per-node 8 -m cora_b200.makesky 21cm --nside 1024 --freq 800 400 2048 --pol none outdir
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


def make_sharded(command, fstate, nside, pol="full", eor=False, oversample=None, seed=0):
    """The same generators, one rank of a multi-GPU run (``torchrun``; also works with one rank): returns
    ``(this rank's channels as a CUDA tensor [freq_local, npix] or [freq_local, 4, npix], first global channel)``.
    The l-sharded root / apply and the channel-sharded SHT of ``dist.ShardedSky`` / ``ShardedPolSky``."""
    from . import corr21cm, galaxy
    from . import dist as cdist

    if command == "21cm":
        model = corr21cm.EoR21cm() if eor else corr21cm.Corr21cm()
        zromb = oversample if oversample is not None else 3
        sh = cdist.ShardedSky(model, nside, fstate.frequencies, lmax=3 * nside - 1, zromb=zromb)
    elif pol == "full":
        sh = cdist.ShardedPolSky(nside, fstate.frequencies)
        sh.exchange = "p2p"
    else:
        sh = cdist.ShardedSky(galaxy.FullSkySynchrotron(), nside, fstate.frequencies, lmax=3 * nside)
    sky = sh.step(seed=seed).clone()
    if getattr(sh, "exchange", None) == "p2p" and sh.size > 1:
        sh.peers.check()
    first = int(sh.plan.chan_lo[sh.rank])
    return sky, first, sh


def main(argv=None):
    import argparse
    import os

    ap = argparse.ArgumentParser(prog="cora_b200.makesky", description=__doc__.split("\n")[0])
    ap.add_argument("command", choices=["21cm", "gaussianfg"])
    ap.add_argument("--nside", type=int, default=128)
    ap.add_argument("--freq", type=float, nargs=3, default=(800.0, 400.0, 1024), metavar=("START", "STOP", "NUM"))
    ap.add_argument("--freq-mode", choices=["centre", "centre_nyquist", "edge"], default="centre")
    ap.add_argument("--pol", choices=["full", "zero", "none"], default="full")
    ap.add_argument("--eor", action="store_true")
    ap.add_argument("--oversample", type=int, default=None)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--sharded", action="store_true",
                    help="generate with the l-/channel-sharded engine and write the frequency-sharded map directory "
                         "(implied under torchrun with more than one rank)")
    ap.add_argument("filename", help="X.npy: one array (single GPU); anything else: a map directory (cora_b200.mapio)")
    a = ap.parse_args(argv)
    fs = FreqState()
    fs.freq = (a.freq[0], a.freq[1], int(a.freq[2]))
    fs.freq_mode = a.freq_mode
    if a.seed is not None:
        np.random.seed(a.seed)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or a.sharded:
        import torch

        from . import mapio

        rank = int(os.environ.get("RANK", "0"))
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        sky, first, sh = make_sharded(a.command, fs, a.nside, a.pol, a.eor, a.oversample, seed=a.seed or 0)
        if a.command == "21cm" and a.pol == "zero":
            sky = sky[:, None, :]          # cora-makesky --pol zero: intensity only is generated, write_map pads (makesky.py:422-426)
        mapio.write_map(a.filename, sky, fs.frequencies, fs.freq_width, include_pol=(a.pol != "none"), freq_start=first,
                        rank=rank, size=world)
        if world > 1:
            dist.barrier()
            if getattr(sh, "exchange", None) == "p2p":
                sh.peers.close()
            dist.destroy_process_group()
        return
    if a.command == "21cm":
        m = make_21cm(fs, a.nside, a.pol, a.eor, a.oversample)
    else:
        m = make_gaussianfg(fs, a.nside, a.pol, seed=a.seed)
    if a.filename.endswith(".npy"):
        np.save(a.filename, m)
    else:
        from . import mapio

        mapio.write_map(a.filename, m, fs.frequencies, fs.freq_width, include_pol=(a.pol != "none"))


if __name__ == "__main__":
    main()
