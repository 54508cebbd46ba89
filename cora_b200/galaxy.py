"""Full-sky synchrotron spectra used by ``cora-makesky gaussianfg``
(parameter sets of ``cora/foreground/galaxy.py:20-40``)."""

from . import gaussianfg


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
