"""cora_b200 -- B200-native (sm_100a) full-sky Gaussian field generator.

Drop-in for the hot path of radiocosmology/cora: ``skysim.clarray`` / ``skysim.mkfullsky``
as driven by ``cora-makesky 21cm`` and ``cora-makesky gaussianfg``.  Module names, function
names, argument meaning and error behaviour follow the reference:

    cora.core.skysim      -> cora_b200.skysim      (clarray, mkfullsky, mkconstrained)
    cora.util.nputil      -> cora_b200.nputil      (matrix_root_manynull, complex_std_normal)
    cora.util.hputil      -> cora_b200.hputil      (pack_alm, unpack_alm, sphtrans_inv_*, sphtrans_real[_pol],
                                                    sphtrans_complex[_pol], sphtrans_sky, sph_ps, ang_positions,
                                                    smoothing, coord_g2c / coord_c2g)
    cora.core.maps        -> cora_b200.maps        (Map3d/Sky3d: getsky, getpolsky, getalms)
    cora.foreground.*     -> cora_b200.gaussianfg, cora_b200.galaxy (SCK spectra; ConstrainedGalaxy, map_variance)
    cora.signal.corrfunc  -> cora_b200.corrfunc    (corr_to_clarray, legendre_array: xi(r) -> C_l(chi, chi'))
    cora.signal.corr21cm  -> cora_b200.corr21cm    (Corr21cm, EoR21cm)
    cora.util.cosmology   -> cora_b200.cosmology   (Cosmology, host side)
    cora.scripts.makesky  -> cora_b200.makesky     (FreqState, 21cm / gaussianfg drivers)
    caput MPIArray path   -> cora_b200.dist        (ShardedSky, ShardedPolSky, mkfullsky_sharded / mkfullsky_mpi: one
                                                    process per GPU, exchange fused into the kernels over peer memory --
                                                    cora_b200.peer), cora_b200.mpiarray (MPIArray over torch.distributed)
    healpy helpers        -> cora_b200.healpix     (reorder, ud_grade, get_interp_val, Galactic <-> celestial; host)
    map output            -> cora_b200.mapio       (frequency-sharded map writer, cora/scripts/makesky.py:412-450)

All arithmetic of the path runs in hand-written CUDA kernels behind the C ABI of
``include/cora_b200.h`` (``libcora_b200.so``, bound with ctypes in ``_lib``).  There is no CPU
fallback: without the library or a CUDA device the entry points raise.
"""

__version__ = "0.1.0"
