"""GPU parity of the multi-GPU kernels on ONE device: every rank's l-shard is run in turn through
cora_b200_draw_apply_slabs, the all-to-all is done by slicing with the plan's split sizes, and
cora_b200_alm_slabs_to_panel + the SHT must reproduce the single-GPU mkfullsky exactly."""

import os

import numpy as np
import pytest

from oracle import skysim as osk
from oracle import spectra as osp

pytestmark = pytest.mark.gpu


def _manual_alltoall(plan, sends):
    import torch

    recvs = []
    for s in range(plan.size):
        chunks = []
        for r in range(plan.size):
            sp = plan.send_splits(r)
            o = int(np.sum(sp[:s]))
            chunks.append(sends[r][o : o + sp[s]])
        recvs.append(torch.cat(chunks))
    return recvs


@pytest.mark.parametrize("size,partition", [(1, "interleaved"), (2, "interleaved"), (3, "block"), (4, "interleaved")])
def test_sharded_equals_single(size, partition):
    import torch
    from cora_b200 import dist as cdist
    from cora_b200 import galaxy, skysim

    nside, nz = 16, 10  # nz not divisible by 3 or 4: ragged channel blocks
    lmax = 3 * nside - 1
    freq = np.linspace(800.0, 400.0, nz, endpoint=False)
    model = galaxy.FullSkySynchrotron()
    cla = skysim.clarray(model.angular_powerspectrum, lmax, freq, device_out=True)
    ref = skysim.mkfullsky(cla, nside, seed=5, device_out=True)

    shards = [cdist.ShardedSky(model, nside, freq, lmax=lmax, rank=r, size=size, partition=partition) for r in range(size)]
    sends = []
    for sh in shards:
        c_loc = sh.fill()
        np.testing.assert_array_equal(c_loc.cpu().numpy(), cla[torch.from_numpy(sh.l_list.astype(np.int64)).cuda()].cpu().numpy())
        sends.append(sh.alm_local(c_loc, seed=5))
    recvs = _manual_alltoall(shards[0].plan, sends)
    for s, sh in enumerate(shards):
        sky = sh.synthesize(recvs[s])
        lo, hi = int(sh.plan.chan_lo[s]), int(sh.plan.chan_hi[s])
        assert sky.shape == (hi - lo, 12 * nside**2)
        # not bit-equal: the phase stage transforms channels in pairs, and the pairing follows the
        # channel offset inside each rank's block
        want = ref[lo:hi].cpu().numpy()
        np.testing.assert_allclose(sky.cpu().numpy(), want, rtol=0, atol=1e-13 * np.abs(want).max())


def test_sharded_injected_draws_match_oracle():
    """identical-draw parity through the sharded path (2 ranks emulated): oracle roots and draws."""
    import torch
    from cora_b200 import dist as cdist
    from cora_b200 import galaxy

    nside, nz = 8, 6
    lmax = 3 * nside - 1
    L = lmax + 1
    freq = np.linspace(800.0, 400.0, nz, endpoint=False)
    cla_o = osk.clarray(osp.full_sky_synchrotron().angular_powerspectrum, lmax, freq)
    rec = {}
    ref = osk.mkfullsky(cla_o, nside, rng=np.random.default_rng(3), record=rec)
    size = 2
    shards = [cdist.ShardedSky(galaxy.FullSkySynchrotron(), nside, freq, lmax=lmax, rank=r, size=size) for r in range(size)]
    sends = []
    for sh in shards:
        g = np.zeros((sh.nl, nz, L), dtype=np.complex128)
        roots = np.zeros((sh.nl, nz, nz))
        for i, l in enumerate(sh.l_list):
            g[i, :, : l + 1] = rec["gauss"][l]
            roots[i] = rec["roots"][l]
        sends.append(sh.alm_local(None, gauss=g, roots=roots))
    recvs = _manual_alltoall(shards[0].plan, sends)
    got = torch.cat([sh.synthesize(recvs[s]) for s, sh in enumerate(shards)]).cpu().numpy()
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-10


# --------------------------------------------------------------------------- alms=True (all-gather) and MPIArray input
@pytest.mark.parametrize("size,partition", [(2, "interleaved"), (3, "block")])
def test_alms_allgather_virtual_ranks(size, partition):
    """mkfullsky(alms=True) over ranks (skysim.py:123-125): every rank's gather slab, padded and stacked the way
    torch.distributed.all_gather delivers them, must unpack to the single-GPU alm array."""
    import torch
    from cora_b200 import _dev, _lib, galaxy, hputil, skysim
    from cora_b200 import dist as cdist

    nside, nz = 8, 5
    lmax = 3 * nside - 1
    L = lmax + 1
    freq = np.linspace(800.0, 400.0, nz, endpoint=False)
    model = galaxy.FullSkySynchrotron()
    cla = skysim.clarray(model.angular_powerspectrum, lmax, freq, device_out=True)
    ref = skysim.mkfullsky(cla, nside, alms=True, seed=9)
    assert ref.shape == (nz, 1, L, L)
    plan = cdist.ShardPlan(lmax, nz, size, partition)
    n = plan.gather_rows() * nz
    allb = torch.zeros(size * n, dtype=torch.complex128, device="cuda")
    base, width = plan.gather_tables()
    for r in range(size):
        sh = cdist.ShardedSky(model, nside, freq, lmax=lmax, rank=r, size=size, partition=partition, exchange="collective")
        slab = torch.empty(int(plan.rows[r]) * nz, dtype=torch.complex128, device="cuda")
        sh.alm_local(sh.fill(), seed=9, slab_tables=(_dev.to_device(base, torch.int64), _dev.to_device(width, torch.int32), slab))
        allb[r * n : r * n + slab.numel()] = slab
    panel = torch.empty((L * (L + 1) // 2, nz), dtype=torch.complex128, device="cuda")
    loff = _dev.to_device(plan.gather_l_offsets(), torch.int64)
    _lib.call("cora_b200_alm_slabs_to_panel", _lib.ptr(allb), _lib.ptr(loff), lmax, nz, _lib.ptr(panel), nz, 0, _lib.stream_ptr())
    got = hputil.panel_to_dense(panel, lmax, nz).reshape(nz, 1, L, L).cpu().numpy()
    np.testing.assert_array_equal(got, ref)


def test_mkfullsky_accepts_distributed_corr_single_rank():
    """The MPIArray branch of mkfullsky (skysim.py:97-103,128-134) with one rank: same draws from the same rng as
    the dense call, maps come back wrapped and distributed over frequency; alms=True returns the full array."""
    from cora_b200 import galaxy, mpiarray, skysim

    nside, nz = 8, 4
    lmax = 3 * nside - 1
    freq = np.linspace(800.0, 400.0, nz, endpoint=False)
    cla = skysim.clarray(galaxy.FullSkySynchrotron().angular_powerspectrum, lmax, freq)
    want = skysim.mkfullsky(cla, nside, rng=np.random.default_rng(11))
    dcorr = mpiarray.MPIArray.wrap(cla, axis=0)
    got = skysim.mkfullsky(dcorr, nside, rng=np.random.default_rng(11))
    assert mpiarray.is_distributed(got) and got.axis == 0 and got.global_shape == want.shape
    np.testing.assert_allclose(got.local_array, want, rtol=0, atol=1e-13 * np.abs(want).max())
    alm_want = skysim.mkfullsky(cla, nside, alms=True, rng=np.random.default_rng(11))
    alm_got = skysim.mkfullsky(dcorr, nside, alms=True, rng=np.random.default_rng(11))
    np.testing.assert_allclose(alm_got, alm_want, rtol=0, atol=1e-13 * np.abs(alm_want).max())
    with pytest.raises(Exception, match="incorrect shape"):
        skysim.mkfullsky(mpiarray.MPIArray.wrap(np.zeros((lmax + 1, nz, nz + 1)), axis=0), nside)


# --------------------------------------------------------------------------- fused exchange (p2p)
def _run_virtual_p2p(model, nside, freq, lmax, size, partition, seed, zromb=3, steps=1):
    """Drive the multi-GPU p2p path with `size` virtual ranks on one GPU (cora_b200.peer.LocalPeers):
    the same kernels and pointer tables as the real one-process-per-GPU run, phases in lock step."""
    import torch
    from cora_b200 import dist as cdist
    from cora_b200 import peer

    lp = peer.LocalPeers(size)
    shards = [cdist.ShardedSky(model, nside, freq, lmax=lmax, zromb=zromb, rank=r, size=size, partition=partition,
                               exchange="p2p", peers=lp.view(r)) for r in range(size)]
    for sh in shards:
        sh._p2p_setup()
    out = None
    for it in range(steps):
        k = it & 1
        for sh in shards:
            sh.p2p_fill(k)
        torch.cuda.synchronize()
        for sh in shards:
            sh.p2p_alm(k, seed=seed + it)
        torch.cuda.synchronize()
        out = torch.cat([sh.p2p_sht(k).clone() for sh in shards])
    return shards, out


@pytest.mark.parametrize("size,partition", [(2, "interleaved"), (3, "block"), (4, "interleaved")])
def test_p2p_virtual_ranks_sck(size, partition):
    """apply -> peer PANEL stores: maps equal the single-GPU maps (SCK model: local l-sharded fill)."""
    from cora_b200 import galaxy, skysim

    nside, nz = 16, 10
    lmax = 3 * nside - 1
    freq = np.linspace(800.0, 400.0, nz, endpoint=False)
    model = galaxy.FullSkySynchrotron()
    cla = skysim.clarray(model.angular_powerspectrum, lmax, freq, device_out=True)
    # two steps: the second uses the other buffer set (double buffering) and seed + 1
    ref = skysim.mkfullsky(cla, nside, seed=6, device_out=True).cpu().numpy()
    _, got = _run_virtual_p2p(model, nside, freq, lmax, size, partition, seed=5, steps=2)
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=1e-13 * np.abs(ref).max())


@pytest.mark.parametrize("size,partition", [(2, "interleaved"), (3, "block")])
def test_p2p_virtual_ranks_21cm_pair_sharded_fill(size, partition, gpu_corr21cm_dist):
    """21cm: the fill is sharded over channel pairs and scatters C_l rows to the owners of l.
    The scattered rows must be bit-identical to clarray's, and the maps equal the single-GPU maps."""
    import torch
    from cora_b200 import skysim

    model = gpu_corr21cm_dist
    nside, nz = 8, 7
    lmax = 3 * nside - 1
    freq = np.linspace(800.0, 700.0, nz, endpoint=False)
    cla = skysim.clarray(model.angular_powerspectrum, lmax, freq, device_out=True)
    ref = skysim.mkfullsky(cla, nside, seed=9, device_out=True).cpu().numpy()
    shards, got = _run_virtual_p2p(model, nside, freq, lmax, size, partition, seed=9)
    for sh in shards:
        mine = sh._p2p["cla"][0].tensor((sh.nl, nz, nz), torch.float64).cpu().numpy()
        want = cla[torch.from_numpy(sh.l_list.astype(np.int64)).cuda()].cpu().numpy()
        # the sharded fill writes the lower triangles only (all the root stage reads); those are bit-identical
        np.testing.assert_array_equal(np.tril(mine), np.tril(want))
        assert np.all(np.triu(mine, 1) == 0)
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=1e-13 * np.abs(ref).max())


@pytest.fixture(scope="module")
def gpu_corr21cm_dist():
    from cora_b200 import corr21cm

    return corr21cm.Corr21cm()


def _p2p_worker(rank, size, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(size))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=size, device_id=torch.device("cuda", rank))
    try:
        from cora_b200 import corr21cm, galaxy
        from cora_b200 import dist as cdist

        res = {}
        for name, model, nside, nz in (("sck", galaxy.FullSkySynchrotron(), 16, 10), ("21cm", corr21cm.Corr21cm(), 8, 7)):
            lmax = 3 * nside - 1
            freq = np.linspace(800.0, 700.0, nz, endpoint=False)
            sh = cdist.ShardedSky(model, nside, freq, lmax=lmax, rank=rank, size=size)
            assert sh.exchange == "p2p"
            for it in range(3):      # exercises both buffer sets and the barrier epochs
                sky = sh.step(seed=20 + it)
            sh.peers.check()
            sh2 = cdist.ShardedSky(model, nside, freq, lmax=lmax, rank=rank, size=size, exchange="collective")
            sky2 = sh2.step(seed=22)
            res[name] = (sky.cpu().numpy(), sky2.cpu().numpy())
            sh.peers.close()
        # the reference's MPI call: l-distributed corr in, frequency-distributed maps out / all-gathered alms
        from cora_b200 import mpiarray, skysim

        nside, nz = 8, 6
        lmax = 3 * nside - 1
        freq = np.linspace(800.0, 400.0, nz, endpoint=False)
        cla = skysim.clarray(galaxy.FullSkySynchrotron().angular_powerspectrum, lmax, freq)
        lo, hi = mpiarray.split_block(lmax + 1, size, rank)
        dcorr = mpiarray.MPIArray.wrap(cla[lo:hi].copy(), axis=0)
        sky = skysim.mkfullsky(dcorr, nside, seed=77)
        alm = skysim.mkfullsky(dcorr, nside, alms=True, seed=77)
        res["mpi"] = (sky.allgather(), skysim.mkfullsky(cla, nside, seed=77), alm, skysim.mkfullsky(cla, nside, alms=True, seed=77))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_p2p_two_processes():
    """Real peer memory: two processes on two GPUs (IPC-mapped buffers, NVLink stores, flag
    barrier); the fused exchange must give the same maps as the NCCL all-to-all path."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        for name in ("sck", "21cm"):
            a, b = got[r][name]
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-13 * np.abs(b).max())
        sky, sky_ref, alm, alm_ref = got[r]["mpi"]
        np.testing.assert_allclose(sky, sky_ref, rtol=0, atol=1e-13 * np.abs(sky_ref).max())
        np.testing.assert_array_equal(alm, alm_ref)


# --------------------------------------------------------------------------- polarised, block-sharded
def _run_virtual_pol(nside, freq, size, partition, seed, steps=1):
    import torch
    from cora_b200 import dist as cdist
    from cora_b200 import peer

    lp = peer.LocalPeers(size)
    shards = [cdist.ShardedPolSky(nside, freq, rank=r, size=size, partition=partition, peers=lp.view(r)) for r in range(size)]
    out = None
    for it in range(steps):
        k = it & 1
        for sh in shards:
            sh.alm_phase(k, seed=seed + it)
        torch.cuda.synchronize()
        out = torch.cat([sh.sht_phase(k).clone() for sh in shards])
    return out


def test_pol_blocked_equals_dense_makesky():
    """gaussianfg --pol full: the block formulation (T, E, B roots apart, Philox counters offset per
    block, global jitter) reproduces the dense (4 nfreq)^2 formulation of makesky.py:368-387 on the
    Cholesky branch (few channels), where both are defined without eigen-degeneracy freedom."""
    from cora_b200 import makesky

    fs = makesky.FreqState()
    fs.freq = (800.0, 700.0, 4)
    nside = 8
    dense = makesky.make_gaussianfg(fs, nside, pol="full", seed=11, blocked=False)
    blk = makesky.make_gaussianfg(fs, nside, pol="full", seed=11, blocked=True)
    assert blk.shape == dense.shape == (4, 4, 12 * nside**2)
    for p in range(3):
        scale = np.abs(dense[:, p]).max()
        assert np.max(np.abs(blk[:, p] - dense[:, p])) / scale < 1e-7, p   # own roots of ill-conditioned SCK blocks
    assert np.all(blk[:, 3] == 0) and np.abs(dense[:, 3]).max() < 1e-4 * np.abs(dense[:, 0]).max()


@pytest.mark.parametrize("size,partition", [(2, "interleaved"), (3, "block")])
def test_pol_sharded_equals_single(size, partition):
    nside, nz = 8, 7
    freq = np.linspace(800.0, 600.0, nz, endpoint=False)
    ref = _run_virtual_pol(nside, freq, 1, "interleaved", seed=4, steps=2).cpu().numpy()
    got = _run_virtual_pol(nside, freq, size, partition, seed=4, steps=2).cpu().numpy()
    assert got.shape == ref.shape == (nz, 4, 12 * nside**2)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13 * np.abs(ref).max())
    assert np.abs(ref[:, 1]).max() > 0 and np.abs(ref[:, 2]).max() > 0


def test_pol_blocked_eigen_branch_statistics():
    """Many channels -> both blocks take the eigen branch.  T: the GPU anafast of the maps follows
    C_l(nu, nu) within cosmic variance; Q/U: <Q^2 + U^2> follows sum (2l+1)/(4 pi) (C_EE + C_BB)
    (dominated by l = 2..4, hence the wide band); V = 0."""
    import torch
    from cora_b200 import galaxy, hputil, skysim

    nside, nz = 16, 24
    freq = np.linspace(800.0, 400.0, nz, endpoint=False)
    sky_d = _run_virtual_pol(nside, freq, 1, "interleaved", seed=5)
    sky = sky_d.cpu().numpy()
    lmax = 3 * nside
    la = 2 * nside
    clT = skysim.clarray(galaxy.FullSkySynchrotron().angular_powerspectrum, lmax, freq)
    alm = hputil.panel_to_dense(hputil.map2alm_device(sky_d[:, 0].contiguous(), nside, la, iter=3), la, nz).cpu().numpy()
    l = np.arange(6, la + 1)
    for i in (0, nz // 2, nz - 1):
        prod = np.abs(alm[i]) ** 2
        est = (prod[:, 0] + 2.0 * prod[:, 1:].sum(axis=1)) / (2.0 * np.arange(la + 1) + 1.0)
        want = clT[l, i, i] * (2.0 * l + 0.5) / (2.0 * l + 1.0)
        assert np.all(np.abs(est[l] - want) < 6.0 * want * np.sqrt(2.0 / (2.0 * l + 1.0))), i
    clP = skysim.clarray(galaxy.FullSkyPolarisedSynchrotron().angular_powerspectrum, lmax, freq)
    ll = np.arange(lmax + 1)
    for i in (0, nz - 1):
        var = 2 * np.sum(((2 * ll + 1) * clP[:, i, i])[2:]) / (4 * np.pi)
        got = (sky[i, 1] ** 2 + sky[i, 2] ** 2).mean()
        assert 0.25 < got / var < 4.0, (i, got / var)
    assert np.all(sky[:, 3] == 0)
