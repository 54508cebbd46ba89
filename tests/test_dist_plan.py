"""Host-side logic of the multi-GPU path (cora_b200/dist.py): partitions, slab offsets and the
all-to-all split sizes, exercised with a world_size-2 gloo group on CPU tensors.  The device
kernels (draw_apply_slabs / slabs_to_panel) are emulated in numpy from the same tables."""

import os
import socket

import numpy as np
import pytest

from oracle import skysim as osk


def _encode(l, m, nu):
    return l * 1e6 + m * 1e3 + nu + 1j * (l + m + nu)


def _emulate_send(plan, r):
    """What cora_b200_draw_apply_slabs writes on rank r, with alm[l, m, nu] = _encode(l, m, nu)."""
    base, width = plan.nu_tables(r)
    send = np.full(int(plan.rows[r]) * plan.nz, np.nan + 0j, dtype=np.complex128)
    row0 = plan.send_row0(r)
    for i, l in enumerate(plan.l_lists[r]):
        for m in range(l + 1):
            for nu in range(plan.nz):
                send[base[nu] + (row0[i] + m) * width[nu]] = _encode(int(l), m, nu)
    return send


def _emulate_panel(plan, s, recv):
    """What cora_b200_alm_slabs_to_panel produces on rank s."""
    lmax, cb = plan.lmax, int(plan.cb[s])
    off = plan.l_offsets(s)
    panel = np.full(((lmax + 1) * (lmax + 2) // 2, cb), np.nan + 0j, dtype=np.complex128)
    for l in range(lmax + 1):
        for m in range(l + 1):
            idx = m * (2 * lmax + 1 - m) // 2 + l
            panel[idx] = recv[off[l] + m * cb : off[l] + (m + 1) * cb]
    return panel


def test_block_partition_matches_caput_rule():
    from cora_b200 import dist as cdist

    for n, size in [(10, 3), (7, 8), (768, 8), (193, 2)]:
        got = [cdist.block_partition(n, size, r) for r in range(size)]
        assert got == [osk.partition(n, size, r) for r in range(size)]
        assert got[0][0] == 0 and got[-1][1] == n
        assert all(got[i][1] == got[i + 1][0] for i in range(size - 1))


@pytest.mark.parametrize("partition", ["interleaved", "block"])
@pytest.mark.parametrize("size,lmax,nz", [(1, 5, 3), (2, 7, 5), (3, 6, 7), (4, 9, 4), (8, 10, 3)])
def test_plan_roundtrip_numpy(partition, size, lmax, nz):
    """send slabs -> (manual) all-to-all -> panel puts every (l, m, nu) where the SHT expects it."""
    from cora_b200 import dist as cdist

    plan = cdist.ShardPlan(lmax, nz, size, partition)
    assert sorted(np.concatenate(plan.l_lists).tolist()) == list(range(lmax + 1))
    assert plan.nalm_total() == (lmax + 1) * (lmax + 2) // 2
    sends = [_emulate_send(plan, r) for r in range(size)]
    for r in range(size):
        assert not np.isnan(sends[r]).any()
        assert sum(plan.send_splits(r)) == sends[r].size
    for s in range(size):
        chunks = []
        for r in range(size):
            sp = plan.send_splits(r)
            o = int(np.sum(sp[:s]))
            chunks.append(sends[r][o : o + sp[s]])
            assert plan.recv_splits(s)[r] == sp[s]
        panel = _emulate_panel(plan, s, np.concatenate(chunks) if chunks else np.zeros(0, complex))
        assert not np.isnan(panel).any()
        for l in range(lmax + 1):
            for m in range(l + 1):
                idx = m * (2 * lmax + 1 - m) // 2 + l
                want = _encode(l, m, np.arange(plan.chan_lo[s], plan.chan_hi[s]))
                np.testing.assert_array_equal(panel[idx], want)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, size, port, lmax, nz, partition, q):
    import torch
    import torch.distributed as dist

    from cora_b200 import dist as cdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        plan = cdist.ShardPlan(lmax, nz, size, partition)
        send = torch.from_numpy(_emulate_send(plan, rank))
        recv = cdist.exchange(send, plan, rank).numpy()
        panel = _emulate_panel(plan, rank, recv)
        ok = True
        for l in range(lmax + 1):
            for m in range(l + 1):
                idx = m * (2 * lmax + 1 - m) // 2 + l
                want = _encode(l, m, np.arange(plan.chan_lo[rank], plan.chan_hi[rank]))
                ok &= bool(np.array_equal(panel[idx], want))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("partition", ["interleaved", "block"])
def test_exchange_gloo_world2(partition):
    """The real torch.distributed all_to_all_single (gloo, world_size 2) with the plan's split sizes."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 9, 5, partition, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


# --------------------------------------------------------------------------- alms=True: the all-gather
def _emulate_gather_slab(plan, r):
    """What cora_b200_draw_apply_slabs writes on rank r with the gather tables: one slab [rows_r][nz]."""
    base, width = plan.gather_tables()
    slab = np.full(int(plan.rows[r]) * plan.nz, np.nan + 0j, dtype=np.complex128)
    row0 = plan.send_row0(r)
    for i, l in enumerate(plan.l_lists[r]):
        for m in range(l + 1):
            for nu in range(plan.nz):
                slab[base[nu] + (row0[i] + m) * width[nu]] = _encode(int(l), m, nu)
    return slab


def _check_gathered(plan, allb):
    off = plan.gather_l_offsets()
    for l in range(plan.lmax + 1):
        for m in range(l + 1):
            row = allb[off[l] + m * plan.nz : off[l] + (m + 1) * plan.nz]
            if not np.array_equal(row, _encode(l, m, np.arange(plan.nz))):
                return False
    return True


@pytest.mark.parametrize("partition", ["interleaved", "block"])
@pytest.mark.parametrize("size,lmax,nz", [(2, 7, 5), (3, 6, 4), (8, 10, 3)])
def test_gather_tables_numpy(partition, size, lmax, nz):
    from cora_b200 import dist as cdist

    plan = cdist.ShardPlan(lmax, nz, size, partition)
    n = plan.gather_rows() * nz
    allb = np.zeros(size * n, dtype=np.complex128)
    for r in range(size):
        slab = _emulate_gather_slab(plan, r)
        assert not np.isnan(slab).any()
        allb[r * n : r * n + slab.size] = slab
    assert _check_gathered(plan, allb)


def _gloo_gather_worker(rank, size, port, q):
    import torch
    import torch.distributed as dist

    from cora_b200 import dist as cdist
    from cora_b200 import mpiarray

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        ok = True
        for partition in ("interleaved", "block"):
            plan = cdist.ShardPlan(9, 5, size, partition)
            allb = cdist.allgather_alm(torch.from_numpy(_emulate_gather_slab(plan, rank)), plan, rank).numpy()
            ok &= _check_gathered(plan, allb)
        # the MPIArray shim: wrap / allgather / redistribute against a global array every rank can build
        glob = np.arange(7 * 3 * 5, dtype=np.float64).reshape(7, 3, 5) + 0.25
        lo, hi = mpiarray.split_block(7, size, rank)
        a = mpiarray.MPIArray.wrap(glob[lo:hi].copy(), axis=0)
        ok &= a.global_shape == (7, 3, 5) and a.local_offset == (lo, 0, 0)
        ok &= bool(np.array_equal(a.allgather(), glob))
        b = a.redistribute(axis=2)
        lo2, hi2 = mpiarray.split_block(5, size, rank)
        ok &= b.axis == 2 and bool(np.array_equal(b.local_array, glob[:, :, lo2:hi2]))
        ok &= bool(np.array_equal(b.redistribute(axis=0).local_array, glob[lo:hi]))
        ok &= [g for _, g in a.enumerate(0)] == list(range(lo, hi))
        z = mpiarray.zeros((7, 2), dtype=np.complex128, axis=0)
        ok &= z.local_array.shape == (hi - lo, 2) and z.dtype == np.complex128
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_allgather_and_mpiarray_gloo_world2():
    """alms=True all-gather (plan tables + torch.distributed.all_gather) and the MPIArray shim on a real
    world_size-2 gloo group."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_mpiarray_single_rank():
    from cora_b200 import mpiarray

    a = mpiarray.MPIArray.wrap(np.arange(12.0).reshape(4, 3), axis=0)
    assert a.global_shape == (4, 3) and a.local_shape == (4, 3) and mpiarray.is_distributed(a)
    assert np.array_equal(a.allgather(), np.arange(12.0).reshape(4, 3))
    assert a.redistribute(axis=1).axis == 1
    assert not mpiarray.is_distributed(np.zeros(3))


@pytest.mark.parametrize("nz", [1, 7, 256, 1024])
@pytest.mark.parametrize("size", [1, 2, 3, 8])
def test_fill_tiles_are_dealt_interleaved_and_cover_everything(nz, size):
    """The sharded 21cm fill deals its channel-pair tiles rank, rank + G, ... (dist.fill_tile_share): every tile of
    cora_b200_cl_fill_21cm_ntiles(nz) exactly once over the ranks, shares equal to within one tile."""
    from cora_b200 import _lib
    from cora_b200 import dist as cdist

    ntile = int(_lib.load().cora_b200_cl_fill_21cm_ntiles(nz))
    assert ntile >= (nz * (nz + 1) // 2 + 3) // 4
    seen = np.zeros(ntile, dtype=int)
    counts = []
    for r in range(size):
        t0, n, step = cdist.fill_tile_share(ntile, r, size)
        assert step == size
        idx = t0 + step * np.arange(n)
        assert n == 0 or idx.max() < ntile
        seen[idx] += 1
        counts.append(n)
    assert np.all(seen == 1)
    assert max(counts) - min(counts) <= 1
