#!/usr/bin/env python
"""Benchmark of the full-sky Gaussian field path (BASELINE.json: map voxels/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

One "step" = one pass of the hot path over the workload: C_l(nu,nu') fill -> batched root ->
Philox draws + apply -> (N > 1: all-to-all) -> inverse SHT -> maps resident in HBM.  The
one-off 21cm P(k) DCT table (the reference's ``_aps_cache``, corr.py:915-942) is built once
before the timed region in both arms and reported separately.

``--impl reference`` times the CPU restatement of the reference path (``oracle/``; numpy/scipy,
all host cores through a fork pool + BLAS threads) on a bounded sample of the same workload
and extrapolates to the whole workload (the sample is described in ``cpu_baseline.sample``).
healpy cannot be installed here, so the SHT leg is the restatement, not healpy.

Prints ONE JSON line (rank 0).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (model, nside, nchan, f_start, f_stop)   (BASELINE.json configs; frequencies = makesky "centre" mode)
    "c1": ("gaussianfg", 64, 32, 800.0, 400.0),
    "c2": ("21cm", 256, 256, 800.0, 400.0),
    "c3": ("21cm", 512, 1024, 800.0, 400.0),
    "c3fg": ("gaussianfg", 512, 1024, 800.0, 400.0),
    "c4": ("gaussianfg_pol", 512, 1024, 800.0, 400.0),
    "c5": ("21cm", 1024, 2048, 800.0, 400.0),
}
WORKLOAD_TEXT = {
    "c1": "cora-makesky gaussianfg nside=64, 32 channels 400-800 MHz, unpolarised",
    "c2": "cora-makesky 21cm (Corr21cm) nside=256, 256 channels 400-800 MHz",
    "c3": "cora-makesky 21cm nside=512, 1024 channels 400-800 MHz",
    "c3fg": "cora-makesky gaussianfg --pol none nside=512, 1024 channels 400-800 MHz (the foreground half of config 3)",
    "c4": "cora-makesky gaussianfg --pol full (T/E/B, spin-2 SHT) nside=512, 1024 channels 400-800 MHz",
    "c5": "cora-makesky 21cm nside=1024, 2048 channels 400-800 MHz",
}
# The default is the largest configuration that fits one GPU (config 3's 21cm half, nside 512 x 1024
# channels, ~0.45 s/step on one B200), so that the driver's 1/2/4/8-GPU runs measure the regime the north star is
# about; `--workload c2` is BASELINE.json's single-GPU config (16 ms/step) and is also measured in every default
# run (the "secondary" block of the line).
DEFAULT_WORKLOAD = "c3"
METRIC = "full-sky map voxels/sec (pixels x channels)"
UNIT = "voxels/s"


def workload_params(name):
    model, nside, nchan, f0, f1 = WORKLOADS[name]
    lmax = 3 * nside if model.startswith("gaussianfg") else 3 * nside - 1
    freq = np.linspace(f0, f1, nchan, endpoint=False)
    return dict(model=model, nside=nside, nchan=nchan, lmax=lmax, freq=freq, npix=12 * nside * nside, zromb=3)


def sht_flops(nside, lmax, nchan):
    """2 real FMAs per (ring pair, l, m, channel) = 4 flop x ceil((4 nside - 1)/2) x nalm x nchan."""
    L = lmax + 1
    return 4.0 * ((4 * nside - 1 + 1) // 2) * (L * (L + 1) / 2.0) * nchan


# =============================================================================== CPU arm
def _cpu_fill_rows(args):
    """worker: channel-averaged C_l of a block of channel rows (all columns) for one l: aps evaluation at the
    (9 nrow) x (9 nz) sample pairs + both Romberg passes (what clarray does for these entries)."""
    import scipy.integrate as si

    l, za_rows, za, nrow, nz, zint, dx, h = args
    clt = _CPU["aps"](np.array([l])[:, None, None], za_rows[None, :, None], za[None, None, :])
    clt = np.broadcast_to(clt, (1, nrow * zint, nz * zint)).reshape(1, nrow, zint, nz, zint)
    clt = si.romb(clt, dx=dx, axis=4)
    clt = si.romb(clt, dx=dx, axis=2)
    return float(clt.sum()) / (2 * h) ** 2


def _cpu_leg_one(args):
    """worker: Legendre stage of one m for a block of channels (lambda recursion + contraction)."""
    from oracle import sht as osht

    m, lmax, nchan, seed = args
    g = _CPU["geom"]
    lam = osht.lambda_lm(lmax, m, g["cth"], g["sth"])
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((nchan, lmax - m + 1)) + 1j * rng.standard_normal((nchan, lmax - m + 1))
    return (lam.T @ a.T).shape[0]


_CPU = {}


def cpu_reference_setup(wp):
    """One-off state of the CPU arm (21cm: the P(k) DCT tables, like the reference's cache)."""
    from oracle import sht as osht
    from oracle import spectra as osp

    t0 = time.time()
    if wp["model"] == "21cm":
        c = osp.Corr21cm()
        c.tables()
        _CPU["aps"] = c.angular_powerspectrum
    else:
        _CPU["aps"] = osp.full_sky_synchrotron().angular_powerspectrum
    _CPU["geom"] = osht.ring_geometry(wp["nside"])
    return time.time() - t0


def cpu_reference_step(wp, pool, ncores, budget=1.0):
    """Time a BOUNDED sample of every stage of the path on the host and extrapolate linearly by work to the whole
    workload (the C3 workload would take hours on the host: SURVEY 8d prescribes the sampling).  ``budget`` scales
    the sample (1.0: about 20-60 s of wall time on 8-32 cores at any workload).

    Returns (extrapolated seconds for the whole workload, per-stage dict, sample description)."""
    import scipy.linalg as la

    from oracle import nputil as onp
    from oracle import sht as osht

    nz, lmax, nside, npix = wp["nchan"], wp["lmax"], wp["nside"], wp["npix"]
    L = lmax + 1
    freq = wp["freq"]
    zs = np.sort(freq)
    h = abs(zs[1] - zs[0]) / 2.0
    zint = 2 ** wp["zromb"] + 1
    dx = 2.0 * h / 2 ** wp["zromb"]
    za = (freq[:, None] + np.linspace(-h, h, zint)[None, :]).flatten()

    # --- stage 1: C_l fill.  Work = evaluations of the spectrum; sample: n_l values of l x a subset of channel
    # rows (all columns), dealt to the pool in blocks of rows.
    evals_budget = 2.0e7 * ncores * budget
    n_l = 8 if nz >= 1024 else 24
    ls = np.unique(np.linspace(1, lmax, n_l).astype(int))
    rows_per_task = max(1, min(nz, int(2.0e6 // (zint * zint * nz)) or 1))
    ntask = max(ncores, int(evals_budget / (len(ls) * rows_per_task * zint * zint * nz)))
    ntask = min(ntask, max(1, nz // rows_per_task))
    starts = np.unique(np.linspace(0, nz - rows_per_task, ntask).astype(int))
    tasks = []
    for l in ls:
        for r0 in starts:
            zr = (freq[r0:r0 + rows_per_task, None] + np.linspace(-h, h, zint)[None, :]).flatten()
            tasks.append((int(l), zr, za, rows_per_task, nz, zint, dx, h))
    t0 = time.time()
    pool.map(_cpu_fill_rows, tasks)
    t_fill_s = time.time() - t0
    evals_done = len(tasks) * rows_per_task * zint * zint * nz
    t_fill = t_fill_s * (float(L) * (zint * nz) ** 2) / evals_done

    # --- stages 2-4: root, draws, apply (BLAS/LAPACK threads) for sampled l on matrices of the workload's size.
    # LAPACK/BLAS time does not depend on the matrix values: a synthetic covariance of the same conditioning
    # class (21cm: positive definite -> Cholesky branch; foregrounds: the SCK closed form -> eigh branch at >= 1024 ch).
    rng = np.random.default_rng(0)
    if wp["model"] == "21cm":
        x = np.arange(nz, dtype=np.float64)
        cm0 = np.exp(-np.abs(x[:, None] - x[None, :]) / 3.0)
    else:
        cm0 = _CPU["aps"](np.array([100.0]), freq[:, None], freq[None, :])
    n_r = 3 if nz >= 1024 else 8
    lr = np.unique(np.linspace(1, lmax, n_r).astype(int))
    t_root_s = t_draw_s = t_apply_s = 0.0
    for l in lr:
        t0 = time.time()
        cm = cm0 + np.identity(nz) * cm0.diagonal().max() * 1e-14
        tr = onp.matrix_root_manynull(cm, truncate=False)
        t1 = time.time()
        g = onp.complex_std_normal((nz, l + 1), rng=rng)
        t2 = time.time()
        _ = np.dot(tr, g)
        t3 = time.time()
        t_root_s += t1 - t0
        t_draw_s += t2 - t1
        t_apply_s += t3 - t2
    wsum = float(np.sum(lr + 1))
    wall = L * (L + 1) / 2.0
    t_root = t_root_s * L / len(lr)
    t_draw = t_draw_s * wall / wsum
    t_apply = t_apply_s * wall / wsum

    # --- stage 5: inverse SHT (restatement of healpy.alm2map): Legendre over sampled m for a block of channels,
    # phase synthesis over sampled rings
    ch_leg = min(nz, 64)
    n_m = max(8, int(2 * ncores * budget))
    ms = np.unique(np.linspace(0, lmax, n_m).astype(int))
    t0 = time.time()
    pool.map(_cpu_leg_one, [(int(m), lmax, ch_leg, 7 + int(m)) for m in ms])
    t_leg_s = time.time() - t0
    t_leg = t_leg_s * (wall / float(np.sum(lmax - ms + 1))) * (nz / float(ch_leg))
    nring = 4 * nside - 1
    ch_ph = min(nz, 32)
    rs = np.unique(np.linspace(0, nring - 1, 16).astype(int))
    Fm = (rng.standard_normal((L, len(rs), ch_ph)) + 1j * rng.standard_normal((L, len(rs), ch_ph)))
    out = np.empty((ch_ph, npix))
    t0 = time.time()
    osht._rings_from_phase(Fm, _CPU["geom"], rs, out)
    t_ph_s = time.time() - t0
    pix_done = float(np.sum(_CPU["geom"]["nph"][rs]))
    t_phase = t_ph_s * (npix / pix_done) * (nz / float(ch_ph))

    stages = {"cl_fill_s": t_fill, "root_s": t_root, "draws_s": t_draw, "apply_s": t_apply, "sht_legendre_s": t_leg,
              "sht_phase_s": t_phase}
    total = sum(stages.values())
    sample = ("oracle port (numpy/scipy), %d-process fork pool + BLAS threads; per step: C_l fill of %d l x %d channel rows x "
              "all %d columns (%.3g of %.3g evaluations, extrapolated by evaluations); root/draw/apply for %d l on a synthetic "
              "covariance of the workload's size (extrapolated x l-count / x sum(l+1)); SHT Legendre for %d of %d m x %d of %d "
              "channels and phase synthesis for %d of %d rings x %d channels (extrapolated by work); the SHT leg is the "
              "restatement, not healpy; sample wall %.1f s"
              % (ncores, len(ls), len(starts) * rows_per_task, nz, evals_done, float(L) * (zint * nz) ** 2, len(lr), len(ms), L,
                 ch_leg, nz, len(rs), nring, ch_ph, t_fill_s + t_root_s + t_draw_s + t_apply_s + t_leg_s + t_ph_s))
    return total, stages, sample


def run_reference(args):
    """``--impl reference``: rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp

    wp = workload_params(args.workload)
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ncores = min(ncores, 32)
    try:   # torchrun exports OMP_NUM_THREADS=1 to its workers: give the BLAS legs all host cores anyway
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=ncores)
    except Exception:
        pass
    oneoff = cpu_reference_setup(wp)
    pool = mp.get_context("fork").Pool(ncores)
    try:
        for _ in range(args.warmup):
            cpu_reference_step(wp, pool, ncores)
        totals, stages, sample = [], None, ""
        t_start = time.time()
        for _ in range(args.steps):
            tot, stages, sample = cpu_reference_step(wp, pool, ncores)
            totals.append(tot)
        wall = time.time() - t_start
    finally:
        pool.close()
        pool.join()
    total = float(np.median(totals))
    voxels = wp["npix"] * wp["nchan"]
    value = voxels / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "nside": wp["nside"], "channels": wp["nchan"],
                   "lmax": wp["lmax"], "zromb": wp["zromb"],
                   "note": "value = voxels / extrapolated whole-workload CPU seconds (%.1f s); each step times a bounded "
                           "sample; one-off P(k) table build %.1f s excluded (as in the GPU arm)" % (total, oneoff)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample,
                         "stages_extrapolated_s": stages},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# =============================================================================== clocks
class ClockSampler(object):
    """SM clock + throttle-reason sampler running during the timed region (NVML from a thread,
    one sample every few ms -- the C2 timed region is only ~0.2 s, too short for `nvidia-smi -lms`;
    falls back to the B200_PROFILING.md nvidia-smi line if NVML is unavailable)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []          # (t, sm_mhz, power_w, reasons bitmask)
        self.stop_flag = False
        self.thr = None
        self.nvml = None
        self.smax = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except Exception:
                return self.gpu
        return self.gpu

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thr = threading.Thread(target=self._loop, daemon=True)
            self.thr.start()
        except Exception:
            self.nvml = None

    def _loop(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken inside [t0, t1] (perf_counter seconds; default: all)."""
        self.stop_flag = True
        if self.thr is not None:
            self.thr.join(timeout=2)
        if self.nvml is None:
            return self._smi_once()
        sel = [x for x in self.samples if (t0 is None or x[0] >= t0) and (t1 is None or x[0] <= t1)]
        if not sel:
            sel = self.samples[-3:]
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": ["no samples"]}
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                0x80: "hw_power_brake_slowdown"}
        reasons = set()
        for x in sel:
            for b, nm in bits.items():
                if x[3] & b:
                    reasons.add(nm)
        return {"sm_mhz": float(np.median([x[1] for x in sel])), "sm_max_mhz": self.smax,
                "power_w_max": float(max(x[2] for x in sel)), "samples": len(sel), "reasons": sorted(reasons),
                "source": "NVML, sampled every ~4 ms inside the timed region"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q,
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().splitlines()[0].split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = [nm for nm, v in zip(names, f[5:9]) if v.lower().startswith("active")]
            return {"sm_mhz": float(f[1]), "sm_max_mhz": float(f[2]), "power_w_max": float(f[3]), "samples": 1,
                    "reasons": reasons, "source": "nvidia-smi after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}


# =============================================================================== GPU arm
def _make_model(wp, torch):
    """(model, one-off table build seconds)."""
    if wp["model"] == "21cm":
        from cora_b200 import corr21cm

        model = corr21cm.Corr21cm()
        t0 = time.time()
        model.table()
        torch.cuda.synchronize()
        table_s = time.time() - t0
    else:
        from cora_b200 import galaxy

        model = galaxy.FullSkySynchrotron()
        table_s = 0.0
    model.nside = wp["nside"]
    model.frequencies = wp["freq"]
    model.oversample = wp["zromb"]
    return model, table_s


def _measure_resident(sh, out, steps, warmup, barrier, lib, sampler=None):
    """W warm-up steps, then K steps between barrier + synchronize, CUDA events on the launching stream.
    Returns (ms for K steps, launches, {kernel kind: (ms, count)}, (t0, t1) perf_counter of the timed region)."""
    import torch

    for i in range(warmup):
        sh.step(seed=1000 + i, out=out)
    barrier()
    lib.cora_b200_timing_enable(1)
    n0 = lib.cora_b200_launch_count()
    if sampler is not None:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tc0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        sh.step(seed=i, out=out)
    e1.record()
    barrier()
    tc1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = lib.cora_b200_launch_count() - n0
    nk = lib.cora_b200_timing_kinds()
    kms = (ctypes_double * nk)()
    kcnt = (ctypes_ll * nk)()
    lib.cora_b200_timing_read(kms, kcnt, nk)
    lib.cora_b200_timing_enable(0)
    kernels = {lib.cora_b200_timing_name(i).decode(): (kms[i], kcnt[i]) for i in range(nk)}
    return ms, launches, kernels, (tc0, tc1)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from cora_b200 import _dev, _lib, build
    from cora_b200 import dist as cdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU arm has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_bound = _dev.bind_host_to_gpu(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0 and build.needs_build():
        build.build()
    if world > 1:
        dist.barrier()
    lib = _lib.load()

    wp = workload_params(args.workload)
    nside, nchan, lmax, npix = wp["nside"], wp["nchan"], wp["lmax"], wp["npix"]
    model, table_s = _make_model(wp, torch)
    pol = wp["model"] == "gaussianfg_pol"
    npol = 4 if pol else 1
    sht_equiv = 5.0 if pol else 1.0    # T scalar + the (E,B)->(Q,U) pair = 4 scalar-equivalents; V is identically 0

    if pol:
        sh = cdist.ShardedPolSky(nside, wp["freq"], lmax=lmax, zromb=wp["zromb"], rank=rank, size=world)
        sh.exchange = "p2p"
        out = torch.zeros((sh.cb, 4, npix), dtype=torch.float64, device="cuda")
    else:
        sh = cdist.ShardedSky(model, nside, wp["freq"], lmax=lmax, zromb=wp["zromb"], rank=rank, size=world)
        out = torch.empty((sh.cb, npix), dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return float(x)
        tt = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- N > 1: the sharded run must reproduce the single-GPU path (same seed) for this rank's channel block
    parity = None
    parity_note = None
    if world > 1:
        if pol:
            parity_note = "not run: the single-GPU polarised path does not fit beside the sharded buffers"
        else:
            need = 8.0 * (lmax + 1) * nchan * nchan * 2.3 + 16.0 * (lmax + 1) * (lmax + 2) / 2 * nchan + 8.0 * sh.cb * npix + (8 << 30)
            free = torch.cuda.mem_get_info()[0]
            if need > 0.8 * free:
                parity_note = "not run: the single-GPU path needs %.0f GB, %.0f GB free" % (need / 1e9, free / 1e9)
            else:
                sky = sh.step(seed=4242, out=out)
                lo = int(sh.plan.chan_lo[rank])
                ref = cdist.single_gpu_block(model, nside, wp["freq"], lmax, wp["zromb"], 4242, lo, lo + sh.cb)
                err = float((sky - ref).abs().max() / ref.abs().max())
                del ref
                torch.cuda.empty_cache()
                parity = allmax(err)
                if not parity <= 1e-12:
                    raise SystemExit("bench.py: sharded maps differ from the single-GPU path: max rel diff %.3e > 1e-12" % parity)
        barrier()

    # ---- device-resident throughput ("value")
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, kernels, (tc0, tc1) = _measure_resident(sh, out, args.steps, args.warmup, barrier, lib, sampler)
    clocks = sampler.stop(tc0, tc1) if rank == 0 else None
    ms = allmax(ms)
    if world > 1:
        lt = torch.tensor([launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    voxels = float(npix) * nchan * npol
    value = voxels * args.steps / (ms * 1e-3)

    # columns the apply kernel really multiplies per l (it skips the zero triangle of Cholesky roots and the
    # leading all-zero columns of eigen-branch roots): read back from the root stage's flags
    nl_local, cb_local = sh.nl, sh.cb
    l_local = np.asarray(sh.l_list, dtype=np.float64)
    try:
        used = sh._buf["root"][1].cpu().numpy().reshape(-1, nl_local)          # [blocks, nl]
        cols = np.where(used == 0, (nchan + 1) / 2.0, nchan - (used - 1.0))
        fields = np.array([1.0, 2.0])[:used.shape[0]] if pol else np.ones(used.shape[0])   # E and B share the polarised root
        apply_flops = float(np.sum(fields[:, None] * 4.0 * nchan * cols * (l_local[None, :] + 1.0)))
        n_eigh = int((used > 0).sum())
    except Exception:
        apply_flops, n_eigh = None, None

    # ---- end to end through the public API, host buffers in / host maps out ("e2e")
    exchange_mode = sh.exchange
    if world > 1 and sh.exchange == "p2p":
        sh.peers.check()             # (non-fatal mode) a timed-out peer barrier would have produced garbage
    if world == 1 and not pol:
        del sh                       # the resident-path buffers go back to the allocator first
        sh = None
    del out
    torch.cuda.empty_cache()
    e2e_steps = max(1, min(args.steps, 5))
    _dev.traffic["h2d"] = _dev.traffic["d2h"] = 0

    def e2e_once(seed):
        if world == 1 and not pol:
            np.random.seed(seed)
            return model.getsky()  # Sky3d.getsky(): clarray + mkfullsky -> numpy float64[nfreq, npix]
        if not pol:
            return sh.getsky(seed=seed)          # ShardedSky.getsky(): this rank's channels -> numpy (pinned, persistent)
        sky = sh.step(seed=seed)
        sky = sky.reshape(sky.shape[0], -1)
        if sky.numel() * 8 > (4 << 30):          # too large to hold pinned: stream it through a staging ring
            _dev.stream_to_host(sky)
            return sky[:, :0].cpu().numpy().reshape(sky.shape[0], 0)
        return _dev.to_host(sky)   # this rank's channels -> pinned host array

    if world == 1 and not pol:
        e2e_api = "%s.getsky() -> numpy" % type(model).__name__
    elif not pol:
        e2e_api = "dist.ShardedSky.getsky() -> numpy (this rank's channels)"
    else:
        e2e_api = "dist.ShardedPolSky.step() -> host"
    e2e_error = None
    e2e_each = []
    try:
        e2e_once(99)  # warm the pinned-buffer cache (two passes: the first one allocates the host block,
        e2e_once(98)  # the second confirms the allocator hands the same block back)
        _dev.traffic["h2d"] = _dev.traffic["d2h"] = 0
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            ts = time.perf_counter()
            res = e2e_once(i)
            assert res.shape[0] == (cb_local if world > 1 else nchan) and res.shape[1] in (npix * npol, 0)
            del res
            e2e_each.append(round(1e3 * (time.perf_counter() - ts), 2))
        barrier()
        e2e_s = time.perf_counter() - t0
    except (torch.OutOfMemoryError, RuntimeError) as exc:   # only the oversized extra workloads get here
        if args.workload in ("c2", DEFAULT_WORKLOAD):
            raise
        e2e_error = "%s: %s" % (type(exc).__name__, str(exc)[:120])
        e2e_s = float("inf")
        torch.cuda.empty_cache()
    if world > 1:
        tt = torch.tensor([e2e_s, _dev.traffic["h2d"], _dev.traffic["d2h"]], dtype=torch.float64, device="cuda")
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        e2e_s, h2d, d2h = float(mx[0].item()), float(tt[1].item()), float(tt[2].item())
    else:
        h2d, d2h = _dev.traffic["h2d"], _dev.traffic["d2h"]
    e2e_value = voxels * e2e_steps / e2e_s

    # ---- roofline of the dominant kernel (rank 0's view): the Legendre stage of the inverse SHT
    peak = (ctypes_double * 1)()
    _lib.call("cora_b200_fp64_peak", 50.0, peak, _lib.stream_ptr())
    leg_ms, leg_n = kernels["sht_legendre"]
    flops_per_launch = sht_equiv * sht_flops(nside, lmax, cb_local) * args.steps / max(1, leg_n)
    achieved = flops_per_launch / (leg_ms / max(1, leg_n) * 1e-3) / 1e12 if leg_ms > 0 else 0.0
    step_ms = ms / args.steps
    traffic = None   # DRAM bytes per launch of the roofline kernel from the committed ncu --set full capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload)
        if tj and tj.get("n_gpus") == world:
            traffic = tj["bytes_per_launch"]
    except Exception:
        traffic = None
    stage_share = {k: round(v[0] / args.steps, 4) for k, v in kernels.items() if v[1]}

    # ---- per-stage roofline: algorithmic work of THIS rank's share / its event-timed stage time
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6500.0, "fallback (B200_PROFILING.md)"
    L = lmax + 1
    nalm = L * (L + 1) / 2.0
    nring = 4 * nside - 1
    sum_l1 = float(np.sum(l_local + 1.0))   # sum over local l of (l+1)

    def _stage(name, work, unit_scale, peak, bound, what):
        t_ms = stage_share.get(name)
        if not t_ms or work is None:
            return None
        ach = work / (t_ms * 1e-3) / unit_scale
        return {"ms": t_ms, "bound": bound, "achieved": round(ach, 3), "peak": round(peak, 1), "frac": round(ach / peak, 4),
                "unit": "TFLOP/s" if unit_scale == 1e12 else "GB/s", "work": what}

    pk = float(peak[0])
    zi = 2 ** wp["zromb"] + 1
    stage_roofline = {
        "sht_legendre": _stage("sht_legendre", sht_equiv * sht_flops(nside, lmax, cb_local), 1e12, pk, "tensor(fp64 DMMA)",
                               "4*ceil(nring/2)*nalm*channels flop"),
        "sht_phase": _stage("sht_phase", (3 if pol else 1) * cb_local * (16.0 * nring * L + 8.0 * npix), 1e9, hbm_peak,
                            "hbm floor (measured: FP64 issue-bound FFT butterflies)",
                            "16*nring*L + 8*npix bytes per channel (F read + map written)"),
        "apply": _stage("apply", apply_flops, 1e12, pk, "tensor(fp64 DMMA)",
                        "4*nz*cols(l)*(l+1) real flop per l (complex draws, real root); cols(l) = the columns the kernel "
                        "multiplies: (nz+1)/2 for a triangular Cholesky root, the retained columns of an eigen-branch root "
                        "(%s of %d local roots are eigen-branch)" % (n_eigh, nl_local * (2 if pol else 1))),
        "cholesky": _stage("cholesky", (2 if pol else 1) * nl_local * nchan ** 3 / 3.0, 1e12, pk, "tensor(fp64 DMMA) / latency",
                           "nz^3/3 flop per l"),
        "draw": _stage("draw", (3 if pol else 1) * 16.0 * nchan * sum_l1, 1e9, hbm_peak, "hbm (measured: FP64 ALU-bound Box-Muller)",
                       "16*nz*(l+1) bytes written per l"),
        "cl_fill": _stage("cl_fill", (2 if pol else 1) * 8.0 * L * nchan * nchan / world, 1e9, hbm_peak, "L1/L2 gather (hbm = output-write floor)",
                          "8*L*nz^2 output bytes; %.3g evaluations of the 2-D interpolant" % (L * (zi * nchan) ** 2 / 2.0 / world)),
    }
    stage_roofline = {k: v for k, v in stage_roofline.items() if v}
    if sh is not None and exchange_mode == "p2p" and (world > 1 or pol):
        sh.peers.check()
        sh.peers.close()
    sh = None
    torch.cuda.empty_cache()

    # ---- secondary: BASELINE.json's single-GPU configuration (C2) measured in the same run, device-resident
    secondary = None
    if args.workload == DEFAULT_WORKLOAD and not args.no_secondary:
        wp2 = workload_params("c2")
        model2, _ = _make_model(wp2, torch)
        sh2 = cdist.ShardedSky(model2, wp2["nside"], wp2["freq"], lmax=wp2["lmax"], zromb=wp2["zromb"], rank=rank, size=world)
        out2 = torch.empty((sh2.cb, wp2["npix"]), dtype=torch.float64, device="cuda")
        ms2, _, k2, _ = _measure_resident(sh2, out2, max(args.steps, 10), args.warmup, barrier, lib)
        ms2 = allmax(ms2) / max(args.steps, 10)
        if sh2.exchange == "p2p" and world > 1:
            sh2.peers.check()
            sh2.peers.close()
        secondary = {"workload": WORKLOAD_TEXT["c2"], "ms_per_step": ms2,
                     "value": float(wp2["npix"]) * wp2["nchan"] / (ms2 * 1e-3), "unit": UNIT, "steps": max(args.steps, 10),
                     "stage_ms_per_step": {k: round(v[0] / max(args.steps, 10), 4) for k, v in k2.items() if v[1]}}
        del sh2, out2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and not pol:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup",
                                "0", "--workload", args.workload], capture_output=True, text=True, timeout=900,
                               preexec_fn=(lambda: os.sched_setaffinity(0, all_cpus)) if all_cpus else None)  # all host cores
            ref = json.loads(p.stdout.strip().splitlines()[-1])
            cpu_baseline = ref["cpu_baseline"]
        except Exception as exc:  # the GPU numbers stand on their own
            cpu_baseline = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: %r" % (exc,)}

    leg_kernel = ("sht_legendre_kernel<2> + sht_legendre_ws_kernel (FP64 DMMA)" if pol else
                  "sht_legendre_ws_kernel (FP64 DMMA, warp-specialised, TMA-fed)")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "nside": nside, "channels": nchan, "lmax": lmax,
                   "zromb": wp["zromb"], "parallelism": "l-sharded root/apply + channel-sharded SHT x%d" % world,
                   "exchange": ("none (1 GPU)" if world == 1 else
                                ("fused: fill/apply kernels store into peer HBM over NVLink + flag barrier" if exchange_mode == "p2p"
                                 else "NCCL all_to_all_single")),
                   "l2": "per-step working set (C_l %.0f MB, alm %.0f MB, maps %.0f MB per GPU) exceeds the 126 MB L2; no flush needed"
                         % (8e-6 * nl_local * nchan * nchan, 16e-6 * (lmax + 1) * (lmax + 2) / 2 * cb_local, 8e-6 * cb_local * npix),
                   "one_off_table_build_s": round(table_s, 3),
                   "stage_ms_per_step": stage_share,
                   "stage_roofline": stage_roofline, "hbm_peak_source": hbm_src},
        "e2e": {"value": None, "unit": UNIT, "error": e2e_error} if e2e_error else {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / e2e_steps, "d2h_bytes_per_step": d2h / e2e_steps,
                "steps": e2e_steps, "ms_each": e2e_each, "host_bound_to_gpu_numa_node": numa_bound, "api": e2e_api},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": leg_kernel, "achieved": achieved,
                     "peak": float(peak[0]), "unit": "TFLOP/s", "frac": achieved / float(peak[0]) if peak[0] else None,
                     "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, profiles/ncu_traffic.json)",
                     "algorithmic_bytes_per_launch": 16.0 * cb_local * ((lmax + 1) * (lmax + 2) / 2 + (4 * nside - 1) * (lmax + 1)),
                     "peak_source": "FP64 DMMA peak measured in this run by cora_b200_fp64_peak (MEASURED_PEAKS.json "
                                    "has no FP64 entry; 37.1 TFLOP/s recorded in profiles/microbench/)",
                     "flops_per_launch": flops_per_launch, "launch_ms": leg_ms / max(1, leg_n),
                     "note": "frac is algorithmic flop (incl. the polar (ring, m, l) octets the kernel skips) over the DMMA peak"},
        "cpu_baseline": cpu_baseline,
    }
    if world > 1:
        line["parity_vs_single_gpu"] = parity
        if parity_note:
            line["parity_vs_single_gpu_note"] = parity_note
    if secondary:
        line["secondary"] = secondary
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


import ctypes  # noqa: E402

ctypes_double = ctypes.c_double
ctypes_ll = ctypes.c_longlong


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-2 block measured next to the default workload")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself one rank per GPU
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                                  str(args.gpus), "--master-addr", "127.0.0.1", "--master-port", "29533",
                                  os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
