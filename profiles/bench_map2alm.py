#!/usr/bin/env python
"""Time the forward SHT (cora_b200_map2alm) at the C2 shape: nside 256, lmax 767, 256 channels.
Prints one JSON line: quadrature pass (iter=0) and the reference's iter=2 call, with the
Legendre-adjoint kernel's TFLOP/s against the DMMA peak measured in the same run."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from cora_b200 import _lib, hputil

nside, nchan = int(os.environ.get("NSIDE", 256)), int(os.environ.get("NCHAN", 256))
lmax = 3 * nside - 1
lib = _lib.load()
maps = torch.randn((nchan, 12 * nside**2), dtype=torch.float64, device="cuda")
panel = hputil.map2alm_device(maps, nside, lmax, iter=0)
torch.cuda.synchronize()
out = {}
for it in (0, 2):
    lib.cora_b200_timing_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        hputil.map2alm_device(maps, nside, lmax, iter=it, panel=panel)
    e1.record()
    torch.cuda.synchronize()
    nk = lib.cora_b200_timing_kinds()
    ms = (ctypes.c_double * nk)()
    cnt = (ctypes.c_longlong * nk)()
    lib.cora_b200_timing_read(ms, cnt, nk)
    lib.cora_b200_timing_enable(0)
    k = {lib.cora_b200_timing_name(i).decode(): (ms[i] / reps, cnt[i] // reps) for i in range(nk) if cnt[i]}
    out["iter%d" % it] = {"ms": e0.elapsed_time(e1) / reps, "kernels_ms": {a: round(b[0], 3) for a, b in k.items()}}
peak = (ctypes.c_double * 1)()
_lib.call("cora_b200_fp64_peak", 50.0, peak, _lib.stream_ptr())
L = lmax + 1
flops = 4.0 * (2 * nside) * (L * (L + 1) / 2.0) * nchan
leg = out["iter0"]["kernels_ms"]["sht_legendre"]
out["legendre_adj_tflops"] = flops / (leg * 1e-3) / 1e12
out["dmma_peak_tflops"] = peak[0]
out["frac"] = out["legendre_adj_tflops"] / peak[0]
out["shape"] = {"nside": nside, "lmax": lmax, "nchan": nchan}
# polarised analysis (Q, U) -> (aE, aB), quadrature pass
lib.cora_b200_timing_enable(0)
Q, U = maps[: nchan // 2].contiguous(), maps[nchan // 2 :].contiguous()
hputil.map2alm_spin2_device(Q, U, nside, lmax, iter=0)
torch.cuda.synchronize()
e0.record()
hputil.map2alm_spin2_device(Q, U, nside, lmax, iter=0)
e1.record()
torch.cuda.synchronize()
out["spin2_iter0_ms_%dch" % (nchan // 2)] = e0.elapsed_time(e1)
# CPU side by side: the oracle's quadrature pass (numpy, one process) on 2 channels, scaled to nchan
if os.environ.get("CPU", "1") == "1":
    import time
    from oracle import sht as osht
    sub = maps[:2].cpu().numpy()
    t0 = time.time()
    osht.map2alm_adjoint(sub, nside, lmax)
    out["cpu_oracle_iter0_s_extrapolated"] = (time.time() - t0) * nchan / 2.0
    out["cpu_note"] = "oracle/sht.py map2alm_adjoint on 2 channels x nchan/2 (numpy; restatement, not healpy)"
print(json.dumps(out))
