#!/usr/bin/env python
"""Latency / throughput of the batched Cholesky (cora_b200_root_batched) vs batch size."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from cora_b200 import _lib, nputil

lib = _lib.load()
nz = int(os.environ.get("NZ", 256))
rng = np.random.default_rng(0)
a = rng.standard_normal((nz, 2 * nz))
spd = torch.from_numpy(a @ a.T / (2 * nz) + np.eye(nz)).cuda()
out = {}
for nl in (1, 37, 74, 148, 296, 444, 592, 768):
    cl = spd.unsqueeze(0).repeat(nl, 1, 1).contiguous()
    ws = nputil.root_workspace(nl, nz, max_eigh=4)
    outb = (torch.empty_like(cl), torch.empty(nl, dtype=torch.int32, device="cuda"), torch.empty(nl, dtype=torch.int32, device="cuda"))
    for _ in range(2):
        nputil.root_batched_device(cl, 1e-14, 1e-16, out=outb, ws=ws)
    torch.cuda.synchronize()
    lib.cora_b200_timing_enable(1)
    reps = 5
    for _ in range(reps):
        nputil.root_batched_device(cl, 1e-14, 1e-16, out=outb, ws=ws)
    torch.cuda.synchronize()
    nk = lib.cora_b200_timing_kinds()
    ms = (ctypes.c_double * nk)(); cnt = (ctypes.c_longlong * nk)()
    lib.cora_b200_timing_read(ms, cnt, nk)
    lib.cora_b200_timing_enable(0)
    k = {lib.cora_b200_timing_name(i).decode(): ms[i] / reps for i in range(nk) if cnt[i]}
    out[nl] = {"cholesky_ms": round(k["cholesky"], 4), "prepare_ms": round(k.get("root_prepare", 0.0), 4),
               "gflops": round(nl * nz**3 / 3.0 / (k["cholesky"] * 1e-3) / 1e9, 1)}
print(json.dumps({"nz": nz, "by_batch": out}))
