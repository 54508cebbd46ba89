#!/usr/bin/env python
"""Time the polarised block-sharded generator (dist.ShardedPolSky, one rank) and print per-kernel
stage times.  NSIDE / NCHAN from the environment (default 256 / 256)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from cora_b200 import _lib
from cora_b200 import dist as cdist

nside, nchan = int(os.environ.get("NSIDE", 256)), int(os.environ.get("NCHAN", 256))
lib = _lib.load()
freq = np.linspace(800.0, 400.0, nchan, endpoint=False)
sh = cdist.ShardedPolSky(nside, freq, rank=0, size=1)
sh.step(seed=0)
torch.cuda.synchronize()
lib.cora_b200_timing_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = int(os.environ.get("REPS", 3))
e0.record()
for i in range(reps):
    sky = sh.step(seed=1 + i)
e1.record()
torch.cuda.synchronize()
nk = lib.cora_b200_timing_kinds()
ms = (ctypes.c_double * nk)()
cnt = (ctypes.c_longlong * nk)()
lib.cora_b200_timing_read(ms, cnt, nk)
k = {lib.cora_b200_timing_name(i).decode(): round(ms[i] / reps, 3) for i in range(nk) if cnt[i]}
used_all = sh._buf["root"][1]   # int32[2 blocks, nl]: > 0 = eigen branch
used = [int((used_all[b] > 0).sum().item()) for b in range(2)]
print(json.dumps({"shape": {"nside": nside, "nchan": nchan, "lmax": sh.lmax, "npol": 4}, "ms_per_step": e0.elapsed_time(e1) / reps,
                  "voxels_per_s": 4.0 * nchan * sh.npix / (e0.elapsed_time(e1) / reps * 1e-3), "kernels_ms": k,
                  "l_on_eigen_branch": {"T": used[0], "P": used[1], "of": sh.nl}}))
