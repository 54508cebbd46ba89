#!/usr/bin/env python
"""Timeline of the kernels of one resident step (cora_b200_timing_trace): start, duration and the gap
to the previous kernel, plus the host enqueue time of the step.

    python profiles/trace_step.py [--workload c3] [--steps 2]
"""
import argparse
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cora_b200 import _lib  # noqa: E402
from cora_b200 import dist as cdist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    lib = _lib.load()
    wp = bench.workload_params(a.workload)
    model, _ = bench._make_model(wp, torch)
    sh = cdist.ShardedSky(model, wp["nside"], wp["freq"], lmax=wp["lmax"], zromb=wp["zromb"], rank=0, size=1)
    out = torch.empty((sh.cb, sh.npix), dtype=torch.float64, device="cuda")
    for i in range(3):
        sh.step(seed=i, out=out)
    torch.cuda.synchronize()
    print("free after warm-up: %.1f GB; persistent buffers: %s" % (torch.cuda.mem_get_info()[0] / 1e9, {
        k: (round(v.numel() * v.element_size() / 1e9, 2) if hasattr(v, "numel") else "tuple") for k, v in sh._buf.items()}))
    lib.cora_b200_timing_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host = []
    for i in range(a.steps):
        t0 = time.perf_counter()
        sh.step(seed=10 + i, out=out)
        host.append(round(1e3 * (time.perf_counter() - t0), 2))
    e1.record()
    torch.cuda.synchronize()
    cap = 4096
    ids = (ctypes.c_int * cap)()
    st = (ctypes.c_double * cap)()
    du = (ctypes.c_double * cap)()
    n = -lib.cora_b200_timing_trace(ids, st, du, cap)
    print("%d steps: %.2f ms each on the device; host enqueue ms per step %s" % (a.steps, e0.elapsed_time(e1) / a.steps, host))
    prev_end = 0.0
    for i in range(n):
        gap = st[i] - prev_end
        print("%4d %-14s start %9.3f  dur %9.3f  gap %8.3f%s" % (i, lib.cora_b200_timing_name(ids[i]).decode(), st[i], du[i], gap,
                                                                 "   <-- gap" if gap > 0.3 else ""))
        prev_end = st[i] + du[i]
    lib.cora_b200_timing_enable(0)


if __name__ == "__main__":
    main()
