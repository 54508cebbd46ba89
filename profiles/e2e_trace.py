#!/usr/bin/env python
"""Timeline of the public end-to-end call (Sky3d.getsky -> numpy) at a bench workload: wall time per call and the
kernel trace (cora_b200_timing_trace) of the last call, to see where the device idles behind the host.

    python profiles/e2e_trace.py [--workload c3] [--calls 4]
"""
import argparse
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cora_b200 import _dev, _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--calls", type=int, default=4)
    a = ap.parse_args()
    lib = _lib.load()
    _dev.bind_host_to_gpu(0)
    wp = bench.workload_params(a.workload)
    model, _ = bench._make_model(wp, torch)
    for i in range(a.calls):
        np.random.seed(i)
        torch.cuda.synchronize()
        if i == a.calls - 1:
            lib.cora_b200_timing_enable(1)
        t0 = time.perf_counter()
        sky = model.getsky()
        t1 = time.perf_counter()
        print("call %d: %.1f ms  (free %.1f GB)" % (i, 1e3 * (t1 - t0), torch.cuda.mem_get_info()[0] / 1e9), flush=True)
        del sky
    cap = 4096
    ids = (ctypes.c_int * cap)()
    st = (ctypes.c_double * cap)()
    du = (ctypes.c_double * cap)()
    n = -lib.cora_b200_timing_trace(ids, st, du, cap)
    prev_end = 0.0
    for i in range(n):
        gap = st[i] - prev_end
        print("%4d %-14s start %9.3f  dur %9.3f  gap %8.3f%s" % (i, lib.cora_b200_timing_name(ids[i]).decode(), st[i], du[i], gap,
                                                                 "   <-- gap" if gap > 1.0 else ""))
        prev_end = st[i] + du[i]


if __name__ == "__main__":
    main()
