#!/usr/bin/env python
"""Is the e2e jitter (one call in 3-4 takes +150..400 ms) in the PCIe copy itself?  Times repeated device -> pinned-host
copies of a map-sized tensor (one big copy, and the same bytes in 16 chunks on a side stream), nothing else running."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cora_b200 import _dev  # noqa: E402

_dev.bind_host_to_gpu(0)
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 25.77
n = int(gb * 1e9 / 8)
dev = torch.zeros(n, dtype=torch.float64, device="cuda")
host = torch.empty(n, dtype=torch.float64, pin_memory=True)
host.copy_(dev)
torch.cuda.synchronize()
one, chunked = [], []
for i in range(12):
    t0 = time.perf_counter()
    host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    one.append(round(1e3 * (time.perf_counter() - t0), 1))
side = torch.cuda.Stream()
c = n // 16
for i in range(12):
    t0 = time.perf_counter()
    with torch.cuda.stream(side):
        for k in range(16):
            host[k * c:(k + 1) * c].copy_(dev[k * c:(k + 1) * c], non_blocking=True)
    side.synchronize()
    chunked.append(round(1e3 * (time.perf_counter() - t0), 1))
print("D2H of %.2f GB, ms per copy: single %s" % (gb, one))
print("                          16 chunks %s" % chunked)
print("best %.1f GB/s" % (gb / (min(one) * 1e-3)))
