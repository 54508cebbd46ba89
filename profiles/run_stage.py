#!/usr/bin/env python
"""Run ONE stage of the path a few times (for ncu captures / quick timings under gpurun).

    python profiles/run_stage.py fill|root|sht [--workload c3] [--reps 3]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cora_b200 import dist as cdist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("stage", choices=["fill", "root", "alm", "sht", "step"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    wp = bench.workload_params(a.workload)
    model, _ = bench._make_model(wp, torch)
    sh = cdist.ShardedSky(model, wp["nside"], wp["freq"], lmax=wp["lmax"], zromb=wp["zromb"], rank=0, size=1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    cla = sh.fill()
    torch.cuda.synchronize()
    if a.stage == "step":
        out = torch.empty((sh.cb, sh.npix), dtype=torch.float64, device="cuda")
        fn = lambda: sh.step(seed=3, out=out)
    elif a.stage == "fill":
        fn = sh.fill
    elif a.stage in ("root", "alm"):
        fn = lambda: sh.alm_local(cla, seed=1)
    else:
        panel = sh.alm_local(cla, seed=1)
        fn = lambda: sh.synthesize(panel)
    fn()
    torch.cuda.synchronize()
    ev[0].record()
    marks = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        fn()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((e, time.perf_counter() - t0))
    ev[1].record()
    torch.cuda.synchronize()
    prev, each = ev[0], []
    for e, _ in marks:
        each.append(round(prev.elapsed_time(e), 1))
        prev = e
    print("%s %s: %.3f ms per call; each %s; host enqueue ms %s" % (a.stage, a.workload, ev[0].elapsed_time(ev[1]) / a.reps, each,
                                                                   [round(1e3 * h, 1) for _, h in marks]))


if __name__ == "__main__":
    main()
