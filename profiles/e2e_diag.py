import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from cora_b200 import corr21cm, _dev, hputil, skysim
_dev.bind_host_to_gpu(0)
cr = corr21cm.Corr21cm(); cr.table()
NS, NC = int(os.environ.get("NSIDE", 256)), int(os.environ.get("NCHAN", 256))
cr.nside, cr.frequencies, cr.oversample = NS, np.linspace(800., 400., NC, endpoint=False), 3
orig_empty = torch.empty
stamps = []
def timed_empty(*a, **k):
    if k.get("pin_memory"):
        t0 = time.perf_counter(); r = orig_empty(*a, **k); stamps.append(("pin_alloc", time.perf_counter() - t0)); return r
    return orig_empty(*a, **k)
torch.empty = timed_empty
for i in range(12):
    stamps.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cl = skysim.clarray(cr.angular_powerspectrum, 3 * NS - 1, cr.frequencies, device_out=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    sky = skysim.mkfullsky(cl, NS, seed=i)
    t2 = time.perf_counter()
    del sky, cl
    t3 = time.perf_counter()
    print("step %d total %.1f ms: clarray %.1f  mkfullsky %.1f  del %.2f  %s" % (i, 1e3*(t3-t0), 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), [(a, round(1e3*b,2)) for a,b in stamps]))
