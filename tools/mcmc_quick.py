import os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from tests.helpers import device_model
import bench
dm, tab = device_model("two_i")
P = 5
C = 4096
rng = np.random.default_rng([1, 0])
starts = torch.from_numpy(np.array(bench.CENTER["two_i"]) * np.exp(0.05 * rng.standard_normal((C, P)))).cuda()
for keep in (False, True):
    for nits in (300, 500):
        kw = dict(nits=nits, rng_mode="philox", seed=0, pnum=P, device_buffers=True, keep_samples=keep)
        dm.mcmc(starts, **dict(kw, nits=20))
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = dm.mcmc(starts, **kw)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print("keep", keep, "nits", nits, "Mchain-steps/s %.1f" % (C * (nits - 1) / dt / 1e6), "kernel_ms %.2f" % dm.last_kernel_ms(), flush=True)
