"""Dev tool (GPU): chain-steps/s of the MCMC kernels against the mapping and prefetching width, per chain count.
speculate: 0 = automatic thread-per-system width, K >= 1 thread-per-system with K lanes per chain, -K cooperative kernel
with K groups of 4 lanes per chain."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model
import bench
MODEL = sys.argv[1] if len(sys.argv) > 1 else "two_i"
dm, tab = device_model(MODEL)
P = dm.n_param
center = {"two_i": bench.CENTER["two_i"], "one_i": [1.238e-08, 3.550e-08, 19.40, 1.835], "zero_i": [1.36e-8, 1.35e-8, 19.44]}[MODEL]
for C in (1, 32, 256, 1024, 4096, 16384, 65536):
    rng = np.random.default_rng(1)
    starts = torch.from_numpy(np.array(center) * np.exp(0.05 * rng.standard_normal((C, P)))).cuda()
    nits = 1000 if C <= 32 else (300 if C <= 16384 else 100)
    row = []
    for K in (0, 1, 8, -1, -2, -4, -8):
        if C * abs(K if K else 1) * (4 if K < 0 else 1) > 1 << 21:
            continue
        kw = dict(nits=nits, rng_mode="philox", seed=0, pnum=P, device_buffers=True, keep_samples=False, speculate=K)
        dm.mcmc(starts, **dict(kw, nits=10))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dm.mcmc(starts, **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        row.append((K, round(C * (nits - 1) / dt / 1e6, 3)))
    print(MODEL, "chains", C, "Mchain-steps/s by speculate:", row, flush=True)
