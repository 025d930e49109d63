"""Dev tool (GPU): chain-steps/s of the MCMC kernel against the prefetching width K and the chain count."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model
import bench
VAR = sys.argv[1] if len(sys.argv) > 1 else ""
if VAR:
    os.environ["ODL_KERNEL_DEFINES"] = VAR
dm, tab = device_model("two_i")
print(VAR, dm.kernel_info("mcmc"))
P = 5
for C in (4096, 16384, 65536):
    rng = np.random.default_rng(1)
    starts = torch.from_numpy(np.array(bench.CENTER["two_i"]) * np.exp(0.05 * rng.standard_normal((C, P)))).cuda()
    nits = 300 if C <= 16384 else 100
    row = []
    for K in (0, 1, 4, 8):
        if C * max(K, 1) > 1 << 21:
            continue
        kw = dict(nits=nits, rng_mode="philox", seed=0, pnum=P, device_buffers=True, keep_samples=False, speculate=K)
        dm.mcmc(starts, **dict(kw, nits=10))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dm.mcmc(starts, **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        row.append((K, round(C * (nits - 1) / dt / 1e6, 1)))
    print("chains", C, "Mchain-steps/s by K:", row, flush=True)
