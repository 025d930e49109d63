"""Dev tool (GPU): config 3 through the facade's chain runner, with the split main run / re-run.
    python tools/c3_split.py [N=4] [nits=10000]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odelib_b200 import workloads  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nits = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
m, center = workloads.nclass(N, device=0)
dm = m._device()
C, P = 4096, dm.n_param
rng = np.random.default_rng([3, N, 0])
starts = torch.from_numpy(center * np.exp(0.02 * rng.standard_normal((C, P)))).cuda()
seeds = list(range(C))
kw = dict(rng="philox", return_raw=True, keep_samples=False)
m._run_chains(starts, seeds, 60, 30, (), **kw)
orig = dm.mcmc
log = []


def timed(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(*a, **k)
    torch.cuda.synchronize()
    log.append((k.get("solver"), k.get("max_steps"), k.get("explicit_budget", 0), len(a[0]), time.perf_counter() - t0))
    return r


dm.mcmc = timed
t0 = time.perf_counter()
res = m._run_chains(starts, seeds, nits, nits // 2, (), **kw)
torch.cuda.synchronize()
t = time.perf_counter() - t0
print("N=%d nits=%d total %.3f s  %.2f M chain-steps/s  rerun %d  fails %d" % (N, nits, t, C * (nits - 1) / t / 1e6, m._last_rerun,
                                                                             int(np.asarray(res["fail_count"]).sum())))
for rec in log:
    print("  mcmc solver=%s max_steps=%s explicit_budget=%s chains=%d: %.3f s" % rec)
