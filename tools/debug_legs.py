"""Dev tool (GPU): the synthetic workloads of bench.py's config legs, one at a time, with progress on stderr."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odelib_b200 import workloads
which = sys.argv[1]
t0 = time.time()
if which.startswith("n") and which != "net":
    m, center = workloads.nclass(int(which[1:]), device=0)
elif which == "net":
    m, center = workloads.network(device=0)
else:
    m, center = workloads.stiff(device=0)
print("built", which, time.time() - t0, "inits", m.get_inits()[:4], flush=True)
pred = m.integrate(predict_obs=True, as_dataframe=False)
print({k: (float(np.min(v)), float(np.max(v))) for k, v in list(pred.items())[:3]}, "chi", m.get_chi(pred), flush=True)
dm = m._device()
C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nits = int(sys.argv[3]) if len(sys.argv) > 3 else 50
starts = torch.from_numpy(center * np.exp(0.02 * np.random.default_rng(0).standard_normal((C, dm.n_param)))).cuda()
ms = int(sys.argv[4]) if len(sys.argv) > 4 else 500000
for n_ in (10, nits, 4 * nits):
    t0 = time.time()
    res = dm.mcmc(starts, nits=n_, rng_mode="philox", seed=1, device_buffers=True, keep_samples=False, max_steps=ms)
    torch.cuda.synchronize()
    print("mcmc", C, n_, "s", time.time() - t0, "accept", float(res["chain_state"][:, 2].mean()) / (n_ - 1), "fails", int(res["fail_count"].sum()),
          "steps/solve", float(res["step_count"].sum()) / (C * n_),
          "Mchain-steps/s", C * (n_ - 1) / (time.time() - t0) / 1e6, flush=True)
