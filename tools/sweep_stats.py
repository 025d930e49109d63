"""Dev tool (GPU): step-count distribution and timing of the 1M-set two_i prior sweep per solver mode."""
import json
import sys
import time

import numpy as np
import torch

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
MODEL = sys.argv[2] if len(sys.argv) > 2 else "two_i"
dm, tab = device_model(MODEL)
from tests.helpers import prior_draws
theta = torch.from_numpy(prior_draws(MODEL, n, seed=0)).cuda()
res = {}
for mode, kw in (("auto_default", dict(solver="auto", max_steps=200000)),
                 ("auto_512_1024", dict(solver="auto", max_steps=200000, pass_caps=(512, 1024))),
                 ("auto_512_1536", dict(solver="auto", max_steps=200000, pass_caps=(512, 1536))),
                 ("auto_512_2048", dict(solver="auto", max_steps=200000, pass_caps=(512, 2048))),
                 ("auto_384_1024", dict(solver="auto", max_steps=200000, pass_caps=(384, 1024))),
                 ("auto_448_0", dict(solver="auto", max_steps=200000, pass_caps=(448, 1))),
                 ("auto_640_0", dict(solver="auto", max_steps=200000, pass_caps=(640, 1))),
                 ("dopri5_cap512", dict(solver="dopri5", max_steps=512))):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = dm.sweep(theta, **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ns = out["nsteps"].cpu().numpy(); st = out["status"].cpu().numpy()
    q = np.percentile(ns, [50, 90, 99, 99.9, 99.99, 100])
    res[mode] = {"seconds": dt, "kernel_ms": dm.last_kernel_ms(), "pass_ms": dm.last_pass_ms(), "solves_per_s": n / dt, "mean_steps": float(ns.mean()),
                 "pct_50_90_99_999_9999_max": q.tolist(), "status_counts": {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
                 "chi_finite": int(np.isfinite(out["chi"].cpu().numpy()).sum())}
    print(mode, json.dumps(res[mode]), flush=True)
json.dump(res, open("gpurun_out/sweep_stats.json", "w"), indent=1)
# worst systems of the last mode: parameters + step counts, to study on the CPU with LSODA
ns = out["nsteps"].cpu().numpy()
order = np.argsort(-ns)[:200]
np.savez("gpurun_out/worst.npz", theta=theta.cpu().numpy()[order], nsteps=ns[order], chi=out["chi"].cpu().numpy()[order])
plain = dm.sweep(theta, solver="dopri5", stiff_check=True, max_steps=512)
st = plain["status"].cpu().numpy()
print("pass0: stiff", int((st == 4).sum()), "maxsteps", int((st == 1).sum()), "of", n)
rad = dm.sweep(theta[torch.from_numpy(np.flatnonzero(st != 0)).cuda()], solver="radau5", max_steps=200000)
rn = rad["nsteps"].cpu().numpy()
print("radau on deferred: n", rn.size, "steps pct 50/90/99/max", np.percentile(rn, [50, 90, 99, 100]).tolist(), "ms", dm.last_kernel_ms())
