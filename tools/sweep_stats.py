"""Dev tool (GPU): step-count distribution and timing of the 1M-set prior sweep per solver mode / kernel variant."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
MODEL = sys.argv[2] if len(sys.argv) > 2 else "two_i"
VARIANTS = sys.argv[3].split(",") if len(sys.argv) > 3 else [""]
theta = torch.from_numpy(prior_draws(MODEL, n, seed=0)).cuda()
K = dict(solver="auto", max_steps=200000)
MODES = (("auto_default", dict(K)),
         ("auto_noearly", dict(K, early_check_steps=-1)),
         ("auto_early128", dict(K, early_check_steps=128)),
         ("auto_unordered", dict(K, auto_flags=1)),
         ("auto_concurrent", dict(K, auto_flags=2, early_check_steps=-1)),
         ("auto_radau", dict(K, tail_solver="radau5", early_check_steps=-1)),
         ("auto_384", dict(K, pass_caps=384)),
         ("auto_448", dict(K, pass_caps=448)),
         ("auto_640", dict(K, pass_caps=640)),
         ("auto_768", dict(K, pass_caps=768)),
         ("dopri5_cap512", dict(solver="dopri5", max_steps=512)))
res = {}
for variant in VARIANTS:
    if variant:
        os.environ["ODL_KERNEL_DEFINES"] = variant
    dm, tab = device_model(MODEL)
    chis = {}
    for mode, kw in MODES:
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = dm.sweep(theta, **kw)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ns = out["nsteps"].cpu().numpy(); st = out["status"].cpu().numpy(); chis[mode] = out["chi"].cpu().numpy()
        q = np.percentile(ns, [50, 90, 99, 99.9, 100])
        res[variant + mode] = {"kernel_ms": round(dm.last_kernel_ms(), 4), "pass_ms": [round(x, 4) for x in dm.last_pass_ms()],
                               "Msolves_per_s": round(n / dt / 1e6, 2), "mean_steps": round(float(ns.mean()), 2), "pct_50_90_99_999_max": q.tolist(),
                               "status_counts": {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
                               "chi_finite": int(np.isfinite(chis[mode]).sum())}
        print(variant, mode, json.dumps(res[variant + mode]), flush=True)
    a, b = chis["auto_radau"], chis["auto_default"]
    both = np.isfinite(a) & np.isfinite(b)
    rel = np.abs(a - b)[both] / (np.abs(a[both]) + 1e-300)
    print("chi: default vs radau tail: max rel", float(rel.max()), "p99.99", float(np.percentile(rel, 99.99)), "n >1e-6", int((rel > 1e-6).sum()),
          "n >1e-4", int((rel > 1e-4).sum()))
    plain = dm.sweep(theta, solver="dopri5", stiff_check=True, max_steps=512)
    hard = theta[plain["status"] != 0].contiguous()
    for solver in ("radau5", "bdf"):
        for rep in range(2):
            r = dm.sweep(hard, solver=solver, max_steps=200000)
        rn = r["nsteps"].cpu().numpy()
        print(variant, solver, "on deferred: n", rn.size, "steps pct 50/90/99/max", np.percentile(rn, [50, 90, 99, 100]).tolist(), "ms", dm.last_kernel_ms())
    dm.close()
json.dump(res, open("gpurun_out/sweep_stats.json", "w"), indent=1)
