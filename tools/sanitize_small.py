"""Dev tool (GPU): the smallest run that touches every kernel family, for compute-sanitizer --tool memcheck."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odelib_b200 import demo_models
from tests.helpers import device_model, prior_draws, synthetic_problem
from tests.test_gpu_models import nclass_problem

dm, tab = device_model("two_i")
theta = prior_draws("two_i", 3000, seed=0)
a = dm.sweep(theta, solver="auto", max_steps=200000)                      # ordering, bulk, bdf consumer (host path)
b = dm.sweep(torch.from_numpy(theta).cuda(), solver="auto", max_steps=200000, auto_flags=2)   # concurrent consumer
c = dm.sweep(theta[:200], solver="radau5", max_steps=200000)
d = dm.sweep(theta[:200], solver="ros23", max_steps=200000)
print("sweeps ok", int((a["status"] == 0).sum()), int((b["status"] == 0).sum().item()), int((c["status"] == 0).sum()), int((d["status"] == 0).sum()))
center = np.array([7.475e-09, 1.069e-07, 19.73, 1.934, 2.799])
starts = center * np.exp(0.05 * np.random.default_rng(1).standard_normal((37, 5)))
for spec in (1, 8):
    r = dm.mcmc(starts, nits=30, seed=1, speculate=spec, trace=True)
print("mcmc ok", float(r["accepted"].mean()))
r = dm.mcmc(starts[:8], nits=12, seed=1, solver="bdf")
th_dev = dm.sample_lhs([("lognorm", 3.0, 0.0, 1e-8)] * 2 + [("lognorm", 1.0, 0.0, 20.0), ("lognorm", 2.0, 0.0, 0.1), ("lognorm", 2.0, 0.0, 1.0)], 5000, seed=3)
res = dm.sweep(th_dev, solver="auto", max_steps=200000)
idx, cnt = dm.select_below(res["chi"], 666.0)
g = dm.gather_rows(th_dev, np.arange(min(cnt, 10)), index=idx)
print("lhs/select/gather ok", cnt, tuple(g.shape))
traj, st, ns = dm.trajectory(theta[:16])
print("traj ok", traj.shape)
rhs, names, sums, cen, y0, orgs = nclass_problem(10)
dmc, tabc = synthetic_problem(rhs, names, sums, cen, y0, orgs, seed=10)
thc = cen * np.exp(0.05 * np.random.default_rng(2).standard_normal((21, 5)))
s1 = dmc.sweep(thc, max_steps=200000)
s2 = dmc.sweep(thc, solver="auto", max_steps=200000)
m1 = dmc.mcmc(thc[:9], nits=16, seed=2)
m2 = dmc.mcmc(thc[:9], nits=16, seed=2, speculate=-4)
print("coop ok", int((s1["status"] == 0).sum()), int((s2["status"] == 0).sum()), bool(np.array_equal(m1["samples"], m2["samples"])))
