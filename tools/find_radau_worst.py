import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from tests.helpers import device_model, prior_draws
res={}
for name in ("one_i","two_i"):
    dm, tab = device_model(name)
    theta = prior_draws(name, 262144, seed=0)
    r = dm.sweep(theta, solver="radau5", max_steps=200000)
    o = np.argsort(-r["nsteps"])[:12]
    res[name+"_theta"]=theta[o]; res[name+"_nsteps"]=r["nsteps"][o]; res[name+"_chi"]=r["chi"][o]
    print(name, r["nsteps"][o])
np.savez("gpurun_out/radau_worst.npz", **res)
