import os, sys
os.environ["ODL_KERNEL_DEFINES"] = "-DODL_DEBUG_TRAJ=1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odelib_b200 import workloads
from odelib_b200.engine import DeviceModel
for N in (6, 7):
    rhs, n, P, groups = workloads.nclass(N, spec_only=True)
    dm = DeviceModel(rhs, n, P, groups, device=0, cache_dir=None)
    y0 = np.zeros(n); y0[0] = 5236900.0; y0[-1] = 10981000.0
    center = np.array([0.3, 1.0e-7, 20.0, 2.0, 2.8 * N / 2])
    dm.set_grid(np.linspace(0, 3, 19), y0)
    traj, status, nsteps = dm.trajectory(center[None])
    torch.cuda.synchronize()
    print("N", N, "status", status, "nsteps", nsteps, flush=True)
