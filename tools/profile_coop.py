"""Dev tool (GPU): one launch of the cooperative MCMC kernel on the 5x5 network (8192 chains x 20 iterations) for ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.coop_perf import network, timed  # noqa: E402

dm, center, P = network()
C, its = int(os.environ.get("CHAINS", 8192)), int(os.environ.get("ITS", 20))
rng = np.random.default_rng(2)
starts = torch.from_numpy(center * np.exp(0.02 * rng.standard_normal((C, P)))).cuda()
t, r = timed(lambda: dm.mcmc(starts, nits=its, rng_mode="philox", seed=1, device_buffers=True, keep_samples=False))
print("chains", C, "its", its, "seconds", t, "k chain-steps/s", C * (its - 1) / t / 1e3, dm.kernel_info("mcmc_coop"))
