"""Dev tool (GPU): where the wall time of ModelFramework.MCMC(chain_inits=4096 from a 1M survey) goes."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_facade_host import make_model  # noqa: E402

m = make_model("two_i")
np.random.seed(0)
m.fit_survey(samples=1000)
n = 1 << 20
for rep in range(2):
    np.random.seed(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = m.MCMC(chain_inits=4096, iterations_per_chain=200, fitsurvey_samples=n, sd_fitdistance=6.0, print_report=False,
               posterior="summary", rng=(sys.argv[1] if len(sys.argv) > 1 else "auto"))
    torch.cuda.synchronize(); print("MCMC wall %.3f s, rerun %d, solver %s" % (time.perf_counter() - t0, m._last_rerun, m._last_solver))
np.random.seed(1)
pr = cProfile.Profile()
pr.enable()
m.MCMC(chain_inits=4096, iterations_per_chain=200, fitsurvey_samples=n, sd_fitdistance=6.0, print_report=False,
       posterior="summary", rng=(sys.argv[1] if len(sys.argv) > 1 else "auto"))
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(22)
