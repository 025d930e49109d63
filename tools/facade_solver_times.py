"""Dev tool (GPU): ModelFramework.MCMC(4096 chains from a 1M survey) per stepper choice / DOPRI5 step budget."""
import os
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
its = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for solver, budget in (("auto", 4096), ("auto", 2048), ("auto", 1024), ("auto", 512), ("auto", 256), ("bdf", 0), ("auto", 4096)):
    m.solver = solver
    if budget:
        m.EXPLICIT_STEP_BUDGET = budget
    for rep in range(2):
        np.random.seed(1)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s = m.MCMC(chain_inits=4096, iterations_per_chain=its, fitsurvey_samples=n, sd_fitdistance=6.0, print_report=False,
                   posterior="summary", rng="philox")
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out = m._last_mcmc
    sc = np.asarray(out["step_count"])
    print(solver, budget, "wall %.3f s" % dt, "last kernel ms %.1f" % m._device().last_kernel_ms(), "rerun", m._last_rerun, "used", m._last_solver,
          "steps/chain-step mean %.0f max %.0f" % (sc.mean() / its, sc.max() / its), "best chi %.3f" % s.best_chi, flush=True)
