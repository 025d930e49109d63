"""Dev tool (GPU): where the wall time of ModelFramework.fit_survey(1M) goes."""
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
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sv = m.fit_survey(samples=n)
    torch.cuda.synchronize(); print("fit_survey wall %.4f s" % (time.perf_counter() - t0))
pr = cProfile.Profile()
pr.enable()
sv = m.fit_survey(samples=n)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
