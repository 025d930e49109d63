"""Dev tool (GPU): fit_survey(samples) end to end -- numpy sampler vs device sampler -- and MCMC(chain_inits=int)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_facade_host import make_model
m = make_model("two_i")
m.fit_survey(samples=1000)
for n in (100000, 1000000):
    for sampler in ("host", "device"):
        np.random.seed(1)
        m.fit_survey(samples=n, sampler=sampler)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sv = m.fit_survey(samples=n, sampler=sampler)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"fit_survey({n}) sampler={sampler}: {dt * 1e3:8.1f} ms, chi<666: {(sv['chi'] < 666).mean():.4f}", flush=True)
for C, fs in ((4096, 1000000),):
    np.random.seed(2)
    t0 = time.perf_counter()
    s = m.MCMC(chain_inits=C, iterations_per_chain=200, fitsurvey_samples=fs, sd_fitdistance=6.0, print_report=False, posterior="summary")
    dt = time.perf_counter() - t0
    print(f"MCMC(chain_inits={C}, 200 its, survey {fs}, posterior='summary'): {dt * 1e3:8.1f} ms total; best chi {s.best_chi:.3f}; rhat max {max(s.rhat.values()):.3f}", flush=True)
