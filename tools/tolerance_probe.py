"""Dev tool (GPU): how far the default-tolerance results really are from the reference (what the thresholds of
tests/test_gpu_sweep.py / test_gpu_mcmc.py should be): GPU(default) vs reference(default), each vs reference(1e-13)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, golden  # noqa: E402


def rel(a, b, atol=0.0):
    return float(np.max(np.abs(a - b) / (np.abs(b) + atol)))


for name in ("zero_i", "one_i", "two_i"):
    g = golden(name)
    dm, _ = device_model(name)
    out = dm.sweep(g["theta"], return_pred=True)
    ok = np.all(g["pred_def"] > 1.0, axis=1) & np.all(g["pred_tight"] > 1.0, axis=1) & (out["status"] == 0)
    print(name, "rows", int(ok.sum()), "of", len(ok))
    for key, a, d, t in (("pred", out["pred"], g["pred_def"], g["pred_tight"]), ("chi", out["chi"], g["chi_def"], g["chi_tight"]),
                         ("r2", out["r2"], g["r2_def"], g["r2_tight"])):
        print("  %-5s gpu_def vs ref_def %.3g | gpu_def vs ref_tight %.3g | ref_def vs ref_tight %.3g" % (
            key, rel(a[ok], d[ok]), rel(a[ok], t[ok]), rel(d[ok], t[ok])))
    for tag, tol in (("def", None), ("tight", 1e-13)):
        pre = f"chain_{tag}_s0_"
        nits = int(g[pre + "nits"])
        o = dm.mcmc(g[pre + "theta0"][None, :], nits=nits, rng_mode="forced", forced=g[pre + "proposals"][None],
                    u=g[pre + "u"][None], rtol=tol, atol=tol, trace=True, pnum=int(g["pnum"]), max_steps=2000000)
        fin = np.isfinite(g[pre + "chinew"])
        c, r = o["chinew"][0][fin], g[pre + "chinew"][fin]
        # margin of every decision: |(chi - chinew) - ln u| along the reference chain
        cur = float(g[pre + "chi0"]); margins = []
        for k in range(len(g[pre + "u"])):
            margins.append(abs((cur - g[pre + "chinew"][k]) - np.log(g[pre + "u"][k])))
            if g[pre + "accepted"][k]:
                cur = g[pre + "chinew"][k]
        margins = np.array(margins)
        print("  chain %-5s chinew max rel %.3g max abs %.3g | decisions differing %d | smallest decision margins %s" % (
            tag, rel(c, r), float(np.max(np.abs(c - r))), int((o["accepted"][0].astype(bool) != g[pre + "accepted"]).sum()),
            np.sort(margins[np.isfinite(margins)])[:3]))
    dm.close()
