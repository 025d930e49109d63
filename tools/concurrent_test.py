"""Dev tool (GPU): ODL_AUTO_CONCURRENT with 64-thread bulk CTAs (finer register granularity beside the stiff warps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import device_model, prior_draws
n = 1 << 20
theta = torch.from_numpy(prior_draws("two_i", n, seed=0)).cuda()
for blk, mb in ((64, 8), (128, 4)):
    dm, tab = device_model("two_i", block_threads=blk, min_blocks=mb)
    for label, kw in (("sequential", dict()), ("sequential noearly", dict(early_check_steps=-1)),
                      ("concurrent tw2 noearly", dict(auto_flags=2, tail_warps=2, early_check_steps=-1)),
                      ("concurrent tw3 noearly", dict(auto_flags=2, tail_warps=3, early_check_steps=-1)),
                      ("concurrent tw4 noearly", dict(auto_flags=2, tail_warps=4, early_check_steps=-1)),
                      ("concurrent tw3 early", dict(auto_flags=2, tail_warps=3)),
                      ("concurrent tw4 early", dict(auto_flags=2, tail_warps=4))):
        for rep in range(3):
            out = dm.sweep(theta, solver="auto", max_steps=200000, **kw)
        torch.cuda.synchronize()
        print(blk, mb, f"{label:26s}", "kernel_ms", round(dm.last_kernel_ms(), 3), [round(x, 3) for x in dm.last_pass_ms()],
              "ok", float((out["status"] == 0).double().mean()), flush=True)
    dm.close()
