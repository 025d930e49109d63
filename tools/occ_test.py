"""Dev tool (GPU): 1M-set AUTO sweep per model against CTA size / resident CTAs per SM of the bulk kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import device_model, prior_draws
n = 1 << 20
for name, variants in (("zero_i", ((128, 4), (128, 6), (128, 8), (256, 3), (256, 4))),
                       ("one_i", ((128, 4), (128, 5), (128, 6), (256, 3)))):
    theta = torch.from_numpy(prior_draws(name, n, seed=0)).cuda()
    for blk, mb in variants:
        dm, tab = device_model(name, block_threads=blk, min_blocks=mb)
        for rep in range(3):
            out = dm.sweep(theta, solver="auto", max_steps=200000)
        torch.cuda.synchronize()
        print(name, blk, mb, dm.kernel_info("sweep"), "kernel_ms", round(dm.last_kernel_ms(), 4), [round(x, 4) for x in dm.last_pass_ms()],
              "ok", float((out["status"] == 0).double().mean()), flush=True)
        dm.close()
