import sys, os, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from tests.helpers import device_model, prior_draws
n = 1 << 20
theta = torch.from_numpy(prior_draws("two_i", n, seed=0)).cuda()
for blk, mb in ((128, 4), (128, 5), (160, 4), (96, 6), (64, 8), (64, 10), (256, 2)):
    dm, tab = device_model("two_i", block_threads=blk, min_blocks=mb)
    for rep in range(3):
        out = dm.sweep(theta, solver="auto", max_steps=200000)
    torch.cuda.synchronize()
    print(blk, mb, dm.kernel_info("sweep"), "kernel_ms", round(dm.last_kernel_ms(), 4), [round(x, 4) for x in dm.last_pass_ms()], flush=True)
    dm.close()
