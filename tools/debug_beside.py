"""Dev tool (GPU): the 1M-row AUTO sweep with the stiff pass beside the bulk pass -- back-to-back calls with / without
host synchronisation and L2 flush in between (what bench.py does), pass times of the last call."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws  # noqa: E402

dm, _ = device_model("two_i")
kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])}
theta = torch.from_numpy(prior_draws("two_i", 1 << 20, seed=0)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    dm.sweep(theta, solver="auto", **kw)
torch.cuda.synchronize()
for label, sync, fl in (("sync, no flush", True, False), ("no sync, flush", False, True), ("no sync, flush", False, True)):
    ev = []
    for rep in range(5):
        if fl:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dm.sweep(theta, solver="auto", **kw)
        b.record()
        ev.append((a, b))
        if sync:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    print("%-20s" % label, "per call ms", [round(a.elapsed_time(b), 3) for a, b in ev], "last passes", [round(x, 3) for x in dm.last_pass_ms()], flush=True)
