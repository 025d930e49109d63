"""Dev tool (GPU): host-memory AUTO sweep of 1M rows (pinned buffers, chi + status back) over (cap, consumer SMs, first piece)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws
n = 1 << 20
dm, tab = device_model("two_i")
theta = torch.from_numpy(prior_draws("two_i", n, seed=0)).pin_memory().numpy()
outs = ("chi", "status")
out = {"chi": torch.empty(n, dtype=torch.float64).pin_memory().numpy(), "status": torch.empty(n, dtype=torch.int32).pin_memory().numpy()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for cfg in sys.argv[1:]:
    cap, early, sms, first = cfg.split(",")
    if first == "default":
        os.environ.pop("ODL_FIRST_PIECE", None)
    else:
        os.environ["ODL_FIRST_PIECE"] = first
    kw = dict(solver="auto", max_steps=500000, out=out, outputs=outs, pass_caps=int(cap), early_check_steps=int(early), tail_warps=int(sms))
    for _ in range(2):
        dm.sweep(theta, **kw)
    ts = []
    for _ in range(8):
        flush.fill_(1); torch.cuda.synchronize(); t0 = time.perf_counter()
        dm.sweep(theta, **kw)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("cap %s early %s sms %s first %s: ms per call median %.3f min %.3f  passes %s" % (cap, early, sms, first, np.median(ts) * 1e3, min(ts) * 1e3,
          [round(x, 3) for x in dm.last_pass_ms()]), flush=True)
