"""Dev tool (GPU): the host-memory two-piece sweep under every scheduling flag, one call at a time."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws  # noqa: E402

dm, _ = device_model("two_i")
n = (1 << 18) + 777
theta = prior_draws("two_i", n, seed=13)
ref = None
for label, kw, dev in (("sequential one piece", dict(auto_flags=8 | 4), False), ("sequential two pieces", dict(auto_flags=8), False),
                       ("beside one piece", dict(auto_flags=4), False), ("beside device", dict(), True),
                       ("beside two pieces", dict(), False)):
    try:
        out = dm.sweep(torch.from_numpy(theta).cuda() if dev else theta, solver="auto", max_steps=200000, **kw)
        torch.cuda.synchronize()
        chi = out["chi"].cpu().numpy() if dev else out["chi"]
        if ref is None:
            ref = chi
        print(label, "ok", "equal" if np.array_equal(ref, chi, equal_nan=True) else "DIFFERENT", dm.last_pass_ms(), flush=True)
    except Exception as exc:  # noqa: BLE001
        print(label, "FAILED", exc, flush=True)
        break
