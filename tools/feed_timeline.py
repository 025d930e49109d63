"""Dev tool (GPU): when do rows reach the stiff pass that runs beside the bulk pass, and when are they done?
Kernels built with -DODL_TIMELINE=1; one 1M-row AUTO sweep per scenario (idle GPU / behind an L2 flush)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["ODL_KERNEL_DEFINES"] = "-DODL_TIMELINE=1 " + os.environ.get("EXTRA_DEFINES", "")
os.environ["ODL_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odelib_b200 import _capi  # noqa: E402
from tests.helpers import device_model, prior_draws  # noqa: E402

n = 1 << 20
dm, _ = device_model("two_i")
theta = torch.from_numpy(prior_draws("two_i", n, seed=0)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])}
for _ in range(3):
    dm.sweep(theta, solver="auto", **kw)
torch.cuda.synchronize()
for label, fl, reps in (("idle GPU", False, 1), ("behind an L2 flush", True, 1), ("4th of 4 back-to-back calls, flush before each", True, 4),
                        ("4th of 4 back-to-back calls", False, 4), ("idle GPU", False, 1)):
    for _ in range(reps):
        if fl:
            flush.fill_(1)
        out = dm.sweep(theta, solver="auto", **kw)
    torch.cuda.synchronize()
    cnt = np.zeros(128, np.int32)
    _capi.check(dm._L.odl_debug_counters(dm._h, cnt.ctypes.data, 128))
    m = int(cnt[16])
    tl = np.zeros((m, 3), np.int64)
    _capi.check(dm._L.odl_debug_timeline(dm._h, tl.ctypes.data, m))
    tail = np.zeros((n, 3), np.int64)
    _capi.check(dm._L.odl_debug_timeline(dm._h, tail.ctypes.data, n))
    sm = tail[::-1][:64]
    sm = sm[sm[:, 0] > 0]
    smids = np.sort(sm[:, 0] - 1)
    t0 = tl[:, 0].min()
    print("  consumer CTAs on SMs", smids.tolist(), "| both SMs of a TPC taken:", int(np.sum(np.diff(smids // 2) == 0)),
          "| resident (ms rel. to first arrival) min/max", round((sm[:, 1].min() - t0) * 1e-6, 3), round((sm[:, 1].max() - t0) * 1e-6, 3))
    arr, beg, end = (tl[:, 0] - t0) * 1e-6, (tl[:, 1] - t0) * 1e-6, (tl[:, 2] - t0) * 1e-6
    picked = tl[:, 1] > 0
    ns = out["nsteps"].cpu().numpy()
    q = lambda x: np.round(np.percentile(x, [0, 10, 50, 90, 99, 100]), 3).tolist()
    print(f"--- {label}: {m} feed entries, consumer took {int(picked.sum())}; passes {[round(x, 3) for x in dm.last_pass_ms()]}")
    print("  arrival ms      pct 0/10/50/90/99/100", q(arr))
    print("  wait (start-arrival)                 ", q((beg - arr)[picked]))
    print("  solve duration                       ", q((end - beg)[picked]))
    print("  end                                  ", q(end[picked]))
    print("  steps/ms of the solves (pace)        ", q((ns[out["status"].cpu().numpy() == 0][:1] * 0 + 1)), "BDF rows:", int((ns > 512).sum()),
          "mean duration", round(float((end - beg)[picked].mean()), 3), "sum of durations (lane-ms)", round(float((end - beg)[picked].sum()), 1))
    late = np.argsort(end)[-5:]
    print("  last five to end: arrival", np.round(arr[late], 3).tolist(), "start", np.round(beg[late], 3).tolist(), "end", np.round(end[late], 3).tolist())
