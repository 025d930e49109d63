"""Dev tool (GPU): A/B of kernel variants (ODL_KERNEL_DEFINES) on the 1M-set two_i AUTO sweep.
    python tools/variant_ab.py "" "-DODL_INNER=4" ...   -> best-of-6 kernel / pass times per variant"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws  # noqa: E402

theta = torch.from_numpy(prior_draws("two_i", 1 << 20, seed=0)).cuda()
for variant in (sys.argv[1:] or [""]):
    if variant:
        os.environ["ODL_KERNEL_DEFINES"] = variant
    else:
        os.environ.pop("ODL_KERNEL_DEFINES", None)
    dm, _ = device_model("two_i")
    best, passes = 1e9, None
    for rep in range(7):
        dm.sweep(theta, solver="auto")
        torch.cuda.synchronize()
        if rep and dm.last_kernel_ms() < best:
            best, passes = dm.last_kernel_ms(), dm.last_pass_ms()
    print("%-70s total %.3f ms  passes %s  regs %s" % (variant or "(default)", best, [round(x, 3) for x in passes], dm.kernel_info("sweep")), flush=True)
    dm.close()
