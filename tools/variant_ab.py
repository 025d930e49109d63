"""Dev tool (GPU): A/B of kernel variants (ODL_KERNEL_DEFINES) on the 1M-set two_i AUTO sweep.
    python tools/variant_ab.py "" "-DODL_INNER=4" ...   -> best-of-6 kernel / pass times per variant"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws  # noqa: E402

MODEL = os.environ.get("MODEL", "two_i")
theta = torch.from_numpy(prior_draws(MODEL, 1 << 20, seed=0)).cuda()
for variant in (sys.argv[1:] or [""]):
    kw, skw = {}, {}
    if "@" in variant:                      # "...@tail_lanes=16,early_check_steps=-1": options of the sweep call
        variant, sopts = variant.split("@", 1)
        skw = {k: int(v) for k, v in (item.split("=") for item in sopts.split(",") if item)}
        label_extra = "@" + sopts
    else:
        label_extra = ""
    if "|" in variant:                      # "block_threads=64,min_blocks=8|-DODL_INNER=8": build options | defines
        opts, variant_defs = variant.split("|", 1)
        kw = {k: int(v) for k, v in (item.split("=") for item in opts.split(",") if item)}
    else:
        variant_defs = variant
    label, variant = variant + label_extra, variant_defs
    if variant:
        os.environ["ODL_KERNEL_DEFINES"] = variant
    else:
        os.environ.pop("ODL_KERNEL_DEFINES", None)
    dm, _ = device_model(MODEL, **kw)
    best, passes = 1e9, None
    for rep in range(7):
        dm.sweep(theta, solver="auto", **skw)
        torch.cuda.synchronize()
        if rep and dm.last_kernel_ms() < best:
            best, passes = dm.last_kernel_ms(), dm.last_pass_ms()
    print("%-70s total %.3f ms  passes %s  regs %s bdf %s" % (label or "(default)", best, [round(x, 3) for x in passes], dm.kernel_info("sweep"), dm.kernel_info("sweep_bdf")), flush=True)
    dm.close()
