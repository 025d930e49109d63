"""Dev tool (GPU): time of the device-side reference stream generator against the host one."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model
from odelib_b200 import _capi
dm, _ = device_model("two_i")
for C, n_iter in ((4096, 999), (4096, 9999), (65536, 999)):
    seeds = np.arange(C, dtype=np.uint32)
    dm.reference_streams(seeds[:64], 10, 5, 5)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    z, u = dm.reference_streams(seeds, n_iter, 5, 5)
    torch.cuda.synchronize(); t_dev = time.perf_counter() - t0
    t_host = float("nan")
    if C * n_iter <= 4096 * 999:
        zh = np.empty((C, n_iter, 5)); uh = np.empty((C, n_iter))
        t0 = time.perf_counter()
        _capi.check(_capi.lib().odl_reference_streams(seeds.ctypes.data, C, n_iter, 5, 5, 0.05, zh.ctypes.data, uh.ctypes.data))
        t_host = time.perf_counter() - t0
        print("  identical u:", bool(np.array_equal(u.cpu().numpy(), uh)), " identical z share:", float((z.cpu().numpy() == zh).mean()),
              " max rel diff:", float(np.max(np.abs(z.cpu().numpy() - zh) / np.abs(zh))))
    print("chains %d iterations %d: device %.3f s (%.1f GB of streams), host %.3f s" % (C, n_iter, t_dev, (z.numel() + u.numel()) * 8 / 1e9, t_host), flush=True)
    del z, u
