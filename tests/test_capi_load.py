"""CPU-only checks of the C-ABI library: it loads, exports what include/odelib_b200.h declares, NVRTC
compiles the demo models for sm_100a offline, and compute entry points refuse loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from odelib_b200 import _capi, demo_models, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "odelib_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(odl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/odelib_b200.h but not exported"
    assert sorted(_capi.EXPORTS) == names
    assert L.odl_abi_version() == 3


@pytest.mark.parametrize("name", ["zero_i", "one_i", "two_i"])
def test_nvrtc_compiles_demo_models_offline(name, tmp_path):
    f, n, P, g = demo_models.MODELS[name]
    m = engine.DeviceModel(f, n, P, g, compile_only=True, cache_dir=str(tmp_path))
    # one NVRTC program (one cubin) per kernel unit, compiled side by side on host threads
    units = ["sweep", "traj", "mcmc", "sweep_bdf", "mcmc_bdf", "sweep_ros23", "mcmc_ros23", "mcmc_auto", "sweep_radau5",
             "mcmc_radau5", "order"]
    cubins = sorted(p for p in os.listdir(tmp_path) if p.endswith(".cubin"))
    assert len(cubins) == len(units) and all(os.path.getsize(tmp_path / c) > 3000 for c in cubins)
    assert sorted(c.split("_", 2)[2][:-len(".cubin")] for c in cubins) == sorted(units)
    for u in units:
        sec, hit = m.unit_seconds(u)
        assert sec > 0 and not hit
    assert m.unit_seconds("sweep_coop") == (-1.0, False)           # n <= 8: no cooperative kernels
    # second build hits the cache, unit by unit
    m2 = engine.DeviceModel(f, n, P, g, compile_only=True, cache_dir=str(tmp_path))
    assert m2.build_log.count("cache hit") == len(units) and all(m2.unit_seconds(u)[1] for u in units)
    m.close(); m2.close()
    # compile_only = 2: only what the default paths launch (ordering, DOPRI5 sweep, BDF stiff pass, trajectories, chains)
    sub = tmp_path / "default_only"
    m3 = engine.DeviceModel(f, n, P, g, compile_only=2, cache_dir=str(sub))
    got = sorted(c.split("_", 2)[2][:-len(".cubin")] for c in os.listdir(sub) if c.endswith(".cubin"))
    assert got == sorted(["order", "sweep", "sweep_bdf", "traj", "mcmc"])
    m3.close()


def test_bad_model_source_reports_nvrtc_log():
    L = _capi.lib()
    h = ctypes.c_void_p()
    bo = _capi.BuildOpts(); bo.device = -1; bo.dense_output = 1; bo.compile_only = 1
    rc = L.odl_model_create(b"#define ODL_N 1\n#define ODL_P 1\n#define ODL_NOUT 1\nthis is not CUDA;\n", 1, 1, 1,
                            ctypes.byref(bo), ctypes.byref(h))
    assert rc == _capi.ECOMPILE
    assert b"error" in L.odl_last_error()


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    f, n, P, g = demo_models.MODELS["zero_i"]
    with pytest.raises(_capi.OdlError) as e:
        engine.DeviceModel(f, n, P, g)
    assert e.value.code == _capi.ENODEVICE
    m = engine.DeviceModel(f, n, P, g, compile_only=True)
    with pytest.raises(_capi.OdlError):
        m.sweep(np.ones((4, 3)))
    # the device-side helpers either side of the path refuse as loudly (no CPU fallback anywhere)
    L = _capi.lib()
    cnt = ctypes.c_longlong(0)
    buf = np.zeros(8)
    kind = np.zeros(3, np.int32)
    assert L.odl_select_below(m._h, buf.ctypes.data, 8, 1.0, buf.ctypes.data, ctypes.byref(cnt), None) == _capi.ENODEVICE
    assert L.odl_gather_rows(m._h, buf.ctypes.data, 2, None, buf.ctypes.data, 1, buf.ctypes.data, None) == _capi.ENODEVICE
    assert L.odl_sample_lhs(m._h, 2, 3, kind.ctypes.data, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, 0,
                            buf.ctypes.data, None) == _capi.ENODEVICE
    assert b"no CPU fallback" in L.odl_last_error()
    m.close()


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    z = engine.philox4x32_10(np.zeros(4, np.uint32), np.zeros(2, np.uint32))
    assert [int(x) for x in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = engine.philox4x32_10(np.full(4, 0xffffffff, np.uint32), np.full(2, 0xffffffff, np.uint32))
    assert [int(x) for x in f] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    zz, uu = engine.philox_streams(7, np.arange(4), 50, 5)
    assert zz.shape == (4, 50, 5) and uu.shape == (4, 50)
    assert 0 <= uu.min() and uu.max() < 1 and abs(zz.std() / 0.05 - 1) < 0.1


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof / field offsets of the option structs as gcc sees include/odelib_b200.h vs the ctypes mirror."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "odelib_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(odl_solver_opts), '
                   'offsetof(odl_solver_opts, max_steps), offsetof(odl_solver_opts, tail_solver), '
                   'offsetof(odl_solver_opts, auto_flags), sizeof(odl_mcmc_opts), offsetof(odl_mcmc_opts, seed), '
                   'sizeof(odl_mcmc_io), sizeof(odl_build_opts)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    S, M = _capi.SolverOpts, _capi.McmcOpts
    assert got == [ctypes.sizeof(S), S.max_steps.offset, S.tail_solver.offset, S.auto_flags.offset, ctypes.sizeof(M),
                   M.seed.offset, ctypes.sizeof(_capi.McmcIO), ctypes.sizeof(_capi.BuildOpts)]
