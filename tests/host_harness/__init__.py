"""Builds tests/host_harness/harness.cpp for one traced model with g++ and loads it with ctypes.
Test infrastructure: lets the CPU suite exercise the stepper code paths of odl_kernels.cuh."""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build(ode, n_state, n_param, groups=None):
    from odelib_b200.tracer import trace
    src = trace(ode, n_state, n_param).cuda_source(fmad=True, observe_groups=groups)
    kern = open(os.path.join(ROOT, "odelib_b200", "csrc", "odl_kernels.cuh")).read()
    key = hashlib.sha1((src + kern + open(os.path.join(HERE, "harness.cpp")).read() +
                        os.environ.get("ODL_HARNESS_DEFINES", "")).encode()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), "odl_harness")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, f"h_{key}.so")
    if not os.path.exists(so):
        hdr = os.path.join(d, f"m_{key}.h")
        open(hdr, "w").write(src)
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", f'-DODL_MODEL_HEADER="{hdr}"',
               *os.environ.get("ODL_HARNESS_DEFINES", "").split(),
               "-I" + os.path.join(ROOT, "odelib_b200", "csrc"), os.path.join(HERE, "harness.cpp"), "-o", so]
        subprocess.run(cmd, check=True)
    lib = C.CDLL(so)
    lib.harness_solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double,
                                  C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    return lib


def solve(lib, solver, theta, slot_t, y0, rtol, atol, max_steps=2000000):
    theta = np.ascontiguousarray(theta, np.float64)
    slot_t = np.ascontiguousarray(slot_t, np.float64)
    y0 = np.ascontiguousarray(y0, np.float64)
    out = np.full((slot_t.size, y0.size), np.nan)
    ns = C.c_int()
    st = lib.harness_solve({"dopri5": 0, "ros23": 1, "radau5": 3, "bdf": 4}[solver], theta.ctypes.data, slot_t.ctypes.data,
                           slot_t.size, y0.ctypes.data, float(slot_t[0]) if slot_t[0] <= 0 else 0.0, rtol, atol,
                           max_steps, out.ctypes.data, C.byref(ns))
    return out, st, ns.value


def solve_observed(lib, theta, slot_t, y0, rtol, atol, max_steps=2000000):
    """DOPRI5 through the observed-columns sink of the sweep / chain kernels -> (out [n_slot, n_out], status, steps)."""
    theta = np.ascontiguousarray(theta, np.float64)
    slot_t = np.ascontiguousarray(slot_t, np.float64)
    y0 = np.ascontiguousarray(y0, np.float64)
    lib.harness_solve_observed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                           C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    out = np.full((slot_t.size, lib.harness_nout()), np.nan)
    ns = C.c_int()
    st = lib.harness_solve_observed(theta.ctypes.data, slot_t.ctypes.data, slot_t.size, y0.ctypes.data,
                                    float(slot_t[0]) if slot_t[0] <= 0 else 0.0, rtol, atol, max_steps, out.ctypes.data,
                                    C.byref(ns))
    return out, st, ns.value


def solve_handover(lib, theta, slot_t, y0, rtol, atol, cap=704, early=384, idle=0):
    """DOPRI5 until it gives the row up, `idle` attempts of the stopped solve, BDF continuing from there (the AUTO sweep's
    hand-over) -> (out [n_slot, n], status, dopri5 attempts, bdf steps, slot at the hand-over, time at the hand-over)."""
    theta = np.ascontiguousarray(theta, np.float64)
    slot_t = np.ascontiguousarray(slot_t, np.float64)
    y0 = np.ascontiguousarray(y0, np.float64)
    lib.harness_solve_handover.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                           C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    out = np.full((slot_t.size, y0.size), np.nan)
    info = np.zeros(3, np.int32)
    th = C.c_double()
    st = lib.harness_solve_handover(theta.ctypes.data, slot_t.ctypes.data, slot_t.size, y0.ctypes.data,
                                    float(slot_t[0]) if slot_t[0] <= 0 else 0.0, rtol, atol, cap, early, idle,
                                    out.ctypes.data, info.ctypes.data, C.byref(th))
    return out, st, int(info[0]), int(info[1]), int(info[2]), th.value
