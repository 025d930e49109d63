"""ctypes binding of libodelib_b200.so (include/odelib_b200.h).  No fallback: if the library is missing or
the GPU is absent, the product raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libodelib_b200.so")
CACHE_DIR = os.path.join(_HERE, "_cubin_cache")

SUCCESS, EINVAL, ECUDA, ECOMPILE, ENODEVICE, EIO = range(6)
MEM_HOST, MEM_DEVICE = 0, 1
SOLVER_DOPRI5, SOLVER_ROS23, SOLVER_AUTO, SOLVER_RADAU5, SOLVER_BDF = 0, 1, 2, 3, 4
RNG_PHILOX, RNG_HOST_STREAMS, RNG_FORCED = 0, 1, 2
SAMPLES_CHAIN_MAJOR, SAMPLES_ITERATION_MAJOR = 0, 1
COMM_ID_BYTES = 128
RHAT_LOCAL = 256
AUTO_UNORDERED, AUTO_CONCURRENT, AUTO_ONE_PIECE, AUTO_SEQUENTIAL, AUTO_NO_HELPER, AUTO_NO_HANDOVER = 1, 2, 4, 8, 16, 32
ST_OK, ST_MAXSTEPS, ST_NONFINITE, ST_HUNDERFLOW, ST_STIFF, ST_ALLMASKED = 0, 1, 2, 3, 4, 8

EXPORTS = ["odl_abi_version", "odl_last_error", "odl_model_create", "odl_model_destroy", "odl_model_build_log",
           "odl_model_kernel_info", "odl_model_set_data", "odl_model_set_grid", "odl_sweep", "odl_trajectory",
           "odl_mcmc", "odl_model_last_kernel_ms", "odl_model_last_pass_ms", "odl_launch_count", "odl_fp64_peak", "odl_debug_counters", "odl_select_below", "odl_gather_rows", "odl_sample_lhs", "odl_reference_streams", "odl_reference_streams_device",
           "odl_model_unit_seconds", "odl_debug_timeline", "odl_comm_unique_id", "odl_comm_init", "odl_comm_destroy",
           "odl_rhat"]


class OdlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libodelib_b200 error {code}: {msg}")
        self.code = code


class BuildOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("block_threads", C.c_int), ("min_blocks", C.c_int), ("dense_output", C.c_int),
                ("compile_only", C.c_int), ("y0_from_param", C.c_int), ("coop_lanes", C.c_int), ("reserved", C.c_int * 1),
                ("cache_dir", C.c_char_p)]


class SolverOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("h0", C.c_double), ("hmax", C.c_double),
                ("max_steps", C.c_int), ("solver", C.c_int), ("stiff_check", C.c_int), ("stiff_min_steps", C.c_int),
                ("pass_cap0", C.c_int), ("tail_warps", C.c_int), ("tail_solver", C.c_int), ("early_check_steps", C.c_int),
                ("tail_lanes", C.c_int), ("auto_flags", C.c_int)]


class McmcOpts(C.Structure):
    _fields_ = [("n_chain", C.c_int), ("chain_offset", C.c_int), ("nits", C.c_int), ("burnin", C.c_int),
                ("it_begin", C.c_int), ("it_end", C.c_int), ("rng_mode", C.c_int), ("n_walk", C.c_int),
                ("walk", C.POINTER(C.c_int)), ("pnum", C.c_int), ("row_stride", C.c_int), ("step_sd", C.c_double),
                ("seed", C.c_ulonglong), ("speculate", C.c_int), ("sample_layout", C.c_int),
                ("stop_failed_chains", C.c_int), ("reserved", C.c_int)]


class McmcIO(C.Structure):
    _fields_ = [("theta", C.c_void_p), ("chain_state", C.c_void_p), ("samples", C.c_void_p),
                ("summaries", C.c_void_p), ("z", C.c_void_p), ("u", C.c_void_p), ("forced", C.c_void_p),
                ("trace_chinew", C.c_void_p), ("trace_accept", C.c_void_p), ("fail_count", C.c_void_p),
                ("step_count", C.c_void_p), ("best_theta", C.c_void_p), ("chain_ids", C.c_void_p), ("prior_table", C.c_void_p)]


_lib = None


def lib():
    """Load the shared library (built in-tree by odelib_b200/csrc/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OdlError(EIO, f"{LIB_PATH} not found - run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.odl_abi_version.restype = C.c_int
    L.odl_last_error.restype = C.c_char_p
    L.odl_launch_count.restype = C.c_longlong
    L.odl_model_build_log.restype = C.c_char_p
    L.odl_model_build_log.argtypes = [C.c_void_p]
    L.odl_model_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(BuildOpts), C.POINTER(C.c_void_p)]
    L.odl_model_destroy.argtypes = [C.c_void_p]
    L.odl_model_kernel_info.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.odl_model_set_data.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double]
    L.odl_model_set_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.odl_sweep.argtypes = [C.c_void_p, C.POINTER(SolverOpts), C.c_longlong, C.c_void_p, C.c_int, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.odl_trajectory.argtypes = [C.c_void_p, C.POINTER(SolverOpts), C.c_longlong, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.odl_mcmc.argtypes = [C.c_void_p, C.POINTER(SolverOpts), C.POINTER(McmcOpts), C.POINTER(McmcIO), C.c_int,
                           C.c_void_p]
    L.odl_model_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.odl_model_last_pass_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.odl_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
    L.odl_debug_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.odl_select_below.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_void_p, C.POINTER(C.c_longlong),
                                   C.c_void_p]
    L.odl_sample_lhs.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_ulonglong, C.c_void_p, C.c_void_p]
    L.odl_reference_streams.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    L.odl_reference_streams_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
    L.odl_gather_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p,
                                  C.c_void_p]
    L.odl_model_unit_seconds.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.odl_comm_unique_id.argtypes = [C.c_void_p]
    L.odl_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.odl_comm_destroy.argtypes = [C.c_void_p]
    L.odl_rhat.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_longlong), C.c_void_p]
    L.odl_fp64_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    if L.odl_abi_version() != 3:
        raise OdlError(EIO, "libodelib_b200.so ABI version mismatch - rebuild")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise OdlError(rc, lib().odl_last_error().decode(errors="replace"))
