"""Loader for the CUDA library behind the C ABI in include/ddp_b200.h.

There is no CPU fallback: if ``libddp_b200.so`` is missing or does not load, importing the
solver raises.  ``build()`` compiles it in-tree with nvcc for sm_100a (cross-compiles without
a GPU), so the built file travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DDP_B200_LIB") or os.path.join(_HERE, "libddp_b200.so")
_CSRC = os.path.join(_HERE, "csrc")
_SOURCES = [os.path.join(_CSRC, "ddp_api.cu")]
_DEPS = _SOURCES + [os.path.join(_CSRC, f) for f in ("kernels.cuh", "backward_sym.cuh", "quadruped_fused.cuh", "quadruped_quat_fused.cuh", "quadruped_rollout.cuh", "arm_rollout.cuh", "models.h", "dual.h")] + [
    os.path.join(_HERE, "..", "include", "ddp_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

# symbols declared in include/ddp_b200.h
SYMBOLS = [
    "ddp_last_error", "ddp_model_dims", "ddp_workspace_bytes", "ddp_create", "ddp_destroy",
    "ddp_set_options", "ddp_set_keypoints", "ddp_set_regularization", "ddp_set_cost", "ddp_set_target",
    "ddp_set_initial_state", "ddp_set_initial_guess", "ddp_reset", "ddp_begin_solve",
    "ddp_iterate", "ddp_iterate_linesearch", "ddp_iterate_finish_async", "ddp_iterate_wait", "ddp_solve",
    "ddp_run_phase", "ddp_get",
    "ddp_put", "ddp_get_int", "ddp_device_ptr", "ddp_array_elems", "ddp_last_timings",
    "ddp_launch_count", "ddp_peak_fp64", "ddp_mpc_shift", "ddp_set_mpc_rearm", "ddp_apply_staged_inputs", "ddp_set_control_limits",
]

# enums of include/ddp_b200.h
X_BAR, U_BAR, K, KAPPA, DV, FX, FU, COST, EPS, IMPROVEMENT, X0, X_NOM = range(12)
CAND_COST, CAND_EXPECTED, CAND_X, CAND_U, CONVERGED_COST = 12, 13, 14, 15, 16
I_STATUS, I_LS_ITERS, I_ITERS, I_NUM_KEYPOINTS, I_KEYPOINTS, I_ACTIVE, I_RESOLVES = range(7)
PHASE_LINESEARCH, PHASE_DERIVATIVES, PHASE_BACKWARD = range(3)
KP_METHODS = {"setInterval": 0, "adaptiveJerk": 1, "iterativeError": 2}
TRAJ_RUNNING, TRAJ_CONVERGED, TRAJ_LINESEARCH_FAILED = 0, 1, 2


# DMMAs backward_sym_kernel issues per step at (n, m) = (36, 12): 270 (W = Vxx S) + 189 (upper tiles of
# S'W) + 30 (K) + 45 (upper tiles of the Vxx update), 8 x 8 tile padding included, + 24 per Newton-
# Schulz pass, three passes as a rule
BWD_DMMA_PER_STEP_36_12 = 534 + 24 * 3


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in _DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> drake_ddp_b200/libddp_b200.so"""
    if not force and not is_stale():
        return LIB_PATH
    tmp = LIB_PATH + f".{os.getpid()}.tmp"
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + _SOURCES + ["-o", tmp]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


# Test-only build for compute-sanitizer's synccheck / racecheck: the same sources with the
# intra-CTA hand-offs of backward_sym_kernel as non-aligned named barriers instead of aligned
# ones + mbarriers (csrc/backward_sym.cuh, DDP_SANITIZER_BUILD, says why).  Never loaded by the
# product: only tests/test_gpu_parity.py points DDP_B200_LIB at it for those two passes.
RACECHECK_LIB_PATH = os.path.join(_HERE, "libddp_b200_racecheck.so")


def build_racecheck(force: bool = False) -> str:
    if not force and os.path.exists(RACECHECK_LIB_PATH):
        t = os.path.getmtime(RACECHECK_LIB_PATH)
        if not any(os.path.exists(d) and os.path.getmtime(d) > t for d in _DEPS):
            return RACECHECK_LIB_PATH
    tmp = RACECHECK_LIB_PATH + f".{os.getpid()}.tmp"
    subprocess.check_call(["nvcc"] + NVCC_FLAGS + ["-DDDP_SANITIZER_BUILD"] + _SOURCES + ["-o", tmp])
    os.replace(tmp, RACECHECK_LIB_PATH)
    return RACECHECK_LIB_PATH


_lib = None


def lib():
    """The loaded library with argtypes set.  Raises if the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    L = ctypes.CDLL(LIB_PATH)
    c_int, c_dbl, c_vp, c_sz = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t
    ip, dp = ctypes.POINTER(c_int), ctypes.POINTER(c_dbl)
    L.ddp_last_error.restype = ctypes.c_char_p
    L.ddp_last_error.argtypes = []
    L.ddp_model_dims.argtypes = [c_int, ip, ip, ip]
    L.ddp_workspace_bytes.restype = c_sz
    L.ddp_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    L.ddp_create.argtypes = [ctypes.POINTER(c_vp), c_int, c_vp, c_int, c_int, c_int, c_int, c_vp,
                             c_sz, c_vp]
    L.ddp_destroy.argtypes = [c_vp]
    L.ddp_set_options.argtypes = [c_vp, c_dbl, c_dbl, c_dbl]
    L.ddp_set_keypoints.argtypes = [c_vp, c_int, c_int, c_int, c_dbl, c_dbl]
    L.ddp_set_cost.argtypes = [c_vp, c_vp, c_vp, c_vp]
    L.ddp_set_target.argtypes = [c_vp, c_vp, c_int]
    L.ddp_set_initial_state.argtypes = [c_vp, c_vp]
    L.ddp_set_initial_guess.argtypes = [c_vp, c_vp]
    L.ddp_reset.argtypes = [c_vp]
    L.ddp_begin_solve.argtypes = [c_vp]
    L.ddp_mpc_shift.argtypes = [c_vp, c_int]
    L.ddp_set_mpc_rearm.argtypes = [c_vp, c_int, c_vp]
    L.ddp_apply_staged_inputs.argtypes = [c_vp, c_vp, c_vp]
    L.ddp_set_control_limits.argtypes = [c_vp, c_vp, c_vp]
    L.ddp_set_regularization.argtypes = [c_vp, c_dbl]
    L.ddp_iterate.argtypes = [c_vp, ip]
    L.ddp_iterate_linesearch.argtypes = [c_vp]
    L.ddp_iterate_finish_async.argtypes = [c_vp]
    L.ddp_iterate_wait.argtypes = [c_vp, ip]
    L.ddp_solve.argtypes = [c_vp, c_int, ip]
    L.ddp_run_phase.argtypes = [c_vp, c_int]
    L.ddp_get.argtypes = [c_vp, c_int, c_vp]
    L.ddp_put.argtypes = [c_vp, c_int, c_vp]
    L.ddp_get_int.argtypes = [c_vp, c_int, c_vp]
    L.ddp_device_ptr.restype = c_vp
    L.ddp_device_ptr.argtypes = [c_vp, c_int]
    L.ddp_array_elems.restype = c_sz
    L.ddp_array_elems.argtypes = [c_vp, c_int]
    L.ddp_last_timings.argtypes = [c_vp, ctypes.POINTER(ctypes.c_float)]
    L.ddp_launch_count.restype = ctypes.c_longlong
    L.ddp_launch_count.argtypes = [c_vp]
    L.ddp_peak_fp64.argtypes = [c_vp, c_int, dp]
    _lib = L
    return L


class DDPError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().ddp_last_error().decode(errors="replace")
        raise DDPError(f"{what} failed with status {rc}: {msg}")
