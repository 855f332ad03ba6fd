"""ctypes binding of libmirkb200.so — the C ABI declared in include/mirk_b200.h.

This is the same boundary a Julia `ccall` backend binds.  There is no CPU
fallback: if the shared library is missing the import fails loudly, and without a CUDA device
`mirk_create` returns MIRK_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmirkb200.so")
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
fp = C.POINTER(C.c_float)

OK, ERR_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE = 0, -1, -2, -3, -4, -5
RET_SUCCESS, RET_FAILURE, RET_MAXITERS, RET_UNSTABLE, RET_STALLED = 0, 1, 2, 3, 4
RETCODE_NAMES = {0: "Success", 1: "Failure", 2: "MaxIters", 3: "Unstable", 4: "Stalled"}


class Desc(C.Structure):
    _fields_ = [("problem_id", C.c_int32), ("order", C.c_int32), ("abstol", C.c_double),
                ("adaptive", C.c_int32), ("defect_threshold", C.c_double),
                ("max_num_subintervals", C.c_int32), ("maxiters", C.c_int32),
                ("reinterp_inplace", C.c_int32), ("chunk", C.c_int32), ("device", C.c_int32),
                ("n_params", C.c_int32), ("params", dp), ("nlsolve", C.c_int32),
                ("controller", C.c_int32), ("ge_method", C.c_int32), ("DE", C.c_double), ("GE", C.c_double)]


class ProblemInfo(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_params", C.c_int32), ("problem_type", C.c_int32),
                ("n_bc", C.c_int32), ("n_bca", C.c_int32), ("max_bc_pts", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("retcode", C.c_int32), ("n_mesh", C.c_int32), ("outer_iters", C.c_int32),
                ("newton_iters", C.c_int32), ("resid_norm", C.c_double), ("defect_norm", C.c_double),
                ("n_hist", C.c_int32), ("hist_n_mesh", C.c_int32 * 64), ("hist_newton", C.c_int32 * 64),
                ("hist_defect", C.c_double * 64)]


class EnsembleDesc(C.Structure):
    _fields_ = [("problem_id", C.c_int32), ("order", C.c_int32), ("abstol", C.c_double),
                ("adaptive", C.c_int32), ("defect_threshold", C.c_double),
                ("max_num_subintervals", C.c_int32), ("maxiters", C.c_int32),
                ("reinterp_inplace", C.c_int32), ("device", C.c_int32), ("node_cap", C.c_int32),
                ("t0", C.c_double), ("t1", C.c_double), ("dt", C.c_double), ("nlsolve", C.c_int32)]


Handle = C.c_void_p
i64 = C.c_int64

# every symbol include/mirk_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mirk_version": (C.c_int, []),
    "mirk_last_error": (C.c_char_p, []),
    "mirk_device_count": (C.c_int, [ip]),
    "mirk_problem_lookup": (C.c_int, [C.c_char_p, ip]),
    "mirk_problem_info_get": (C.c_int, [C.c_int32, C.POINTER(ProblemInfo)]),
    "mirk_problem_register_plugin": (C.c_int, [C.c_char_p, C.c_char_p, ip]),
    "mirk_mesh_uniform": (C.c_int, [C.c_double, C.c_double, C.c_int32, dp]),
    "mirk_create": (C.c_int, [C.POINTER(Desc), C.POINTER(Handle)]),
    "mirk_destroy": (C.c_int, [Handle]),
    "mirk_set_params": (C.c_int, [Handle, dp, C.c_int32]),
    "mirk_set_mesh_guess": (C.c_int, [Handle, C.c_int32, dp, dp]),
    "mirk_set_mesh_guess_device": (C.c_int, [Handle, C.c_int32, dp, C.c_void_p]),
    "mirk_get_solution_device": (C.c_int, [Handle, C.c_void_p]),
    "mirk_ensemble_set_inputs_device": (C.c_int, [Handle, C.c_void_p, C.c_void_p, C.c_int32]),
    "mirk_set_uniform_guess": (C.c_int, [Handle, C.c_double, C.c_double, C.c_double, dp]),
    "mirk_residual": (C.c_int, [Handle, dp, dp]),
    "mirk_jacobian_blocks": (C.c_int, [Handle, dp, dp, ip, dp, ip]),
    "mirk_linear_solve": (C.c_int, [Handle, dp]),
    "mirk_abd_solve": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp, dp, C.c_int32, ip, dp, dp, dp, C.c_int32]),
    "mirk_newton_step": (C.c_int, [Handle, dp]),
    "mirk_newton_solve": (C.c_int, [Handle, ip, dp]),
    "mirk_nlsolve_stats": (C.c_int, [Handle, ip, ip]),
    "mirk_defect": (C.c_int, [Handle, dp, dp]),
    "mirk_refine_mesh": (C.c_int, [Handle, ip]),
    "mirk_mesh_select": (C.c_int, [C.c_int32, C.c_double, C.c_int32, C.c_int32, dp, dp, ip, dp, C.c_int32]),
    "mirk_solve": (C.c_int, [Handle, C.POINTER(Result)]),
    "mirk_get_mesh_size": (C.c_int, [Handle, ip]),
    "mirk_get_solution": (C.c_int, [Handle, dp, dp]),
    "mirk_get_stages": (C.c_int, [Handle, dp, dp]),
    "mirk_get_residual": (C.c_int, [Handle, dp]),
    "mirk_interp": (C.c_int, [Handle, dp, C.c_int32, C.c_int32, dp]),
    "mirk_bench_newton_steps": (C.c_int, [Handle, C.c_int32, fp, fp, C.POINTER(C.c_int64)]),
    "mirk_measure_peaks": (C.c_int, [C.c_int32, dp, dp]),
    "mirk_nccl_unique_id": (C.c_int, [C.c_void_p, C.c_char_p]),
    "mirk_partition_attach": (C.c_int, [Handle, C.c_int32, C.c_int32, C.c_void_p, C.c_char_p]),
    "mirk_partition_p2p_export": (C.c_int, [Handle, C.c_int32, C.c_int32, C.c_void_p]),
    "mirk_partition_attach_p2p": (C.c_int, [Handle, C.c_int32, C.c_int32, C.c_void_p]),
    "mirk_ensemble_create": (C.c_int, [C.POINTER(EnsembleDesc), C.c_int64, C.POINTER(Handle)]),
    "mirk_ensemble_destroy": (C.c_int, [Handle]),
    "mirk_ensemble_set_inputs": (C.c_int, [Handle, dp, dp, C.c_int32]),
    "mirk_ensemble_run": (C.c_int, [Handle, fp]),
    "mirk_ensemble_get_results": (C.c_int, [Handle, ip, ip, ip, ip, dp, dp, dp]),
    "mirk_ensemble_node_cap": (C.c_int, [Handle, ip]),
    "mirk_ensemble_get_trajectory": (C.c_int, [Handle, C.c_int64, ip, dp, dp]),
    "mirk_ensemble_solve": (C.c_int, [C.POINTER(EnsembleDesc), C.c_int64, dp, dp, C.c_int32, ip, ip, ip, dp]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libmirkb200.so in-tree with csrc/Makefile (nvcc, sm_100a only)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libmirkb200.so failed")
    return LIB_PATH


def lib():
    """The loaded C ABI.  Raises if the CUDA library was not built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(this package has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the header and the library ever drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class MirkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmirkb200 error {code}: {msg}")
        self.code = code


def check(code: int) -> int:
    """Raise on argument / environment errors (< 0); numerical outcomes (>= 0) are returned."""
    if code < 0:
        msg = lib().mirk_last_error().decode()
        if code == ERR_ARG and "dt must be positive" in msg:
            raise ValueError("dt must be positive")  # ArgumentError in the reference, CORE/utils.jl:354
        raise MirkError(code, msg)
    return code
