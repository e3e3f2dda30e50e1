"""ctypes binding of ``include/qpmpc_b200.h`` (the C-ABI CUDA library).

There is no CPU fallback: if the library is missing, or a solve is requested
without a CUDA device, a :class:`BackendError` is raised.
"""

import ctypes
import os

from .exceptions import BackendError

_HERE = os.path.dirname(os.path.abspath(__file__))
# QPMPC_B200_LIB selects another build of the same library (A/B measurements)
LIB_PATH = os.environ.get("QPMPC_B200_LIB") or os.path.join(_HERE, "lib", "libqpmpc_b200.so")

ABSENT, SHARED_LTI, SHARED_LTV, BATCH_LTI, BATCH_LTV = range(5)
VEC_ABSENT, VEC_SHARED, VEC_BATCH = range(3)
F64, F32 = 0, 1
ACTIVE_SET, PDIP = 0, 1
FLAG_NO_POLISH = 1  # desc.flags: PDIP without the active-set polish
STATUS_SOLVED, STATUS_MAX_ITER, STATUS_INFEASIBLE, STATUS_NOT_SPD = range(4)

EXPORTS = (
    "qpmpc_b200_solve", "qpmpc_b200_solve_host", "qpmpc_b200_condense",
    "qpmpc_b200_integrate", "qpmpc_b200_workspace_bytes", "qpmpc_b200_max_vars",
    "qpmpc_b200_max_rows", "qpmpc_b200_launch_count", "qpmpc_b200_strerror",
    "qpmpc_b200_version", "qpmpc_b200_fp64_peak", "qpmpc_b200_pendulum_closed_loop",
    "qpmpc_b200_solve_scatter", "qpmpc_b200_factor_bytes", "qpmpc_b200_factor",
    "qpmpc_b200_solve_factored", "qpmpc_b200_lipm_closed_loop",
)


class Desc(ctypes.Structure):
    """``qpmpc_b200_desc``."""

    _fields_ = [
        ("batch", ctypes.c_int32), ("N", ctypes.c_int32),
        ("nx", ctypes.c_int32), ("nu", ctypes.c_int32), ("nc", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("mode_A", ctypes.c_int32), ("mode_B", ctypes.c_int32),
        ("mode_C", ctypes.c_int32), ("mode_D", ctypes.c_int32),
        ("mode_e", ctypes.c_int32),
        ("mode_x0", ctypes.c_int32), ("mode_goal", ctypes.c_int32),
        ("mode_targets", ctypes.c_int32),
        ("has_wt", ctypes.c_int32), ("has_wx", ctypes.c_int32),
        ("w_t", ctypes.c_double), ("w_x", ctypes.c_double), ("w_u", ctypes.c_double),
        ("method", ctypes.c_int32), ("max_iter", ctypes.c_int32),
        ("tol", ctypes.c_double),
        ("paired", ctypes.c_int32), ("flags", ctypes.c_int32),
    ]


class Operands(ctypes.Structure):
    """``qpmpc_b200_operands``."""

    _fields_ = [(k, ctypes.c_void_p) for k in
                ("A", "B", "C", "D", "e", "x0", "goal", "targets")]


class Outputs(ctypes.Structure):
    """``qpmpc_b200_outputs``."""

    _fields_ = [(k, ctypes.c_void_p) for k in ("U", "status", "iters", "Z")]


class QPFields(ctypes.Structure):
    """``qpmpc_b200_qp_fields``."""

    _fields_ = [(k, ctypes.c_void_p) for k in
                ("P", "q", "G", "h", "Phi", "Psi", "phi_last", "psi_last")]


class Peers(ctypes.Structure):
    """``qpmpc_b200_peers``."""

    _fields_ = [
        ("count", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("row_offset", ctypes.c_int64),
        ("U", ctypes.c_void_p * 8), ("status", ctypes.c_void_p * 8),
    ]


class ClosedLoop(ctypes.Structure):
    """``qpmpc_b200_closed_loop``."""

    _fields_ = [
        ("cycles", ctypes.c_int32), ("substeps", ctypes.c_int32),
        ("dt", ctypes.c_double), ("sampling_period", ctypes.c_double),
        ("length", ctypes.c_double), ("gravity", ctypes.c_double),
        ("v_target", ctypes.c_void_p), ("trajectory", ctypes.c_void_p),
        ("unsolved", ctypes.c_void_p), ("record", ctypes.c_void_p),
        ("upright", ctypes.c_void_p), ("iterations", ctypes.c_void_p),
    ]


class LipmLoop(ctypes.Structure):
    """``qpmpc_b200_lipm_loop``."""

    _fields_ = [
        ("cycles", ctypes.c_int32), ("substeps", ctypes.c_int32),
        ("nb_dsp_steps", ctypes.c_int32), ("nb_ssp_steps", ctypes.c_int32),
        ("sampling_period", ctypes.c_double), ("foot_size", ctypes.c_double),
        ("max_zmp_dist", ctypes.c_double),
        ("support_foot", ctypes.c_void_p), ("strides", ctypes.c_void_p),
        ("phase_index", ctypes.c_void_p), ("stride_index", ctypes.c_void_p),
        ("trajectory", ctypes.c_void_p), ("unsolved", ctypes.c_void_p),
        ("record", ctypes.c_void_p),
    ]


_lib = None


def load():
    """Load the library once; raises BackendError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BackendError(
            f"{LIB_PATH} is missing: build it with `python -m qpmpc_b200.build` "
            "(there is no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    P = ctypes.POINTER
    lib.qpmpc_b200_solve.argtypes = [P(Desc), P(Operands), P(Outputs), ctypes.c_void_p]
    lib.qpmpc_b200_solve_host.argtypes = [P(Desc), P(Operands), P(Outputs), ctypes.c_int]
    lib.qpmpc_b200_condense.argtypes = [P(Desc), P(Operands), P(QPFields), ctypes.c_void_p]
    lib.qpmpc_b200_integrate.argtypes = [
        P(Desc), P(Operands), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.qpmpc_b200_pendulum_closed_loop.argtypes = [
        P(Desc), P(Operands), P(Outputs), P(ClosedLoop), ctypes.c_void_p]
    lib.qpmpc_b200_solve_scatter.argtypes = [
        P(Desc), P(Operands), P(Outputs), P(Peers), ctypes.c_void_p]
    lib.qpmpc_b200_factor_bytes.argtypes = [P(Desc)]
    lib.qpmpc_b200_factor_bytes.restype = ctypes.c_size_t
    lib.qpmpc_b200_factor.argtypes = [P(Desc), P(Operands), ctypes.c_void_p, ctypes.c_void_p]
    lib.qpmpc_b200_solve_factored.argtypes = [P(Desc), P(Operands), ctypes.c_void_p, P(Outputs), ctypes.c_void_p]
    lib.qpmpc_b200_lipm_closed_loop.argtypes = [
        P(Desc), P(Operands), P(Outputs), P(LipmLoop), ctypes.c_void_p]
    lib.qpmpc_b200_lipm_closed_loop.restype = ctypes.c_int
    for name in ("solve", "solve_host", "condense", "integrate", "version", "factor", "solve_factored",
                 "max_vars", "max_rows", "pendulum_closed_loop", "solve_scatter"):
        getattr(lib, f"qpmpc_b200_{name}").restype = ctypes.c_int
    lib.qpmpc_b200_max_vars.argtypes = [ctypes.c_int]
    lib.qpmpc_b200_max_rows.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.qpmpc_b200_workspace_bytes.argtypes = [P(Desc)]
    lib.qpmpc_b200_workspace_bytes.restype = ctypes.c_size_t
    lib.qpmpc_b200_launch_count.restype = ctypes.c_longlong
    lib.qpmpc_b200_fp64_peak.argtypes = [ctypes.c_int, P(ctypes.c_double)]
    lib.qpmpc_b200_fp64_peak.restype = ctypes.c_int
    lib.qpmpc_b200_strerror.argtypes = [ctypes.c_int]
    lib.qpmpc_b200_strerror.restype = ctypes.c_char_p
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    """Turn a non-zero return code of the C ABI into a BackendError."""
    if code != 0:
        msg = load().qpmpc_b200_strerror(code).decode()
        raise BackendError(f"{what} failed ({code}): {msg}")


def launch_count() -> int:
    return int(load().qpmpc_b200_launch_count())
