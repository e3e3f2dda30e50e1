"""ctypes front-end of ``mpc_oracle.c`` (test infrastructure, see package doc)."""

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmpc_oracle.so")
_lib = None

STATUS_OK = 0
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile; returns the .so path."""
    src = os.path.join(_HERE, "mpc_oracle.c")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    )
    if stale:
        subprocess.run(["make", "-s", "-C", _HERE, "-B"], check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_qp_gi.restype = ctypes.c_int
        _lib.oracle_solve_batch.restype = ctypes.c_int
        _lib.oracle_condense.restype = ctypes.c_int
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_kkt.restype = None
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a) -> Optional[np.ndarray]:
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def num_threads() -> int:
    return _load().oracle_num_threads()


def qp_gi(P, q, G, h):
    """Solve one QP exactly; returns (status, x, z, iters)."""
    lib = _load()
    P, q, G, h = _f64(P), _f64(q), _f64(G), _f64(h)
    n, m = q.size, h.size
    x, z = np.zeros(n), np.zeros(max(m, 1))
    it = ctypes.c_int(0)
    st = lib.oracle_qp_gi(
        ctypes.c_int(n), ctypes.c_int(m), _ptr(P), _ptr(q), _ptr(G), _ptr(h),
        _ptr(x), _ptr(z), ctypes.byref(it),
    )
    return st, x, z[:m], it.value


def kkt(P, q, G, h, x, z) -> np.ndarray:
    """[stationarity, primal, dual, complementarity] residuals of (x, z)."""
    lib = _load()
    P, q, G, h, x, z = map(_f64, (P, q, G, h, x, z))
    out = np.zeros(4)
    lib.oracle_kkt(
        ctypes.c_int(q.size), ctypes.c_int(h.size), _ptr(P), _ptr(q), _ptr(G),
        _ptr(h), _ptr(x), _ptr(z), _ptr(out),
    )
    return out


def _strides(arr: Optional[np.ndarray], per_instance: bool, per_step: bool, item: int):
    """(batch stride, step stride) in elements for a canonical operand."""
    if arr is None:
        return 0, 0
    step = item if per_step else 0
    inst = (arr.size // arr.shape[0]) if per_instance else 0
    return inst, step


def condense(N, nx, nu, nc, A, B, C, D, e, x0, goal, targets, w_t, w_x, w_u):
    """Condense ONE instance.  A..e are (r, c) arrays (LTI) or (N, r, c) (LTV)."""
    lib = _load()
    A, B, C, D, e, x0, goal, targets = map(_f64, (A, B, C, D, e, x0, goal, targets))
    n, m = N * nu, N * nc
    out = dict(
        P=np.zeros((n, n)), q=np.zeros(n), G=np.zeros((m, n)), h=np.zeros(m),
        Phi=np.zeros((N * nx, nx)), Psi=np.zeros((N * nx, n)),
        phi_last=np.zeros((nx, nx)), psi_last=np.zeros((nx, n)),
    )

    def step(a, ltv_ndim, item):
        return 0 if a is None or a.ndim < ltv_ndim else item

    lib.oracle_condense(
        ctypes.c_int(N), ctypes.c_int(nx), ctypes.c_int(nu), ctypes.c_int(nc),
        _ptr(A), ctypes.c_long(step(A, 3, nx * nx)),
        _ptr(B), ctypes.c_long(step(B, 3, nx * nu)),
        _ptr(C), ctypes.c_long(step(C, 3, nc * nx)),
        _ptr(D), ctypes.c_long(step(D, 3, nc * nu)),
        _ptr(e), ctypes.c_long(step(e, 2, nc)),
        _ptr(x0), _ptr(goal), _ptr(targets),
        ctypes.c_int(w_t is not None), ctypes.c_double(w_t or 0.0),
        ctypes.c_int(w_x is not None), ctypes.c_double(w_x or 0.0),
        ctypes.c_double(w_u),
        *[_ptr(out[k]) for k in ("P", "q", "G", "h", "Phi", "Psi", "phi_last", "psi_last")],
    )
    return out


def solve_batch(
    batch, N, nx, nu, nc, ops, w_t, w_x, w_u, want_kkt=False, nthreads=0
):
    """Condense + solve a batch.

    ``ops`` maps operand name -> (array or None, per_instance, per_step) for
    A, B, C, D, e and -> (array or None, per_instance) for x0, goal, targets.
    Returns dict(U, status, iters, kkt).
    """
    lib = _load()
    n = N * nu
    items = dict(A=nx * nx, B=nx * nu, C=nc * nx, D=nc * nu, e=nc)
    args = [ctypes.c_int(v) for v in (batch, N, nx, nu, nc)]
    keep = []
    for name in ("A", "B", "C", "D", "e"):
        arr, per_inst, per_step = ops[name]
        arr = _f64(arr)
        keep.append(arr)
        item = items[name]
        bs = (item * (N if per_step else 1)) if (per_inst and arr is not None) else 0
        ss = item if (per_step and arr is not None) else 0
        args += [_ptr(arr), ctypes.c_long(bs), ctypes.c_long(ss)]
    for name, size in (("x0", nx), ("goal", nx), ("targets", N * nx)):
        arr, per_inst = ops[name]
        arr = _f64(arr)
        keep.append(arr)
        args += [_ptr(arr), ctypes.c_long(size if (per_inst and arr is not None) else 0)]
    U = np.zeros((batch, n))
    status = np.zeros(batch, dtype=np.int32)
    iters = np.zeros(batch, dtype=np.int32)
    kk = np.zeros((batch, 4)) if want_kkt else None
    args += [
        ctypes.c_int(w_t is not None), ctypes.c_double(w_t or 0.0),
        ctypes.c_int(w_x is not None), ctypes.c_double(w_x or 0.0),
        ctypes.c_double(w_u),
        _ptr(U), status.ctypes.data_as(_ip), iters.ctypes.data_as(_ip), _ptr(kk),
        ctypes.c_int(nthreads),
    ]
    bad = lib.oracle_solve_batch(*args)
    return dict(U=U, status=status, iters=iters, kkt=kk, bad=bad)
