"""Host emulator of the engine's warp kernels.  TEST INFRASTRUCTURE ONLY.

``kernel_emu.cpp`` compiles the DEVICE SOURCE (``qpmpc_b200/csrc/mpc_kernels.cuh``,
``mpc_pdip.cuh``) for the host against ``warp_emu.h`` (one fiber per CUDA
thread) and drives it with the product's own host-side parameter code
(``mpc_host_params.h``, ``layout_smem``).  This module builds the library on
demand and wraps its entry points; ``tests/test_kernel_emu.py`` and
``tests/test_pdip_emu.py`` compare the results with the oracle.  Nothing under
``qpmpc_b200/`` imports this, nothing here is shipped or timed.
"""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB_PATH = os.path.join(HERE, "libkernel_emu.so")
_vp = ctypes.c_void_p
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_lib = None


def _declare(lib):
    for name in ("emu_divergent_collectives", "emu_selftest", "emu_sync_points", "emu_sync_points_at"):
        getattr(lib, name).restype = ctypes.c_long
    return lib


def load():
    """Build (if stale) and load ``libkernel_emu.so``."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("QPMPC_EMU_LIB"):  # another build of kernel_emu.cpp, e.g. -march=native -ffp-contract=fast
        _lib = _declare(ctypes.CDLL(os.environ["QPMPC_EMU_LIB"]))
        return _lib
    csrc = os.path.join(ROOT, "qpmpc_b200", "csrc")
    deps = [os.path.join(HERE, "kernel_emu.cpp"), os.path.join(HERE, "warp_emu.h"),
            os.path.join(ROOT, "include", "qpmpc_b200.h")]
    deps += [os.path.join(csrc, f) for f in ("mpc_common.cuh", "mpc_kernels.cuh", "mpc_pdip.cuh",
                                             "mpc_cta_kernel.cuh", "mpc_lr_kernel.cuh", "mpc_integrate.cuh", "mpc_plant.cuh",
                                             "mpc_host_params.h")]
    if not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DQPMPC_HOST_EMU",
                        "-Wno-unknown-pragmas", f"-I{ROOT}", "-o", LIB_PATH, deps[0]], check=True)
    _lib = _declare(ctypes.CDLL(LIB_PATH))
    return _lib


class _Checked:
    """Fails the test if the emulated launch stored past its shared-memory request
    or entered a warp-level collective with only part of the lanes it names."""

    def __enter__(self):
        lib = load()
        self.before = (lib.emu_smem_overruns(), lib.emu_divergent_collectives())
        return lib

    def __exit__(self, *exc):
        lib = load()
        lib.emu_set_lane_order(0)
        if exc[0] is None:
            assert lib.emu_smem_overruns() == self.before[0], "a CTA wrote past its shared-memory request"
            assert lib.emu_divergent_collectives() == self.before[1], \
                "a shuffle / vote / reduction was entered by only part of the lanes it names (deadlock on the device)"
        return False


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def describe(w, method=0, max_iter=0, tol=0.0, flags=0, dtype=np.float64, paired=None):
    """(desc, operands, keep-alive arrays) of a workload dict, host pointers:
    the structs of include/qpmpc_b200.h filled the way BatchedMPCProblem fills them."""
    from qpmpc_b200 import _capi
    from qpmpc_b200.workloads import operand_layout

    d, keep = _capi.Desc(), {}
    d.batch, d.N, d.nx, d.nu, d.nc = w["batch"], w["N"], w["nx"], w["nu"], w["nc"]
    d.dtype = 0 if dtype == np.float64 else 1
    for name in ("A", "B", "C", "D", "e"):
        arr, mode = w[name], 0
        if arr is not None:
            mode = {(False, False): 1, (False, True): 2, (True, False): 3, (True, True): 4}[operand_layout(w, name)]
            keep[name] = np.ascontiguousarray(arr, dtype=dtype)
        setattr(d, "mode_" + name, mode)
    for name in ("x0", "goal", "targets"):
        arr, mode = w[name], 0
        if arr is not None:
            mode = 2 if arr.ndim == 2 else 1
            keep[name] = np.ascontiguousarray(arr, dtype=dtype)
        setattr(d, "mode_" + name, mode)
    d.has_wt, d.has_wx = w["w_t"] is not None, w["w_x"] is not None
    d.w_t, d.w_x, d.w_u = float(w["w_t"] or 0.0), float(w["w_x"] or 0.0), w["w_u"]
    d.method, d.max_iter, d.tol, d.flags = method, max_iter, tol, flags
    if paired is None:
        from qpmpc_b200.workloads import rows_are_paired

        paired = rows_are_paired(w)
    d.paired = int(bool(paired))
    ops = _capi.Operands(*[_ptr(keep.get(k)) for k in ("A", "B", "C", "D", "e", "x0", "goal", "targets")])
    return d, ops, keep


def solve(w, method="active_set", wpc=0, max_iter=0, tol=0.0, polish=True, descending=False, dtype=np.float64,
          paired=None):
    """qpmpc_b200_solve on the emulator: mpc_solve_kernel / mpc_solve_cta_kernel /
    mpc_pdip_kernel, in ``dtype`` (float64 or float32; U and z come back as float64)."""
    from qpmpc_b200 import _capi

    lib = load()
    meth = {"active_set": _capi.ACTIVE_SET, "pdip": _capi.PDIP}[method]
    d, ops, keep = describe(w, meth, max_iter, tol, 0 if polish else _capi.FLAG_NO_POLISH, dtype, paired)
    B, n, m = w["batch"], w["N"] * w["nu"], w["N"] * w["nc"]
    U, Z = np.zeros((B, n), dtype=dtype), np.zeros((B, max(m, 1)), dtype=dtype)
    st, it = np.full(B, -1, np.int32), np.zeros(B, np.int32)
    outs = _capi.Outputs(_ptr(U), _ptr(st), _ptr(it), _ptr(Z))
    with _Checked() as lib:
        lib.emu_set_lane_order(int(descending))
        rc = lib.emu_solve(ctypes.byref(d), ctypes.byref(ops), ctypes.byref(outs), wpc)
    return dict(rc=rc, U=U.astype(np.float64), status=st, iters=it, z=Z[:, :m].astype(np.float64))


def condense(w, fields=("P", "q", "G", "h", "Phi", "Psi", "phi_last", "psi_last")):
    """qpmpc_b200_condense on the emulator: mpc_condense_kernel."""
    from qpmpc_b200 import _capi

    lib = load()
    d, ops, keep = describe(w)
    B, N, nx, n, m = w["batch"], w["N"], w["nx"], w["N"] * w["nu"], w["N"] * w["nc"]
    shapes = dict(P=(B, n, n), q=(B, n), G=(B, m, n), h=(B, m), Phi=(B, N * nx, nx), Psi=(B, N * nx, n),
                  phi_last=(B, nx, nx), psi_last=(B, nx, n))
    out = {k: np.full(shapes[k], np.nan) for k in fields}
    qf = _capi.QPFields(*[_ptr(out.get(k)) for k in ("P", "q", "G", "h", "Phi", "Psi", "phi_last", "psi_last")])
    with _Checked() as lib:
        rc = lib.emu_condense(ctypes.byref(d), ctypes.byref(ops), ctypes.byref(qf))
    assert rc == 0, rc
    return out


def pdip_core(P, q, G, h, np_, mr, dtype=0, max_iter=50, tol=1e-9, polish=True):
    """pdip_core() on explicit QPs: one emulated warp, up to 32 // np_ QPs side by side."""
    lib = load()
    P, q, G, h = (np.ascontiguousarray(a, dtype=np.float64) for a in (P, q, G, h))
    B, m, n = G.shape
    U, Z = np.zeros((B, n)), np.zeros((B, max(m, 1)))
    st, it = np.zeros(B, np.int32), np.zeros(B, np.int32)
    pmin = float(np.linalg.eigvalsh(P).min())  # the kernel gets w_u <= lambda_min(P) from the MPC problem
    with _Checked() as lib:
        rc = lib.pdip_emu_solve(dtype, np_, mr, B, n, m, P.ctypes.data_as(_dp), q.ctypes.data_as(_dp),
                                G.ctypes.data_as(_dp), h.ctypes.data_as(_dp), max_iter, ctypes.c_double(tol),
                                int(polish), ctypes.c_double(pmin), U.ctypes.data_as(_dp), Z.ctypes.data_as(_dp),
                                st.ctypes.data_as(_ip), it.ctypes.data_as(_ip))
    assert rc == 0, rc
    return dict(U=U, z=Z[:, :m], status=st, iters=it)


def integrate(w, U):
    """qpmpc_b200_integrate on the emulator (mpc_integrate_kernel): X [B, N+1, nx]."""
    d, ops, keep = describe(w)
    U = np.ascontiguousarray(U, dtype=np.float64).reshape(w["batch"], -1)
    X = np.full((w["batch"], w["N"] + 1, w["nx"]), np.nan)
    with _Checked() as lib:
        rc = lib.emu_integrate(ctypes.byref(d), ctypes.byref(ops), _ptr(U), _ptr(X))
    assert rc == 0, rc
    return X


def pendulum_closed_loop(w, cycles, substeps=15, length=0.6, gravity=9.81, method="active_set"):
    """qpmpc_b200_pendulum_closed_loop on the emulator: pendulum_step_kernel and the
    fused solve alternate; returns (trajectory [cycles+1, B, 4], unsolved count)."""
    from qpmpc_b200 import _capi

    B, N = w["batch"], w["N"]
    w = dict(w, goal=np.zeros((B, 4)), targets=np.zeros((B, N * 4)), x0=w["x0"].copy())
    meth = {"active_set": _capi.ACTIVE_SET, "pdip": _capi.PDIP}[method]
    d, ops, keep = describe(w, meth)
    U, st, it = np.zeros((B, N)), np.zeros(B, np.int32), np.zeros(B, np.int32)
    outs = _capi.Outputs(_ptr(U), _ptr(st), _ptr(it), None)
    v = np.ascontiguousarray(w["v_target"], dtype=np.float64)
    traj = np.full((cycles + 1, B, 4), np.nan)
    unsolved = np.zeros(1, np.int32)
    loop = _capi.ClosedLoop(int(cycles), int(substeps), w["T"] / substeps, w["T"], length, gravity,
                            _ptr(v), _ptr(traj), _ptr(unsolved))
    with _Checked() as lib:
        rc = lib.emu_pendulum_closed_loop(ctypes.byref(d), ctypes.byref(ops), ctypes.byref(outs), ctypes.byref(loop))
    assert rc == 0, rc
    return traj, int(unsolved[0])


class EmulatedLibrary:
    """Stands in for ``libqpmpc_b200.so`` in CPU tests of the HOST-side surface: the entry points
    ``qpmpc_b200/batched.py`` calls, served by the device source compiled for the host (this
    module) on CPU tensors.  Test infrastructure: installed by the ``emulated_engine`` fixture of
    ``tests/conftest.py`` only; the product never imports it."""

    def __init__(self):
        self.lib = load()
        self.calls = 0

    def qpmpc_b200_solve(self, desc, ops, outs, stream):
        self.calls += 1
        with _Checked() as lib:
            return lib.emu_solve(desc, ops, outs, 0)

    def qpmpc_b200_solve_host(self, desc, ops, outs, device):
        # host buffers are all the emulator has: the host entry is the device entry
        return self.qpmpc_b200_solve(desc, ops, outs, None)

    def qpmpc_b200_condense(self, desc, ops, fields, stream):
        self.calls += 1
        with _Checked() as lib:
            return lib.emu_condense(desc, ops, fields)

    def qpmpc_b200_integrate(self, desc, ops, U, X, stream):
        self.calls += 1
        with _Checked() as lib:
            return lib.emu_integrate(desc, ops, U, X)

    def qpmpc_b200_factor_bytes(self, desc):
        self.lib.emu_factor_bytes.restype = ctypes.c_size_t
        return self.lib.emu_factor_bytes(desc)

    def qpmpc_b200_factor(self, desc, ops, record, stream):
        self.calls += 1
        with _Checked() as lib:
            return lib.emu_factor(desc, ops, record)

    def qpmpc_b200_solve_factored(self, desc, ops, record, outs, stream):
        self.calls += 1
        with _Checked() as lib:
            return lib.emu_solve_factored(desc, ops, record, outs, 0)

    def qpmpc_b200_pendulum_closed_loop(self, desc, ops, outs, loop, stream):
        self.calls += 1
        with _Checked() as lib:
            return lib.emu_pendulum_closed_loop(desc, ops, outs, loop)

    def qpmpc_b200_lipm_closed_loop(self, desc, ops, outs, loop, stream):
        self.calls += 1
        with _Checked() as lib:
            return lib.emu_lipm_closed_loop(desc, ops, outs, loop)

    def qpmpc_b200_strerror(self, code):
        return f"emulated engine: error {code}".encode()

    def qpmpc_b200_launch_count(self):
        return self.calls
