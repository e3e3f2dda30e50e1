"""One ``MPCProblem`` through the host entry of the C ABI: the path ``solve_mpc`` takes.

``solve_mpc(problem)`` is what a control loop built on the reference calls once per cycle
(``qpmpc/solve_mpc.py:16-44``; ``examples/*.py``), so its latency matters as much as the
batch throughput.  Going through device tensors costs one small host->device copy per operand
and one device->host read per result.  This path instead writes the operands into ONE
page-locked staging block (cached per problem shape) and calls ``qpmpc_b200_solve_host``: a
single kernel launch whose staging reads the operands from host memory and whose epilogue
stores U, the multipliers, the status and the iteration count back into it -- no copy call at
all, one stream synchronisation per solve.

Every operand is declared per instance (a batch of one has nothing to share), so the kernel
reads the caller's block directly.  There is no CPU path here either: the arithmetic happens
in the CUDA library, or the call raises ``BackendError``.
"""

import ctypes
import threading
from typing import Dict, Tuple

import numpy as np
import torch

from . import _capi

_NP = {torch.float64: np.float64, torch.float32: np.float32}
_CODE = {torch.float64: _capi.F64, torch.float32: _capi.F32}


def _rows_paired(pk: dict, nc: int) -> bool:
    """Rows [M; -M] in every step (what ``BatchedMPCProblem`` detects on its tensors)."""
    if nc == 0 or nc % 2 or (pk["C"] is None and pk["D"] is None):
        return False
    h = nc // 2
    for t in (pk["C"], pk["D"]):
        if t is not None and (t[..., :h, :] != -t[..., h:, :]).any():
            return False
    return True


class _Slot:
    """Staging block of one problem shape: operands, outputs and their ctypes views."""

    def __init__(self, shapes: Dict[str, Tuple[int, ...]], n: int, m: int, dtype):
        npdt = _NP[dtype]
        es = np.dtype(npdt).itemsize
        regions, off = {}, 0

        def carve(name, count, itemsize):
            nonlocal off
            regions[name] = (off, count * itemsize)
            off += (count * itemsize + 15) // 16 * 16  # 16-byte aligned: the bulk copies can take them

        for name, shape in shapes.items():
            carve(name, int(np.prod(shape)), es)
        carve("U", n, es)
        carve("Z", max(m, 1), es)
        carve("status", 1, 4)
        carve("iters", 1, 4)
        # page-locked when there is a device (the emulated engine of the CPU tests has none)
        self.block = torch.zeros(max(off, 16), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        base = self.block.numpy()
        self.view, self.ptr = {}, {}
        for name, (o, nbytes) in regions.items():
            kind = np.int32 if name in ("status", "iters") else npdt
            arr = base[o:o + nbytes].view(kind)
            self.view[name] = arr.reshape(shapes[name]) if name in shapes else arr
            self.ptr[name] = ctypes.c_void_p(self.block.data_ptr() + o)
        self.operands = _capi.Operands(*[self.ptr.get(k) for k in ("A", "B", "C", "D", "e", "x0", "goal", "targets")])
        self.outputs = _capi.Outputs(self.ptr["U"], self.ptr["status"], self.ptr["iters"], self.ptr["Z"] if m else None)


_local = threading.local()  # one set of staging blocks per calling thread


_ITEMS = ("A", "B", "C", "D", "e")
_VECS = ("x0", "goal", "targets")


def solve_single(problem, pk: dict, method: int, max_iter: int, tol: float, dtype=torch.float64, polish: bool = True):
    """Solve the packed problem ``pk`` (``batched.pack_problem``).  Returns
    ``(U [n] float64, Z [m] float64, status, iters)``."""
    from .batched import _require_cuda  # (late: the CPU tests replace it)

    device = _require_cuda(None)
    lib = _capi.load()
    N, nx, nu, nc = problem.nb_timesteps, problem.state_dim, problem.input_dim, int(pk["nc"])
    # shape signature: per operand 0 = absent, 1 = one block, 2 = a stack of N blocks
    sig = tuple(0 if pk[k] is None else pk[k].ndim for k in _ITEMS) + tuple(pk[k] is not None for k in _VECS)
    key = (N, nx, nu, nc, sig, dtype)
    slots = _local.__dict__.setdefault("slots", {})
    slot = slots.get(key)
    if slot is None:
        n, m = N * nu, N * nc
        item = dict(A=(nx, nx), B=(nx, nu), C=(nc, nx), D=(nc, nu), e=(nc,))
        shapes, modes = {}, {}
        for name in _ITEMS:
            arr = pk[name]
            if arr is None or 0 in item[name]:
                modes[name] = _capi.ABSENT
                continue
            ltv = arr.ndim == len(item[name]) + 1
            shapes[name] = ((N,) if ltv else ()) + item[name]
            modes[name] = _capi.BATCH_LTV if ltv else _capi.BATCH_LTI
        for name, size in (("x0", nx), ("goal", nx), ("targets", N * nx)):
            if pk[name] is not None:
                shapes[name] = (size,)
        slot = slots[key] = _Slot(shapes, n, m, dtype)
        slot.names, slot.m = tuple(shapes), m
        d = slot.desc = _capi.Desc()
        d.batch, d.N, d.nx, d.nu, d.nc = 1, N, nx, nu, nc
        d.dtype = _CODE[dtype]
        d.mode_A, d.mode_B, d.mode_C = modes["A"], modes["B"], modes["C"]
        d.mode_D, d.mode_e = modes["D"], modes["e"]
        d.mode_x0 = _capi.VEC_BATCH
        d.mode_goal = _capi.VEC_BATCH if "goal" in shapes else _capi.VEC_ABSENT
        d.mode_targets = _capi.VEC_BATCH if "targets" in shapes else _capi.VEC_ABSENT
        slot.call = (ctypes.byref(d), ctypes.byref(slot.operands), ctypes.byref(slot.outputs), device.index or 0)
    view = slot.view
    for name in slot.names:
        dst = view[name]
        dst[...] = np.asarray(pk[name]).reshape(dst.shape)
    d = slot.desc
    d.has_wt = problem.terminal_cost_weight is not None
    d.has_wx = problem.stage_state_cost_weight is not None
    d.w_t = float(problem.terminal_cost_weight or 0.0)
    d.w_x = float(problem.stage_state_cost_weight or 0.0)
    d.w_u = float(problem.stage_input_cost_weight)
    d.method, d.max_iter, d.tol = method, int(max_iter), float(tol)
    d.flags = 0 if polish else _capi.FLAG_NO_POLISH
    d.paired = int(_rows_paired(pk, nc))
    rc = lib.qpmpc_b200_solve_host(*slot.call)
    if rc:
        _capi.check(rc, "qpmpc_b200_solve_host")
    return (np.array(view["U"], dtype=np.float64), np.array(view["Z"][:slot.m], dtype=np.float64),
            int(view["status"][0]), int(view["iters"][0]))
