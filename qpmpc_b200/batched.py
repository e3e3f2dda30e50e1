"""Tensor-native batched surface: many independent MPC problems in one call.

The reference solves one problem per ``solve_mpc`` call (``qpmpc/solve_mpc.py:
42-44``, reference tree).  Here a :class:`BatchedMPCProblem` holds B problems
of one shape as CUDA tensors and :func:`solve_mpc_batch` runs the fused
condense + QP kernel on all of them; :class:`BatchedPlan` is the batched
counterpart of ``Plan`` (``qpmpc/plan.py:18-109``).

Operand layouts (see ``include/qpmpc_b200.h``): every matrix operand is either
shared by the batch or per instance, and either time-invariant or per step:

    ndim 2  [r, c]          shared,       time-invariant
    ndim 3  [B, r, c]       per instance, time-invariant   (default for ndim 3)
    ndim 3  [N, r, c]       shared,       per step         (name listed in ``ltv``)
    ndim 4  [B, N, r, c]    per instance, per step

``ineq_vector`` follows the same rule with r = nc and no trailing axis.
"""

import ctypes
import os
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _capi
from .exceptions import BackendError, ProblemDefinitionError, StateError
from .mpc_problem import MPCProblem

_TORCH_DTYPE = {_capi.F64: torch.float64, _capi.F32: torch.float32}
_DTYPE_CODE = {torch.float64: _capi.F64, torch.float32: _capi.F32}


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise BackendError(
            "the qpmpc_b200 engine needs a CUDA device (sm_100a); there is no CPU path"
        )
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _device_guard(device):
    """Context that makes ``device`` current for the launch."""
    return torch.cuda.device(device)


def _stream_ptr(device):
    """The caller's current CUDA stream on ``device`` as a ``cudaStream_t``."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class BatchedMPCProblem:
    """B linear MPC problems sharing (N, nx, nu, nc), stored as tensors.

    Argument names follow ``MPCProblem`` (``qpmpc/mpc_problem.py:88-102``).
    Tensors are moved to ``device`` / ``dtype`` and made contiguous once, here.

    Args:
        ltv: Names among ``"A", "B", "C", "D", "e"`` whose 3-D (2-D for e)
            tensor is a per-step stack shared by the batch rather than a
            per-instance stack.
        batch_size: Needed only when no operand carries the batch axis.
        paired_rows: Whether every ``C_k``, ``D_k`` has the form ``[M; -M]`` (two-sided
            bounds, e.g. ``|u| <= u_max``): the kernels then keep one row per pair
            (``qpmpc_b200_desc.paired``).  ``None`` (default) checks the operands once, here.
    """

    def __init__(
        self,
        transition_state_matrix,
        transition_input_matrix,
        ineq_state_matrix,
        ineq_input_matrix,
        ineq_vector,
        nb_timesteps: int,
        terminal_cost_weight: Optional[float],
        stage_state_cost_weight: Optional[float],
        stage_input_cost_weight: float,
        initial_state=None,
        goal_state=None,
        target_states=None,
        ltv: Iterable[str] = (),
        batch_size: Optional[int] = None,
        dtype: torch.dtype = torch.float64,
        device=None,
        paired_rows: Optional[bool] = None,
    ) -> None:
        if stage_input_cost_weight <= 0.0:
            raise ProblemDefinitionError("the input weight must be positive")
        if terminal_cost_weight is None and stage_state_cost_weight is None:
            raise ProblemDefinitionError(
                "set a terminal cost weight, a stage state cost weight, or both"
            )
        if dtype not in _DTYPE_CODE:
            raise ProblemDefinitionError("dtype must be torch.float64 or torch.float32")
        self.device = _require_cuda(device)
        self.dtype = dtype
        self.nb_timesteps = int(nb_timesteps)
        self.terminal_cost_weight = terminal_cost_weight
        self.stage_state_cost_weight = stage_state_cost_weight
        self.stage_input_cost_weight = float(stage_input_cost_weight)
        self._ltv = set(ltv)
        N = self.nb_timesteps

        A = self._tensor(transition_state_matrix)
        B = self._tensor(transition_input_matrix)
        self.state_dim = int(A.shape[-1])
        self.input_dim = int(B.shape[-1])
        nx, nu = self.state_dim, self.input_dim
        e = self._tensor(ineq_vector)
        self.ineq_dim = int(e.shape[-1]) if e is not None else 0
        nc = self.ineq_dim
        C = self._tensor(ineq_state_matrix)
        D = self._tensor(ineq_input_matrix)

        self._batch = batch_size
        self.A, self.mode_A = self._matrix("A", A, (nx, nx), False)
        self.B, self.mode_B = self._matrix("B", B, (nx, nu), False)
        self.C, self.mode_C = self._matrix("C", C, (nc, nx), True)
        self.D, self.mode_D = self._matrix("D", D, (nc, nu), True)
        self.e, self.mode_e = self._matrix("e", e, (nc,), nc == 0)
        self.paired_rows = self._rows_paired() if paired_rows is None else bool(paired_rows)
        self.x0 = self.goal = self.targets = None
        self.mode_x0 = self.mode_goal = self.mode_targets = _capi.VEC_ABSENT
        if initial_state is not None:
            self.update_initial_state(initial_state)
        if goal_state is not None:
            self.update_goal_state(goal_state)
        if target_states is not None:
            self.update_target_states(target_states)
        if self._batch is None:
            raise ProblemDefinitionError(
                "no operand carries a batch axis: pass batch_size explicitly"
            )

    # -- packing -----------------------------------------------------------

    def _tensor(self, value) -> Optional[torch.Tensor]:
        if value is None:
            return None
        if not isinstance(value, torch.Tensor):
            value = torch.as_tensor(np.asarray(value))
        return value.to(device=self.device, dtype=self.dtype).contiguous()

    def _rows_paired(self) -> bool:
        nc = self.ineq_dim
        if nc == 0 or nc % 2 or (self.C is None and self.D is None):
            return False
        h = nc // 2
        return all(bool(torch.equal(t[..., :h, :], -t[..., h:, :])) for t in (self.C, self.D) if t is not None)

    def _set_batch(self, b: int, what: str) -> None:
        if self._batch is None:
            self._batch = int(b)
        elif self._batch != int(b):
            raise ProblemDefinitionError(
                f"{what} has batch {b}, other operands have {self._batch}"
            )

    def _matrix(self, name, t, item, optional):
        if t is None or 0 in item:
            if not optional and 0 not in item:
                raise ProblemDefinitionError(f"operand {name} is required")
            return None, _capi.ABSENT
        k = len(item)
        N = self.nb_timesteps
        if tuple(t.shape[-k:]) != tuple(item):
            raise ProblemDefinitionError(
                f"operand {name} has trailing shape {tuple(t.shape[-k:])}, expected {item}"
            )
        lead = t.shape[:-k]
        if len(lead) == 0:
            return t, _capi.SHARED_LTI
        if len(lead) == 1 and name in self._ltv:
            if lead[0] != N:
                raise ProblemDefinitionError(f"operand {name}: per-step stack needs {N} steps")
            return t, _capi.SHARED_LTV
        if len(lead) == 1:
            self._set_batch(lead[0], f"operand {name}")
            return t, _capi.BATCH_LTI
        if len(lead) == 2:
            if lead[1] != N:
                raise ProblemDefinitionError(f"operand {name}: per-step stack needs {N} steps")
            self._set_batch(lead[0], f"operand {name}")
            return t, _capi.BATCH_LTV
        raise ProblemDefinitionError(f"operand {name} has too many axes: {tuple(t.shape)}")

    def _vector(self, value, size, what):
        t = self._tensor(value)
        if t.ndim <= 1:
            if t.numel() != size:
                raise StateError(f"{what} has {t.numel()} entries, expected {size}")
            return t.reshape(size), _capi.VEC_SHARED
        t = t.reshape(t.shape[0], -1)
        if t.shape[1] != size:
            raise StateError(f"{what} has {t.shape[1]} entries per instance, expected {size}")
        self._set_batch(t.shape[0], what)
        return t.contiguous(), _capi.VEC_BATCH

    # -- state setters (mpc_problem.py:247-295) ------------------------------

    def update_initial_state(self, initial_state) -> None:
        """Set x_0: [B, nx] per instance or [nx] shared."""
        self.x0, self.mode_x0 = self._vector(initial_state, self.state_dim, "initial state")

    def update_goal_state(self, goal_state) -> None:
        """Set the terminal goal: [B, nx] or [nx]."""
        self.goal, self.mode_goal = self._vector(goal_state, self.state_dim, "goal state")

    def update_target_states(self, target_states) -> None:
        """Set the stage targets for x_0..x_{N-1}: [B, N*nx] (or [B, N, nx]) or [N*nx]."""
        t = self._tensor(target_states)
        size = self.state_dim * self.nb_timesteps
        if t.ndim == 3:
            t = t.reshape(t.shape[0], -1)
        elif t.ndim == 2 and t.numel() == size:
            t = t.reshape(size)
        self.targets, self.mode_targets = self._vector(t, size, "target states")

    @property
    def batch_size(self) -> int:
        return int(self._batch)

    @property
    def nb_vars(self) -> int:
        return self.nb_timesteps * self.input_dim

    @property
    def nb_rows(self) -> int:
        return self.nb_timesteps * self.ineq_dim

    # -- C ABI views ----------------------------------------------------------

    def __setattr__(self, name, value):
        # the C-ABI views below are cached (a small batch is bound by this call path, not by the
        # kernel); rebinding any attribute -- an operand, a weight, a state -- drops them.  Writing
        # INTO a tensor changes neither its address nor the descriptor.
        object.__setattr__(self, name, value)
        if name[:3] != "_c_":
            object.__setattr__(self, "_c_ops", None)
            object.__setattr__(self, "_c_desc", {})

    def desc(self, method: int = _capi.ACTIVE_SET, max_iter: int = 0, tol: float = 0.0,
             polish: bool = True) -> _capi.Desc:
        """The descriptor of the C ABI (cached: treat it as read-only)."""
        key = (method, max_iter, tol, polish)
        d = self._c_desc.get(key)
        if d is not None:
            return d
        if self.x0 is None:
            raise ProblemDefinitionError("initial state is undefined")  # mpc_qp.py:49-51
        d = self._c_desc[key] = _capi.Desc()
        d.batch, d.N = self.batch_size, self.nb_timesteps
        d.nx, d.nu, d.nc = self.state_dim, self.input_dim, self.ineq_dim
        d.dtype = _DTYPE_CODE[self.dtype]
        d.mode_A, d.mode_B, d.mode_C = self.mode_A, self.mode_B, self.mode_C
        d.mode_D, d.mode_e = self.mode_D, self.mode_e
        d.mode_x0, d.mode_goal, d.mode_targets = self.mode_x0, self.mode_goal, self.mode_targets
        d.has_wt = self.terminal_cost_weight is not None
        d.has_wx = self.stage_state_cost_weight is not None
        d.w_t = float(self.terminal_cost_weight or 0.0)
        d.w_x = float(self.stage_state_cost_weight or 0.0)
        d.w_u = self.stage_input_cost_weight
        d.method, d.max_iter, d.tol = method, int(max_iter), float(tol)
        d.flags = 0 if polish else _capi.FLAG_NO_POLISH
        d.paired = int(self.paired_rows)
        return d

    def operands(self) -> _capi.Operands:
        """The operand pointers of the C ABI (cached: treat them as read-only)."""
        ops = self._c_ops
        if ops is None:
            ops = _capi.Operands(
                _ptr(self.A), _ptr(self.B), _ptr(self.C), _ptr(self.D), _ptr(self.e),
                _ptr(self.x0), _ptr(self.goal), _ptr(self.targets),
            )
            object.__setattr__(self, "_c_ops", ops)
        return ops

    # -- construction from host-side problems ---------------------------------

    @classmethod
    def from_problems(cls, problems: Sequence[MPCProblem], dtype=torch.float64, device=None):
        """Stack host-side ``MPCProblem`` objects of one shape (convenience;
        for throughput build the tensors directly)."""
        first = problems[0]
        N = first.nb_timesteps
        packed = [pack_problem(p) for p in problems]
        for p_, pk in zip(problems, packed):
            same = (p_.nb_timesteps == N and p_.terminal_cost_weight == first.terminal_cost_weight
                    and p_.stage_state_cost_weight == first.stage_state_cost_weight
                    and p_.stage_input_cost_weight == first.stage_input_cost_weight
                    and pk["row_map"] == packed[0]["row_map"] and pk["nc"] == packed[0]["nc"])
            if not same:
                raise ProblemDefinitionError(
                    "from_problems needs problems with the same horizon, weights and per-step row counts")

        def stack(key):
            vals = [pk[key] for pk in packed]
            if vals[0] is None:
                return None
            return np.stack(vals)

        obj = cls(
            stack("A"), stack("B"), stack("C"), stack("D"), stack("e"), N,
            first.terminal_cost_weight, first.stage_state_cost_weight,
            first.stage_input_cost_weight,
            initial_state=stack("x0"), goal_state=stack("goal"),
            target_states=stack("targets"), dtype=dtype, device=device,
        )
        obj.row_map = packed[0]["row_map"]
        return obj


def pack_problem(problem: MPCProblem) -> dict:
    """Canonical arrays of one ``MPCProblem``: A, B [N?, ...] etc.

    LTV operands (Python lists, ``mpc_problem.py:178-180``) become [N, r, c]
    stacks.  Ragged per-step row counts are padded to the largest ``nc`` with
    all-zero rows with bound 1, and ``row_map`` lists the real rows.
    """
    N = problem.nb_timesteps
    if problem.initial_state is None:
        raise ProblemDefinitionError("initial state is undefined")  # mpc_qp.py:49-51
    if not isinstance(problem.ineq_vector, list):
        # time-invariant rows (the common case, and the one a control loop calls every cycle):
        # nothing is ragged, nothing needs stacking
        e0 = np.asarray(problem.ineq_vector, dtype=float).reshape(-1)
        nc = e0.shape[0]
        nx, nu = problem.state_dim, problem.input_dim
        A, B = problem.transition_state_matrix, problem.transition_input_matrix
        C, D = problem.ineq_state_matrix, problem.ineq_input_matrix
        if not any(isinstance(op, list) for op in (A, B, C, D)):
            return dict(A=np.asarray(A, dtype=float), B=np.asarray(B, dtype=float).reshape(nx, nu),
                        C=None if C is None else np.asarray(C, dtype=float).reshape(-1, nx),
                        D=None if D is None else np.asarray(D, dtype=float).reshape(-1, nu),
                        e=e0, x0=problem.initial_state, goal=problem.goal_state, targets=problem.target_states,
                        nc=nc, row_map=list(range(N * nc)))
    e_steps = [np.asarray(problem.get_ineq_vector(k), dtype=float).reshape(-1) for k in range(N)]
    ncs = [ek.shape[0] for ek in e_steps]
    nc = max(ncs)
    ragged = len(set(ncs)) > 1
    # bound of the padding rows: 0 . u <= 1 is never active and is well scaled for both methods
    # (a huge "no bound" constant would dominate the interior point's scale estimates)
    PAD_BOUND = 1.0

    def mat(op, cols):
        if op is None:
            return None
        ltv = isinstance(op, list)
        if not ltv and not ragged:
            return np.asarray(op, dtype=float).reshape(-1, cols) if cols else np.asarray(op, dtype=float)
        out = np.zeros((N, nc, cols))
        for k in range(N):
            blk = op[k] if ltv else op
            if blk is not None:
                blk = np.asarray(blk, dtype=float).reshape(-1, cols)
                out[k, : blk.shape[0]] = blk
        return out

    def dyn(op):
        return np.stack([np.asarray(a, dtype=float) for a in op]) if isinstance(op, list) else np.asarray(op, dtype=float)

    nx, nu = problem.state_dim, problem.input_dim
    A = dyn(problem.transition_state_matrix)
    B = dyn(problem.transition_input_matrix)
    B = B.reshape((N, nx, nu)) if isinstance(problem.transition_input_matrix, list) else B.reshape((nx, nu))
    C, D = problem.ineq_state_matrix, problem.ineq_input_matrix
    if isinstance(C, list) and any(c is None for c in C):
        C = [np.zeros((ncs[k], nx)) if c is None else c for k, c in enumerate(C)]
    if isinstance(D, list) and any(d is None for d in D):
        D = [np.zeros((ncs[k], nu)) if d is None else d for k, d in enumerate(D)]
    Cm, Dm = mat(C, nx), mat(D, nu)
    if isinstance(problem.ineq_vector, list) or ragged:
        e = np.full((N, nc), PAD_BOUND)
        for k in range(N):
            e[k, : ncs[k]] = e_steps[k]
    else:
        e = e_steps[0]
    row_map = [k * nc + r for k in range(N) for r in range(ncs[k])]
    return dict(A=A, B=B, C=Cm, D=Dm, e=e, x0=problem.initial_state,
                goal=problem.goal_state, targets=problem.target_states,
                nc=nc, row_map=row_map)


def problem_to_batch(problem: MPCProblem, dtype=torch.float64, device=None) -> BatchedMPCProblem:
    """One host-side ``MPCProblem`` as a batch of one (what ``solve_mpc`` does)."""
    pk = pack_problem(problem)
    names = [k for k in ("A", "B", "C", "D", "e")
             if pk[k] is not None and pk[k].ndim == (2 if k == "e" else 3)]
    obj = BatchedMPCProblem(
        pk["A"], pk["B"], pk["C"], pk["D"], pk["e"], problem.nb_timesteps,
        problem.terminal_cost_weight, problem.stage_state_cost_weight,
        problem.stage_input_cost_weight, initial_state=pk["x0"],
        goal_state=pk["goal"], target_states=pk["targets"], ltv=names,
        batch_size=1, dtype=dtype, device=device,
    )
    obj.row_map = pk["row_map"]
    return obj


class BatchedPlan:
    """Result of :func:`solve_mpc_batch` -- batched ``Plan`` (``qpmpc/plan.py``).

    Attributes:
        problem: The batched problem that was solved.
        inputs: U, tensor [B, N, nu]; rows of unsolved instances are NaN.
        status: int32 [B]: 0 solved, 1 iteration limit, 2 infeasible, 3 not SPD.
        iters: int32 [B] solver iterations.
        multipliers: [B, N*nc] or None.
    """

    def __init__(self, problem, U, status, iters, Z=None):
        self.problem = problem
        self.inputs = U.view(problem.batch_size, problem.nb_timesteps, problem.input_dim)
        self.status = status
        self.iters = iters
        self.multipliers = Z
        self._states = None

    @property
    def found(self) -> torch.Tensor:
        """bool [B]: ``qpsol.found`` per instance."""
        return self.status == 0

    @property
    def is_empty(self) -> torch.Tensor:
        """bool [B]: batched ``Plan.is_empty`` (``plan.py:45-48``)."""
        return self.status != 0

    @property
    def first_input(self) -> torch.Tensor:
        """u_0 of every instance, [B, nu] (``plan.py:50-62``)."""
        return self.inputs[:, 0, :]

    @property
    def states(self) -> torch.Tensor:
        """X [B, N+1, nx], integrated on the device on first access
        (``plan.py:81-109`` -> ``mpc_problem.py:316-335``)."""
        if self._states is None:
            self._states = integrate_batch(self.problem, self.inputs)
        return self._states


class FactoredModel:
    """The record of :func:`factor_model`: everything of the condensed QP that does not depend
    on the initial / goal / target states, computed once for a model shared by the batch --
    the batched counterpart of keeping one ``MPCQP`` and only calling ``update_cost_vector`` /
    ``update_constraint_vector`` between solves (``qpmpc/mpc_qp.py:129-163``).  Valid as long as
    A, B, C, D, the weights and the presence of goal / targets stay as they were."""

    def __init__(self, record: torch.Tensor, signature):
        self.record = record
        self.signature = signature


def _model_signature(problem: "BatchedMPCProblem"):
    return (problem.nb_timesteps, problem.state_dim, problem.input_dim, problem.ineq_dim, problem.dtype,
            problem.terminal_cost_weight, problem.stage_state_cost_weight, problem.stage_input_cost_weight,
            problem.mode_A, problem.mode_B, problem.mode_C, problem.mode_D,
            problem.mode_goal != _capi.VEC_ABSENT, problem.mode_targets != _capi.VEC_ABSENT,
            *(None if t is None else t.data_ptr() for t in (problem.A, problem.B, problem.C, problem.D)))


def _resolve_factored(problem: "BatchedMPCProblem", factored):
    """``factored`` of the closed loops: a :class:`FactoredModel`, ``False`` (never: condense and
    factor every instance every cycle) or ``None`` (factor the model here when it is shared by
    the batch -- inside one loop call it cannot change -- else fall back to the full path)."""
    if factored is False:
        return None
    if factored is not None:
        if factored.signature != _model_signature(problem):
            raise ProblemDefinitionError("the factored model belongs to another problem")
        return factored
    desc = problem.desc()
    if int(_capi.load().qpmpc_b200_factor_bytes(ctypes.byref(desc))) == 0:
        return None
    return factor_model(problem)


def factor_model(problem: BatchedMPCProblem) -> FactoredModel:
    """Condense and factor the model of ``problem`` once (``qpmpc_b200_factor``).  Needs A, B, C,
    D shared by the batch and paired constraint rows; raises ``BackendError`` otherwise."""
    lib = _capi.load()
    desc = problem.desc()
    nbytes = int(lib.qpmpc_b200_factor_bytes(ctypes.byref(desc)))
    if nbytes == 0:
        raise BackendError("the shared-model fast path needs A, B, C, D shared by the batch, paired "
                           "constraint rows and N*nu <= 32")
    with _device_guard(problem.device):
        record = torch.empty(nbytes // problem.x0.element_size(), dtype=problem.dtype, device=problem.device)
        ops = problem.operands()
        rc = lib.qpmpc_b200_factor(ctypes.byref(desc), ctypes.byref(ops), _ptr(record), _stream_ptr(problem.device))
    _capi.check(rc, "qpmpc_b200_factor")
    return FactoredModel(record, _model_signature(problem))


def solve_mpc_batch(
    problem: BatchedMPCProblem,
    method: str = "active_set",
    max_iter: int = 0,
    tol: float = 0.0,
    return_multipliers: bool = False,
    out: Optional[torch.Tensor] = None,
    polish: bool = True,
    factored: Optional[FactoredModel] = None,
) -> BatchedPlan:
    """Condense and solve every instance of ``problem`` on its CUDA device.

    ``factored`` (a :class:`FactoredModel` of this problem's shared model) skips condensing
    and factoring: only q and h are rebuilt per instance.
    ``method="active_set"`` (default) is the exact Goldfarb-Idnani kernel;
    ``method="pdip"`` the Mehrotra interior-point kernel stopped at ``tol``
    (default 1e-9), followed by an active-set ``polish`` -- the role a
    ``solver="proxqp"`` / ``"osqp"`` string plays at ``qpmpc/solve_mpc.py:43``.
    Asynchronous on the current CUDA stream, like any torch op.
    """
    lib = _capi.load()
    meth = {"active_set": _capi.ACTIVE_SET, "pdip": _capi.PDIP}.get(method)
    if meth is None:
        raise ProblemDefinitionError(f"unknown method {method!r}")
    desc = problem.desc(meth, max_iter, tol, polish)
    B, n, m = problem.batch_size, problem.nb_vars, problem.nb_rows
    if out is not None and (tuple(out.shape) != (B, n) or out.dtype != problem.dtype
                            or out.device != problem.device or not out.is_contiguous()):
        raise ProblemDefinitionError(
            f"out must be a contiguous {problem.dtype} tensor of shape {(B, n)} on {problem.device}")
    with _device_guard(problem.device):
        U = out if out is not None else torch.empty((B, n), dtype=problem.dtype, device=problem.device)
        status = torch.empty(B, dtype=torch.int32, device=problem.device)
        iters = torch.empty(B, dtype=torch.int32, device=problem.device)
        Z = torch.empty((B, m), dtype=problem.dtype, device=problem.device) if return_multipliers else None
        outs = _capi.Outputs(_ptr(U), _ptr(status), _ptr(iters), _ptr(Z))
        ops = problem.operands()
        stream = _stream_ptr(problem.device)
        if factored is not None:
            if factored.signature != _model_signature(problem):
                raise ProblemDefinitionError("the factored model belongs to another problem (model, weights or "
                                             "presence of goal / targets changed): call factor_model again")
            rc = lib.qpmpc_b200_solve_factored(ctypes.byref(desc), ctypes.byref(ops), _ptr(factored.record),
                                               ctypes.byref(outs), stream)
        else:
            rc = lib.qpmpc_b200_solve(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs), stream)
    _capi.check(rc, "qpmpc_b200_solve_factored" if factored is not None else "qpmpc_b200_solve")
    return BatchedPlan(problem, U, status, iters, Z)


def condense_batch(problem: BatchedMPCProblem, fields: Sequence[str] = ("P", "q", "G", "h")) -> dict:
    """Materialise condensed-QP fields of every instance (MPCQP, batched)."""
    lib = _capi.load()
    desc = problem.desc()
    B, N = problem.batch_size, problem.nb_timesteps
    n, m, nx = problem.nb_vars, problem.nb_rows, problem.state_dim
    shapes = dict(P=(B, n, n), q=(B, n), G=(B, m, n), h=(B, m), Phi=(B, N * nx, nx),
                  Psi=(B, N * nx, n), phi_last=(B, nx, nx), psi_last=(B, nx, n))
    with _device_guard(problem.device):
        out = {k: torch.zeros(shapes[k], dtype=problem.dtype, device=problem.device) for k in fields}
        qf = _capi.QPFields(*[_ptr(out.get(k)) for k in
                              ("P", "q", "G", "h", "Phi", "Psi", "phi_last", "psi_last")])
        ops = problem.operands()
        stream = _stream_ptr(problem.device)
        rc = lib.qpmpc_b200_condense(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(qf), stream)
    _capi.check(rc, "qpmpc_b200_condense")
    return out


def integrate_batch(problem: BatchedMPCProblem, inputs: torch.Tensor) -> torch.Tensor:
    """X [B, N+1, nx] from x0 and U on the device (batched ``integrate``)."""
    lib = _capi.load()
    desc = problem.desc()
    B, N, nx = problem.batch_size, problem.nb_timesteps, problem.state_dim
    U = inputs.reshape(B, -1).to(problem.dtype).contiguous()
    with _device_guard(problem.device):
        X = torch.empty((B, N + 1, nx), dtype=problem.dtype, device=problem.device)
        ops = problem.operands()
        stream = _stream_ptr(problem.device)
        rc = lib.qpmpc_b200_integrate(ctypes.byref(desc), ctypes.byref(ops), _ptr(U), _ptr(X), stream)
    _capi.check(rc, "qpmpc_b200_integrate")
    return X


def pendulum_closed_loop(
    problem: BatchedMPCProblem,
    v_target,
    cycles: int,
    substeps: int = 15,
    length: float = 0.6,
    gravity: float = 9.81,
    sampling_period: float = 0.1,
    record: bool = False,
    factored=None,
    stats: bool = False,
):
    """Receding-horizon closed loop of the wheeled inverted pendulum on the
    device (``examples/wheeled_inverted_pendulum.py:99-118``, batched): per
    cycle the reference trajectory is rebuilt from the state, the MPC is
    condensed and solved, and the nonlinear plant advances ``substeps`` steps
    under the first input.  ``problem.x0`` is the state and is updated in
    place; goal / targets are (re)allocated per instance.

    ``factored``: the model's :func:`factor_model` record (goal and targets must be present in
    the problem it was made from), ``None`` (default: the loop factors the model itself when it is
    shared by the batch) or ``False`` (condense and factor every instance every cycle).  With a
    factored model every cycle only rebuilds q and h, and the whole loop
    runs inside ONE kernel launch (each lane group solves, moves its plant and rewrites its
    targets in shared memory; ``QPMPC_B200_LOOP_FUSED=0``: a solve and a plant launch per cycle).

    Returns ``(plan_of_last_cycle, trajectory or None, unsolved_count_tensor)``;
    trajectory is [cycles + 1, B, 4].  With ``stats`` a fourth value: dict(upright = device
    counter of (instance, cycle) pairs solved with |pitch| <= 1.2 rad, iterations = [cycles]
    solver iterations summed over the batch).  Asynchronous on the current stream.
    """
    lib = _capi.load()
    B, N, nx = problem.batch_size, problem.nb_timesteps, problem.state_dim
    if nx != 4 or problem.input_dim != 1:
        raise ProblemDefinitionError("the pendulum loop needs state_dim 4 and input_dim 1")
    dev, dt_ = problem.device, problem.dtype
    if problem.x0 is None or problem.mode_x0 != _capi.VEC_BATCH:
        raise ProblemDefinitionError("per-instance initial states [B, 4] are required")
    with _device_guard(dev):
        problem.goal = torch.empty((B, nx), dtype=dt_, device=dev)
        problem.mode_goal = _capi.VEC_BATCH
        problem.targets = torch.empty((B, N * nx), dtype=dt_, device=dev)
        problem.mode_targets = _capi.VEC_BATCH
        v = torch.as_tensor(v_target).to(device=dev, dtype=dt_).reshape(-1)
        if v.numel() == 1:
            v = v.expand(B)
        v = v.contiguous()
        if v.numel() != B:
            raise StateError(f"v_target has {v.numel()} entries, expected {B}")
        n = problem.nb_vars
        U = torch.empty((B, n), dtype=dt_, device=dev)
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        iters = torch.zeros(B, dtype=torch.int32, device=dev)
        traj = torch.empty((cycles + 1, B, nx), dtype=dt_, device=dev) if record else None
        unsolved = torch.zeros(1, dtype=torch.int32, device=dev)
        upright = torch.zeros(1, dtype=torch.int32, device=dev) if stats else None
        iter_hist = torch.zeros(max(cycles, 1), dtype=torch.int64, device=dev) if stats else None
        factored = _resolve_factored(problem, factored)
        desc = problem.desc()
        ops = problem.operands()
        outs = _capi.Outputs(_ptr(U), _ptr(status), _ptr(iters), None)
        loop = _capi.ClosedLoop(int(cycles), int(substeps), sampling_period / substeps,
                                float(sampling_period), float(length), float(gravity),
                                _ptr(v), _ptr(traj), _ptr(unsolved),
                                _ptr(factored.record) if factored is not None else None,
                                _ptr(upright), _ptr(iter_hist))
        stream = _stream_ptr(dev)
        rc = lib.qpmpc_b200_pendulum_closed_loop(ctypes.byref(desc), ctypes.byref(ops),
                                                 ctypes.byref(outs), ctypes.byref(loop), stream)
    _capi.check(rc, "qpmpc_b200_pendulum_closed_loop")
    if stats:
        return BatchedPlan(problem, U, status, iters), traj, unsolved, dict(upright=upright, iterations=iter_hist)
    return BatchedPlan(problem, U, status, iters), traj, unsolved


def lipm_walking_closed_loop(
    problem: BatchedMPCProblem,
    support_foot,
    strides,
    phase_index,
    stride_index,
    cycles: int,
    substeps: int = 15,
    dsp_duration: float = 0.1,
    ssp_duration: float = 0.7,
    sampling_period: float = 0.1,
    foot_size: float = 0.065,
    max_zmp_dist: float = 100.0,
    record: bool = False,
    factored=None,
):
    """Walking loop of ``examples/lipm_walking_controller.py:307-335`` on the device, batched:
    per cycle the phase machine writes the ZMP bounds ``e_k`` of the horizon and the goal, the
    MPC is condensed and solved, the state advances ``substeps`` constant-jerk steps under the
    first input, and the phase (and, at a foot switch, the support foot) advances.

    ``problem.x0`` [B, 3] is the state and is updated in place; ``problem.e`` [B, N, 2] and
    ``problem.goal`` [B, 3] are (re)allocated and rewritten every cycle.  ``support_foot`` [B],
    ``phase_index`` [B] and ``stride_index`` [B] are the phase machine's state (updated in place
    when given as device tensors), ``strides`` [B, 2] the alternating strides.

    ``factored`` as in :func:`pendulum_closed_loop` (default: the loop factors the shared model
    itself and runs all its cycles inside one kernel launch; ``False``: the full path per cycle).

    Returns ``(plan_of_last_cycle, trajectory or None, unsolved_count_tensor, phase_state)``
    with ``phase_state = dict(support_foot, phase_index, stride_index)``.  Asynchronous.
    """
    lib = _capi.load()
    B, N = problem.batch_size, problem.nb_timesteps
    if problem.state_dim != 3 or problem.input_dim != 1 or problem.ineq_dim != 2:
        raise ProblemDefinitionError("the walking loop needs state_dim 3, input_dim 1 and two rows per step")
    dev, dt_ = problem.device, problem.dtype
    if problem.x0 is None or problem.mode_x0 != _capi.VEC_BATCH:
        raise ProblemDefinitionError("per-instance initial states [B, 3] are required")
    with _device_guard(dev):
        if problem.e is None or problem.mode_e != _capi.BATCH_LTV:
            problem.e = torch.empty((B, N, 2), dtype=dt_, device=dev)
            problem.mode_e = _capi.BATCH_LTV
        if problem.goal is None or problem.mode_goal != _capi.VEC_BATCH:
            problem.goal = torch.zeros((B, 3), dtype=dt_, device=dev)
            problem.mode_goal = _capi.VEC_BATCH

        def vec(v, dtype, shape):
            t = torch.as_tensor(v).to(device=dev, dtype=dtype)
            if t.numel() * B == int(np.prod(shape)):
                t = t.expand(shape)
            t = t.reshape(shape).contiguous()
            return t

        foot = vec(support_foot, dt_, (B,))
        strd = vec(strides, dt_, (B, 2))
        pidx = vec(phase_index, torch.int32, (B,))
        sidx = vec(stride_index, torch.int32, (B,))
        n = problem.nb_vars
        U = torch.empty((B, n), dtype=dt_, device=dev)
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        iters = torch.zeros(B, dtype=torch.int32, device=dev)
        traj = torch.empty((cycles + 1, B, 3), dtype=dt_, device=dev) if record else None
        unsolved = torch.zeros(1, dtype=torch.int32, device=dev)
        factored = _resolve_factored(problem, factored)
        desc = problem.desc()
        ops = problem.operands()
        outs = _capi.Outputs(_ptr(U), _ptr(status), _ptr(iters), None)
        loop = _capi.LipmLoop(int(cycles), int(substeps), int(round(dsp_duration / sampling_period)),
                              int(round(ssp_duration / sampling_period)), float(sampling_period), float(foot_size),
                              float(max_zmp_dist), _ptr(foot), _ptr(strd), _ptr(pidx), _ptr(sidx), _ptr(traj),
                              _ptr(unsolved), _ptr(factored.record) if factored is not None else None)
        rc = lib.qpmpc_b200_lipm_closed_loop(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs),
                                             ctypes.byref(loop), _stream_ptr(dev))
    _capi.check(rc, "qpmpc_b200_lipm_closed_loop")
    return (BatchedPlan(problem, U, status, iters), traj, unsolved,
            dict(support_foot=foot, phase_index=pidx, stride_index=sidx))
