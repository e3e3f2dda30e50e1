"""One-stop entry point: ``solve_mpc(problem, solver) -> Plan``.

Mirror of ``qpmpc/solve_mpc.py:16-44`` (reference tree).  The ``solver`` string
keeps the role it has there (``qpsolvers.solve_problem(..., solver=solver)``,
``solve_mpc.py:43``):

* ``"b200"`` / ``"cuda"`` -- the fused CUDA kernel, as a batch of one;
  ``"b200_pdip"`` selects its interior-point method (same as
  ``method="pdip"``).
* a qpsolvers backend name (``"proxqp"``, ``"quadprog"``, ...) -- handed to
  ``qpsolvers`` when that package is importable (the reference's own path, with
  this package's condensing); otherwise refused with ``BackendError`` unless
  the caller opted in to having those names served by the CUDA engine
  (:func:`serve_qpsolvers_names`, or ``QPMPC_B200_SERVE_QPSOLVERS_NAMES=1``;
  importing the ``qpmpc`` compatibility alias opts in, which is what lets the
  reference's tests and examples run unmodified).
* anything else raises ``BackendError``.

There is no CPU fallback: without the CUDA library or a device a
``BackendError`` is raised.
"""

import os
import warnings

import torch

from . import _capi
from .batched import pack_problem, problem_to_batch, solve_mpc_batch
from .exceptions import BackendError, ProblemDefinitionError
from .mpc_problem import MPCProblem
from .mpc_qp import MPCQP  # noqa: F401  (reference tests import it from here)
from .plan import Plan
from .single import solve_single
from .solution import Solution

NATIVE_SOLVERS = ("b200", "cuda", "b200_pdip")
# Backend names of qpsolvers 3.x/4.x (what a reference call site may pass).
QPSOLVERS_NAMES = (
    "clarabel", "cvxopt", "daqp", "ecos", "gurobi", "highs", "hpipm", "jaxopt_osqp", "kvxopt",
    "mosek", "nppro", "osqp", "piqp", "proxqp", "qpalm", "qpax", "qpoases", "qpswift", "quadprog",
    "scs", "sip",
)
# Interior-point / augmented-Lagrangian backends map to the interior-point kernel when their
# name is served by the engine; the active-set ones (and everything else) to the exact method.
_IPM_NAMES = ("clarabel", "cvxopt", "ecos", "hpipm", "piqp", "qpswift", "gurobi", "mosek")
_OWN_KWARGS = ("max_iter", "method", "tol", "dtype")

_serve_names = None  # None: follow the environment variable


def serve_qpsolvers_names(enable: bool = True) -> None:
    """Opt in (or out) of serving qpsolvers backend names with the CUDA engine."""
    global _serve_names
    _serve_names = bool(enable)


def _names_served() -> bool:
    if _serve_names is not None:
        return _serve_names
    return os.environ.get("QPMPC_B200_SERVE_QPSOLVERS_NAMES", "0") not in ("", "0")


def _solve_with_qpsolvers(problem: MPCProblem, solver: str, sparse: bool, kwargs) -> Plan:
    """The reference's own lines 42-44 with this package's (CUDA) condensing."""
    import qpsolvers  # noqa: PLC0415

    mpc_qp = MPCQP(problem, sparse=sparse)
    qp = qpsolvers.Problem(mpc_qp.P, mpc_qp.q, mpc_qp.G, mpc_qp.h)
    return Plan(problem, qpsolvers.solve_problem(qp, solver=solver, **kwargs))


def solve_mpc(
    problem: MPCProblem,
    solver: str = "b200",
    sparse: bool = False,
    **kwargs,
) -> Plan:
    """Solve a linear time-variant MPC problem.

    Args:
        problem: The problem to solve.
        solver: See the module docstring.
        sparse: Passed to ``MPCQP`` on the qpsolvers route; the CUDA engine is
            dense and ignores it.
        kwargs: ``max_iter``, ``method`` (``"active_set"`` or ``"pdip"``),
            ``tol``, ``dtype`` for the CUDA engine.  Of the keywords qpsolvers
            callers pass, ``eps_abs`` is taken as ``tol`` of the interior
            point; the others (``eps_rel``, ``initvals``, ``verbose`` ...) do
            not apply to an exact method and are dropped with a warning.

    Returns:
        The plan; ``plan.is_empty`` when the QP has no solution.
    """
    method = kwargs.get("method")
    if solver not in NATIVE_SOLVERS:
        if solver not in QPSOLVERS_NAMES:
            raise BackendError(
                f"unknown solver {solver!r}: this engine provides {NATIVE_SOLVERS} and, through "
                "qpsolvers when it is installed, its backends")
        try:
            import qpsolvers  # noqa: F401, PLC0415
            have_qpsolvers = True
        except ImportError:
            have_qpsolvers = False
        if have_qpsolvers and not _names_served():
            return _solve_with_qpsolvers(problem, solver, sparse, kwargs)
        if not _names_served():
            raise BackendError(
                f"solver {solver!r} is a qpsolvers backend and qpsolvers is not installed; use "
                f"one of {NATIVE_SOLVERS}, or opt in to serving qpsolvers names with the CUDA "
                "engine (qpmpc_b200.serve_qpsolvers_names(), QPMPC_B200_SERVE_QPSOLVERS_NAMES=1, "
                "or `import qpmpc`)")
        if method is None and solver in _IPM_NAMES:
            method = "pdip"
    elif solver == "b200_pdip" and method is None:
        method = "pdip"
    method = method or "active_set"
    tol = float(kwargs.get("tol", 0.0))
    if "eps_abs" in kwargs and "tol" not in kwargs and method == "pdip":
        tol = float(kwargs["eps_abs"])
    dropped = sorted(k for k in kwargs if k not in _OWN_KWARGS and not (k == "eps_abs" and method == "pdip"))
    if dropped:
        warnings.warn(
            f"solve_mpc: keyword(s) {dropped} do not apply to the CUDA engine's {method} method "
            "and were ignored", stacklevel=2)
    dtype = kwargs.get("dtype", torch.float64)
    max_iter = int(kwargs.get("max_iter", 0))
    if os.environ.get("QPMPC_B200_SINGLE_ZEROCOPY", "1") != "0":
        # one launch, no copy call: operands and results in one page-locked block (single.py)
        code = {"active_set": _capi.ACTIVE_SET, "pdip": _capi.PDIP}.get(method)
        if code is None or dtype not in (torch.float64, torch.float32):
            raise ProblemDefinitionError(f"unknown method {method!r}" if code is None else
                                         "dtype must be torch.float64 or torch.float32")
        pk = pack_problem(problem)
        x, z, status, iters = solve_single(problem, pk, code, max_iter, tol, dtype)
        rows = pk["row_map"]
    else:
        # through device tensors, as a batch of one (the path the batched surface takes)
        batch = problem_to_batch(problem, dtype=dtype)
        plan = solve_mpc_batch(batch, method=method, max_iter=max_iter, tol=tol, return_multipliers=True)
        status, iters = int(plan.status[0].item()), int(plan.iters[0].item())
        x = plan.inputs[0].reshape(-1).double().cpu().numpy()
        z = plan.multipliers[0].double().cpu().numpy()
        rows = batch.row_map
    found = status == 0
    qpsol = Solution(
        found=found,
        x=x if found else None,
        z=z[rows] if found and rows else None,
        extras={"status": status, "iters": iters, "method": method},
    )
    return Plan(problem, qpsol)
