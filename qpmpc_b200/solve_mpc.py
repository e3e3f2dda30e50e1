"""One-stop entry point: ``solve_mpc(problem, solver) -> Plan``.

Mirror of ``qpmpc/solve_mpc.py:16-44`` (reference tree).  ``solver="b200"``
(aliases ``"cuda"``, and -- because every solver name the reference's callers
pass means "solve this QP" -- any other name too, see ``strict``) routes the
problem through the fused CUDA kernel as a batch of one.  There is no CPU
fallback: without the CUDA library or a device a ``BackendError`` is raised.
"""

import torch

from .batched import problem_to_batch, solve_mpc_batch
from .exceptions import BackendError
from .mpc_problem import MPCProblem
from .mpc_qp import MPCQP  # noqa: F401  (reference tests import it from here)
from .plan import Plan
from .solution import Solution

NATIVE_SOLVERS = ("b200", "cuda")


def solve_mpc(
    problem: MPCProblem,
    solver: str = "b200",
    sparse: bool = False,
    strict: bool = False,
    **kwargs,
) -> Plan:
    """Solve a linear time-variant MPC problem.

    Args:
        problem: The problem to solve.
        solver: ``"b200"`` / ``"cuda"``.  Names of qpsolvers backends
            (``"proxqp"``, ``"quadprog"``, ...) are accepted and served by the
            CUDA engine as well, so reference call sites run unchanged; with
            ``strict=True`` they raise ``BackendError`` instead.
        sparse: Accepted for signature parity; the engine is dense.
        kwargs: ``max_iter``, ``method`` (``"active_set"`` or ``"pdip"``),
            ``tol``, ``dtype``.  Other solver keywords (``eps_abs`` ...) are
            ignored: the active-set method is exact.

    Returns:
        The plan; ``plan.is_empty`` when the QP has no solution.
    """
    del sparse
    if strict and solver not in NATIVE_SOLVERS:
        raise BackendError(
            f"solver {solver!r} is a qpsolvers backend; this engine provides "
            f"{NATIVE_SOLVERS}"
        )
    batch = problem_to_batch(problem, dtype=kwargs.get("dtype", torch.float64))
    plan = solve_mpc_batch(
        batch,
        method=kwargs.get("method", "active_set"),
        max_iter=int(kwargs.get("max_iter", 0)),
        tol=float(kwargs.get("tol", 0.0)),
        return_multipliers=True,
    )
    status = int(plan.status[0].item())
    found = status == 0
    rows = batch.row_map
    qpsol = Solution(
        found=found,
        x=plan.inputs[0].reshape(-1).double().cpu().numpy() if found else None,
        z=plan.multipliers[0].double().cpu().numpy()[rows] if found and rows else None,
        extras={"status": status, "iters": int(plan.iters[0].item())},
    )
    return Plan(problem, qpsol)
