"""B200-native batched linear MPC with the qpmpc surface.

``MPCProblem`` / ``MPCQP`` / ``Plan`` / ``solve_mpc`` mirror
``qpmpc/__init__.py:9-19`` of the reference; ``BatchedMPCProblem`` /
``solve_mpc_batch`` / ``BatchedPlan`` are the tensor-native batched surface
the CUDA kernels were built for.  Importing this package needs no GPU; solving
does (there is no CPU fallback).
"""

from .exceptions import (
    BackendError,
    PlanError,
    ProblemDefinitionError,
    QPMPCException,
    StateError,
)
from .mpc_problem import MPCProblem
from .plan import Plan
from .solution import QPProblem, Solution

__version__ = "0.1.0"

from .batched import (  # noqa: E402
    BatchedMPCProblem,
    BatchedPlan,
    FactoredModel,
    condense_batch,
    factor_model,
    integrate_batch,
    lipm_walking_closed_loop,
    pendulum_closed_loop,
    problem_to_batch,
    solve_mpc_batch,
)
from .mpc_qp import MPCQP  # noqa: E402
from .solve_mpc import solve_mpc  # noqa: E402  (rebinds the name from module to function)

__all__ = [
    "BackendError", "BatchedMPCProblem", "BatchedPlan", "FactoredModel", "MPCProblem", "MPCQP",
    "Plan", "PlanError", "ProblemDefinitionError", "QPMPCException", "QPProblem",
    "Solution", "StateError", "condense_batch", "factor_model", "integrate_batch", "lipm_walking_closed_loop", "pendulum_closed_loop",
    "solve_mpc", "solve_mpc_batch",
]
