"""Error types raised by the host-side surface.

Same hierarchy and names as the reference (``qpmpc/exceptions.py:10-23``) so
that ``except`` clauses written against qpmpc keep working.
"""


class QPMPCException(Exception):
    """Root of every error this package raises on purpose."""


class ProblemDefinitionError(QPMPCException):
    """The MPC problem is ill-posed (bad weights, missing state, ...)."""


class PlanError(QPMPCException):
    """A plan is inconsistent with the problem it was computed for."""


class StateError(QPMPCException):
    """A state vector does not have the expected size."""


class BackendError(QPMPCException):
    """The CUDA engine (or a requested third-party QP backend) is unavailable.

    Not in the reference: there the failure would surface from ``qpsolvers``.
    The engine never falls back to a CPU path, it raises this instead.
    """
