"""Duck-typed stand-in for ``qpsolvers.Solution``.

The reference's ``Plan`` reads only ``.found`` and ``.x`` of the solver's
return value (``qpmpc/plan.py:36-37``).  ``qpsolvers`` is not a dependency of
this engine, so the CUDA path returns this small record instead; it carries the
same two fields plus what the kernel reports per instance.
"""

from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np


@dataclass
class Solution:
    """Result of one condensed-QP solve.

    Attributes:
        found: True if the solver converged (kernel status 0).
        x: Stacked input vector U (N * nu), or None when not found.
        z: Multipliers of ``G u <= h`` (m), when requested.
        obj: Optimal cost ``0.5 u'Pu + q'u``, when available.
        extras: ``{"status", "iters"}`` from the kernel.
    """

    found: bool
    x: Optional[np.ndarray] = None
    z: Optional[np.ndarray] = None
    obj: Optional[float] = None
    extras: Dict[str, Any] = field(default_factory=dict)


@dataclass
class QPProblem:
    """(P, q, G, h) record returned by ``MPCQP.problem``; stands in for
    ``qpsolvers.Problem`` (``qpmpc/mpc_qp.py:124-127``): no equalities, no
    bounds."""

    P: Any
    q: np.ndarray
    G: Optional[Any] = None
    h: Optional[np.ndarray] = None
    A: Optional[np.ndarray] = None
    b: Optional[np.ndarray] = None
    lb: Optional[np.ndarray] = None
    ub: Optional[np.ndarray] = None
