"""Result wrapper for one MPC solve.

Mirror of ``qpmpc/plan.py:18-109``: ``Plan(problem, qpsol)`` keeps the problem
and the solver result alive, exposes ``inputs`` as an (N, nu) array,
``first_input``, ``is_empty`` and a lazily integrated, memoised ``states``.
"""

import logging
from typing import Optional

import numpy as np

from .mpc_problem import MPCProblem


class Plan:
    """State and input trajectories that optimise an MPC problem.

    Attributes:
        problem: The MPC problem that was solved.
        qpsol: Solver result; only ``found`` and ``x`` are read.
    """

    def __init__(self, problem: MPCProblem, qpsol) -> None:
        self.problem = problem
        self.qpsol = qpsol
        self._states: Optional[np.ndarray] = None
        self._inputs: Optional[np.ndarray] = (
            np.asarray(qpsol.x).reshape(
                (problem.nb_timesteps, problem.input_dim)
            )
            if qpsol.found
            else None
        )

    @property
    def is_empty(self) -> bool:
        """True when the solver found no solution."""
        return self._inputs is None

    @property
    def inputs(self) -> Optional[np.ndarray]:
        """U as an (N, nu) array, or None for an empty plan."""
        return self._inputs

    @property
    def first_input(self) -> Optional[np.ndarray]:
        """u_0, the input a receding-horizon controller applies."""
        return None if self._inputs is None else self._inputs[0]

    @property
    def states(self) -> Optional[np.ndarray]:
        """X as an (N + 1, nx) array; O(N) on first access, then cached."""
        if self._inputs is None:
            return None
        if self._states is None:
            x_init = self.problem.initial_state
            if x_init is None:
                logging.warning("Problem has undefined initial state")
                return None
            self._states = self.problem.integrate(x_init, self._inputs)
        return self._states
