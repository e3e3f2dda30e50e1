"""Host-side container for one linear (time-varying) MPC problem.

Mirror of the reference data model (``qpmpc/mpc_problem.py:16-335``): same
constructor arguments, attribute names, accessors and error behaviour, so user
code that builds an ``MPCProblem`` for qpmpc builds one for this engine
unchanged.  No arithmetic of the hot path lives here; the container only
validates and hands arrays to the packer in :mod:`qpmpc_b200.batched`.

System:       x_{k+1} = A_k x_k + B_k u_k,          k = 0..N-1
Constraints:  C_k x_k + D_k u_k <= e_k
Cost:         w_t |x_N - goal|^2 + w_x sum_k |x_k - target_k|^2 + w_u sum_k |u_k|^2

``A, B, C, D, e`` are either one ndarray (time-invariant) or a Python ``list``
of N ndarrays (time-varying).  As in the reference the test is
``isinstance(x, list)`` (``mpc_problem.py:178-180``).
"""

from typing import List, Optional, Union

import numpy as np

from .exceptions import ProblemDefinitionError, StateError

ArrayOrList = Union[np.ndarray, List[np.ndarray]]
OptArrayOrList = Union[None, np.ndarray, List[np.ndarray]]


def _per_step(operand, k):
    """LTV operands are lists indexed by step; anything else is time-invariant."""
    return operand[k] if isinstance(operand, list) else operand


def _column_count(operand) -> int:
    first = operand if isinstance(operand, np.ndarray) else operand[0]
    return first.shape[1]


class MPCProblem:
    """Linear time-variant model predictive control problem.

    Attributes (names follow ``qpmpc/mpc_problem.py:75-88``):
        transition_state_matrix, transition_input_matrix: A_k, B_k.
        ineq_state_matrix, ineq_input_matrix, ineq_vector: C_k, D_k, e_k;
            ``None`` for C or D means the null matrix.
        nb_timesteps: N.  state_dim, input_dim: nx, nu.
        terminal_cost_weight, stage_state_cost_weight: w_t, w_x (``None``
            disables the term), stage_input_cost_weight: w_u > 0.
        initial_state, goal_state, target_states: flattened vectors or None.
    """

    def __init__(
        self,
        transition_state_matrix: ArrayOrList,
        transition_input_matrix: ArrayOrList,
        ineq_state_matrix: OptArrayOrList,
        ineq_input_matrix: OptArrayOrList,
        ineq_vector: ArrayOrList,
        nb_timesteps: int,
        terminal_cost_weight: Optional[float],
        stage_state_cost_weight: Optional[float],
        stage_input_cost_weight: float,
        initial_state: Optional[np.ndarray] = None,
        goal_state: Optional[np.ndarray] = None,
        target_states: Optional[np.ndarray] = None,
    ) -> None:
        # Validation order and messages' meaning: mpc_problem.py:104-111.
        if stage_input_cost_weight <= 0.0:
            raise ProblemDefinitionError(
                "the input weight must be positive: it regularizes the QP"
            )
        if terminal_cost_weight is None and stage_state_cost_weight is None:
            raise ProblemDefinitionError(
                "set a terminal cost weight, a stage state cost weight, or both"
            )
        self.transition_state_matrix = transition_state_matrix
        self.transition_input_matrix = transition_input_matrix
        self.ineq_state_matrix = ineq_state_matrix
        self.ineq_input_matrix = ineq_input_matrix
        self.ineq_vector = ineq_vector
        self.nb_timesteps = nb_timesteps
        self.state_dim = _column_count(transition_state_matrix)
        self.input_dim = _column_count(transition_input_matrix)
        self.terminal_cost_weight = terminal_cost_weight
        self.stage_state_cost_weight = stage_state_cost_weight
        self.stage_input_cost_weight = stage_input_cost_weight
        self.initial_state: Optional[np.ndarray] = None
        self.goal_state: Optional[np.ndarray] = None
        self.target_states: Optional[np.ndarray] = None
        if goal_state is not None:
            self.update_goal_state(goal_state)
        if initial_state is not None:
            self.update_initial_state(initial_state)
        # Quirk kept for parity: like the reference (mpc_problem.py:88-139) the
        # ``target_states`` argument is accepted but NOT stored; callers set it
        # with update_target_states(), as every reference example does.
        del target_states

    # -- cost-term predicates (mpc_problem.py:141-166) ----------------------

    _WEIGHT_FLOOR = 1e-10

    @property
    def has_terminal_cost(self) -> bool:
        """True if w_t is set above 1e-10; raises if the goal is then missing."""
        w = self.terminal_cost_weight
        active = w is not None and w > self._WEIGHT_FLOOR
        if active and self.goal_state is None:
            raise ProblemDefinitionError(
                "a terminal cost is set but there is no goal state"
            )
        return active

    @property
    def has_stage_state_cost(self) -> bool:
        """True if w_x is set above 1e-10; raises if targets are then missing."""
        w = self.stage_state_cost_weight
        active = w is not None and w > self._WEIGHT_FLOOR
        if active and self.target_states is None:
            raise ProblemDefinitionError(
                "a stage state cost is set but there are no target states"
            )
        return active

    # -- per-step accessors (mpc_problem.py:168-245) ------------------------

    def get_transition_state_matrix(self, k) -> np.ndarray:
        """A_k."""
        return _per_step(self.transition_state_matrix, k)

    def get_transition_input_matrix(self, k) -> np.ndarray:
        """B_k."""
        return _per_step(self.transition_input_matrix, k)

    def get_ineq_state_matrix(self, k) -> Optional[np.ndarray]:
        """C_k, or None."""
        return _per_step(self.ineq_state_matrix, k)

    def get_ineq_input_matrix(self, k) -> Optional[np.ndarray]:
        """D_k, or None."""
        return _per_step(self.ineq_input_matrix, k)

    def get_ineq_vector(self, k) -> np.ndarray:
        """e_k."""
        return _per_step(self.ineq_vector, k)

    # -- state setters (mpc_problem.py:247-295): size check, then flatten ---

    def _checked(self, vec: np.ndarray, size: int, what: str) -> np.ndarray:
        vec = np.asarray(vec)
        if vec.size != size:
            raise StateError(
                f"{what} has shape {vec.shape}, expected {size} entries"
            )
        return vec.flatten()

    def update_goal_state(self, goal_state: np.ndarray) -> None:
        """Set the goal (terminal) state."""
        self.goal_state = self._checked(goal_state, self.state_dim, "goal state")

    def update_initial_state(self, initial_state: np.ndarray) -> None:
        """Set the initial state x_0."""
        self.initial_state = self._checked(
            initial_state, self.state_dim, "initial state"
        )

    def update_target_states(self, target_states: np.ndarray) -> None:
        """Set the reference trajectory for x_0..x_{N-1} (N * nx entries)."""
        self.target_states = self._checked(
            target_states,
            self.state_dim * self.nb_timesteps,
            "reference state trajectory (nb_timesteps * state_dim)",
        )

    def __repr__(self) -> str:
        fields = (
            "goal_state ineq_input_matrix ineq_state_matrix ineq_vector "
            "initial_state input_dim nb_timesteps stage_input_cost_weight "
            "stage_state_cost_weight state_dim terminal_cost_weight "
            "transition_input_matrix transition_state_matrix"
        ).split()
        body = ", ".join(f"{name}={getattr(self, name)}" for name in fields)
        return f"MPCProblem({body})"

    def integrate(self, initial_state: np.ndarray, inputs: np.ndarray) -> np.ndarray:
        """Roll the dynamics forward: returns X of shape (N + 1, nx).

        Host-side convenience for a single plan (``mpc_problem.py:316-335``).
        Batched plans integrate on the device instead.
        """
        N = self.nb_timesteps
        X = np.zeros((N + 1, self.state_dim))
        X[0] = initial_state
        for k in range(N):
            A_k = self.get_transition_state_matrix(k)
            B_k = self.get_transition_input_matrix(k)
            X[k + 1] = A_k.dot(X[k]) + B_k.dot(inputs[k])
        return X
