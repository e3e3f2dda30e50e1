"""Condensed QP of one MPC problem, computed by the CUDA condensing kernel.

Mirror of ``qpmpc/mpc_qp.py:21-163`` (reference tree): same constructor,
public fields (``P, q, G, h, Phi, Psi, phi_last, psi_last, e, C``) and update
methods.  The arithmetic runs in ``mpc_condense_kernel`` through the C ABI
(``qpmpc_b200_condense``); this class only moves arrays.  The constructor and
``update_constraint_vector`` obtain ``h`` from the same kernel, so the two are
bit-identical (what ``tests/test_update_constraint_vector.py:71-80`` of the
reference asserts).
"""

import logging
from typing import Optional

import numpy as np

from .batched import condense_batch, problem_to_batch
from .exceptions import ProblemDefinitionError
from .mpc_problem import MPCProblem
from .solution import QPProblem


class MPCQP:
    """Dense QP ``min 1/2 U'PU + q'U  s.t.  G U <= h`` of an MPC problem.

    Attributes:
        P, q: Cost matrix (n x n) and vector (n), n = N * nu.
        G, h: Inequality matrix (m x n) and vector (m).
        Phi, Psi: Stacked state maps, x_k = Phi_k x_0 + Psi_k U for k < N.
        phi_last, psi_last: The same maps for k = N.
        e, C: Stacked inequality vector and block-diagonal state matrix.
    """

    def __init__(self, mpc_problem: MPCProblem, sparse: bool = False) -> None:
        if mpc_problem.initial_state is None:
            raise ProblemDefinitionError("initial state is undefined")
        batch = problem_to_batch(mpc_problem)
        self._rows = np.asarray(batch.row_map, dtype=np.int64)
        fields = condense_batch(
            batch, ("P", "q", "G", "h", "Phi", "Psi", "phi_last", "psi_last"))
        host = {k: v[0].cpu().numpy() for k, v in fields.items()}
        self.P = host["P"]
        self.G = host["G"][self._rows]
        self.h = host["h"][self._rows]
        self.q = host["q"]
        self.Phi, self.Psi = host["Phi"], host["Psi"]
        self.phi_last, self.psi_last = host["phi_last"], host["psi_last"]
        N, nx = mpc_problem.nb_timesteps, mpc_problem.state_dim
        e_steps = [np.asarray(mpc_problem.get_ineq_vector(k), dtype=float).reshape(-1)
                   for k in range(N)]
        self.e = np.hstack(e_steps)
        self.C = self._stack_state_blocks(mpc_problem, e_steps, N, nx)
        if self.h.size and np.any(self.h[: e_steps[0].size] < 0.0) and (
            mpc_problem.get_ineq_input_matrix(0) is None
        ):
            # mpc_qp.py:79-85: the inputs cannot repair a violated k = 0 row
            logging.warning(
                "initial state is unfeasible: the first inequality rows are "
                "violated and do not depend on the inputs"
            )
        if sparse:
            from scipy.sparse import csc_matrix  # mpc_qp.py:108-109

            self.P = csc_matrix(self.P)
            self.G = csc_matrix(self.G)

    @staticmethod
    def _stack_state_blocks(problem, e_steps, N, nx) -> Optional[np.ndarray]:
        """Block-diagonal C (pure data placement, no arithmetic)."""
        if any(problem.get_ineq_state_matrix(k) is None for k in range(N)):
            return None
        m = sum(ek.size for ek in e_steps)
        C = np.zeros((m, N * nx))
        row = 0
        for k in range(N):
            blk = np.asarray(problem.get_ineq_state_matrix(k), dtype=float)
            C[row:row + blk.shape[0], k * nx:(k + 1) * nx] = blk
            row += blk.shape[0]
        return C

    @property
    def problem(self) -> QPProblem:
        """The QP as a (P, q, G, h) record (``mpc_qp.py:124-127``)."""
        return QPProblem(self.P, self.q, self.G, self.h)

    def update_cost_vector(self, mpc_problem: MPCProblem) -> None:
        """Recompute q for a new initial / goal / target state
        (``mpc_qp.py:129-149``)."""
        if mpc_problem.initial_state is None:
            raise ProblemDefinitionError("initial state is undefined")
        # Evaluating the predicates raises exactly where the reference does.
        mpc_problem.has_terminal_cost
        mpc_problem.has_stage_state_cost
        batch = problem_to_batch(mpc_problem)
        self.q = condense_batch(batch, ("q",))["q"][0].cpu().numpy()

    def update_constraint_vector(self, mpc_problem: MPCProblem) -> None:
        """Recompute h for a new initial state (``mpc_qp.py:151-163``)."""
        if mpc_problem.initial_state is None:
            raise ProblemDefinitionError("initial state is undefined")
        batch = problem_to_batch(mpc_problem)
        h = condense_batch(batch, ("h",))["h"][0].cpu().numpy()
        self.h = h[self._rows]
