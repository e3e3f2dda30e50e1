"""NumPy restatement of the reference condensing.  TEST INFRASTRUCTURE ONLY.

Follows ``qpmpc/mpc_qp.py:39-163`` (reference tree) step for step, on any
object offering the ``MPCProblem`` accessors, and returns plain arrays.  Pinned
against the reference's own ``MPCQP`` by ``tests/golden/`` (see
``tests/golden/make_golden.py``).

With x_k = phi_k x0 + psi_k U  (phi_0 = I, psi_0 = 0):
    G_k = C_k psi_k + [0 .. D_k .. 0]          (mpc_qp.py:67,73-78)
    h_k = e_k - (C_k phi_k) x0                 (mpc_qp.py:68-72)
    phi_{k+1} = A_k phi_k; psi_{k+1} = A_k psi_k, block column k := B_k  (:88-90)
    P = w_u I + [w_t set] w_t psi_N'psi_N + [w_x set] w_x Psi'Psi        (:99-105)
    q = [w_t>1e-10] w_t psi_N'(phi_N x0 - goal)
      + [w_x>1e-10] w_x Psi'(Phi x0 - targets)                          (:139-149)
"""

from typing import Dict, Optional

import numpy as np

WEIGHT_FLOOR = 1e-10  # qpmpc/mpc_problem.py:146,159


def cost_vector(
    phi_last, psi_last, Phi, Psi, x0, goal, targets, w_t, w_x
) -> np.ndarray:
    """q as ``update_cost_vector`` builds it, including quirk Q2:
    a missing goal / target trajectory aborts accumulation where it is met."""
    q = np.zeros(psi_last.shape[1])
    if w_t is not None and w_t > WEIGHT_FLOOR:
        if goal is None:
            return q
        q = q + w_t * (phi_last @ x0 - goal) @ psi_last
    if w_x is not None and w_x > WEIGHT_FLOOR:
        if targets is None:
            return q
        q = q + w_x * (Phi @ x0 - targets) @ Psi
    return q


def constraint_vector(e, C_steps, Phi, x0, nx) -> np.ndarray:
    """h = e - blockdiag(C) Phi x0, row block by row block (mpc_qp.py:161-163)."""
    h, row = np.array(e, dtype=float), 0
    for k, C_k in enumerate(C_steps):
        nc_k = C_k.shape[0]
        h[row:row + nc_k] -= (C_k @ Phi[k * nx:(k + 1) * nx]) @ x0
        row += nc_k
    return h


def condense(problem) -> Dict[str, Optional[np.ndarray]]:
    """All fields of the reference ``MPCQP`` for one problem."""
    N, nx, nu = problem.nb_timesteps, problem.state_dim, problem.input_dim
    n = N * nu
    x0 = problem.initial_state
    if x0 is None:
        raise ValueError("initial state is undefined")  # mpc_qp.py:49-51
    phi, psi = np.eye(nx), np.zeros((nx, n))
    rows_G, rows_h, rows_e, phis, psis, Cs = [], [], [], [], [], []
    for k in range(N):
        A_k = problem.get_transition_state_matrix(k)
        B_k = problem.get_transition_input_matrix(k)
        C_k = problem.get_ineq_state_matrix(k)
        D_k = problem.get_ineq_input_matrix(k)
        e_k = np.asarray(problem.get_ineq_vector(k), dtype=float)
        phis.append(phi)
        psis.append(psi)
        G_k = np.zeros((e_k.shape[0], n))
        if D_k is not None:
            G_k[:, k * nu:(k + 1) * nu] = D_k
        if C_k is not None:
            G_k = G_k + C_k @ psi
            h_k = e_k - (C_k @ phi) @ x0
        else:
            h_k = e_k
        rows_G.append(G_k)
        rows_h.append(h_k)
        rows_e.append(e_k)
        Cs.append(C_k)
        phi = A_k @ phi
        psi = A_k @ psi
        psi[:, k * nu:(k + 1) * nu] = B_k
    G = np.vstack(rows_G).astype(float)
    h = np.hstack(rows_h).astype(float)
    Phi = np.vstack(phis).astype(float)
    Psi = np.vstack(psis).astype(float)
    e = np.hstack(rows_e).astype(float)
    w_t, w_x = problem.terminal_cost_weight, problem.stage_state_cost_weight
    P = problem.stage_input_cost_weight * np.eye(n)
    if w_t is not None:
        P = P + w_t * (psi.T @ psi)
    if w_x is not None:
        P = P + w_x * (Psi.T @ Psi)
    q = cost_vector(
        phi, psi, Phi, Psi, x0, problem.goal_state, problem.target_states, w_t, w_x
    )
    return dict(P=P, q=q, G=G, h=h, Phi=Phi, Psi=Psi, phi_last=phi,
                psi_last=psi, e=e, C_steps=Cs)
