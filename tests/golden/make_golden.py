#!/usr/bin/env python3
"""Generate tests/golden/condense_*.npz by running the REFERENCE condensing.

Run in the build container only (``/root/reference`` does not exist on the GPU
box): ``python tests/golden/make_golden.py``.  The reference package needs a
module called ``qpsolvers`` at import time (``qpmpc/mpc_qp.py:13``,
``qpmpc/solve_mpc.py:9``, ``qpmpc/plan.py:13``); the real one is not installable
offline, so a three-name stub is put on ``sys.modules`` first.  Only the
reference's own NumPy condensing (``MPCQP.__init__``, ``update_cost_vector``,
``update_constraint_vector``) is executed -- no solver.

Each fixture stores the problem inputs (canonical per-step stacks) next to the
reference outputs P, q, G, h, Phi, Psi, phi_last, psi_last, e, so the tests can
rebuild the problem without the reference and compare field by field.
"""

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"


def import_reference():
    stub = types.ModuleType("qpsolvers")

    class Problem:  # what MPCQP.problem constructs (mpc_qp.py:127)
        def __init__(self, P, q, G=None, h=None, A=None, b=None, lb=None, ub=None):
            self.P, self.q, self.G, self.h = P, q, G, h

    class Solution:
        def __init__(self, problem=None):
            self.problem, self.found, self.x = problem, False, None

    def solve_problem(problem, solver=None, **kwargs):
        raise RuntimeError("stub qpsolvers: no solver in the golden generator")

    stub.Problem, stub.Solution, stub.solve_problem = Problem, Solution, solve_problem
    stub.available_solvers = []
    sys.modules["qpsolvers"] = stub
    sys.path.insert(0, REFERENCE)
    import qpmpc  # noqa: E402  (the reference)

    assert os.path.realpath(qpmpc.__file__).startswith(REFERENCE), qpmpc.__file__
    return qpmpc


def triple_integrator(N=16, horizon=1.0, x0=(0.0, 0.0, 0.0), goal=(1.0, 0.0, 0.0),
                      w_t=1.0, w_x=None, w_u=1e-6, max_accel=3.0):
    """Data of examples/triple_integrator.py:15-42."""
    T = horizon / N
    A = np.array([[1.0, T, T**2 / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]])
    B = np.array([T**3 / 6.0, T**2 / 2.0, T]).reshape((3, 1))
    C = np.array([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0]])
    e = np.array([max_accel, max_accel])
    return dict(A=A, B=B, C=C, D=None, e=e, N=N, w_t=w_t, w_x=w_x, w_u=w_u,
                x0=np.array(x0), goal=np.array(goal), targets=None)


def humanoid(N=16, start=0.0, end=0.3, foot=0.1, com_height=0.8, horizon=2.5,
             dsp=0.1, ssp=0.7):
    """Data of tests/test_humanoid_one_step.py:29-70 (LTV e_k)."""
    T = horizon / N
    n_dsp0, n_ssp0, n_dsp1 = (int(round(v / T)) for v in (dsp, ssp, dsp))
    A = np.array([[1.0, T, T**2 / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]])
    B = np.array([T**3 / 6.0, T**2 / 2.0, T]).reshape((3, 1))
    zmp = np.array([1.0, 0.0, -com_height / 9.81])
    C = np.array([+zmp, -zmp])
    e = []
    for i in range(N):
        if i < n_dsp0:
            e.append(np.array([1000.0, 1000.0]))
        elif i - n_dsp0 <= n_ssp0:
            e.append(np.array([start + 0.5 * foot, -(start - 0.5 * foot)]))
        elif i - n_dsp0 - n_ssp0 < n_dsp1:
            e.append(np.array([1000.0, 1000.0]))
        else:
            e.append(np.array([end + 0.5 * foot, -(end - 0.5 * foot)]))
    return dict(A=A, B=B, C=C, D=None, e=e, N=N, w_t=1.0, w_x=None, w_u=1e-3,
                x0=np.array([start, 0.0, 0.0]), goal=np.array([end, 0.0, 0.0]),
                targets=None)


def pendulum(qpmpc, x0=(0.0, 0.2, 0.0, 0.5), v_target=0.5):
    """WIP problem of examples/wheeled_inverted_pendulum.py:90-108."""
    from qpmpc.systems import WheeledInvertedPendulum

    wip = WheeledInvertedPendulum()
    prob = wip.build_mpc_problem(
        terminal_cost_weight=10.0, stage_state_cost_weight=1.0,
        stage_input_cost_weight=1e-3,
    )
    N, nx, T = wip.nb_timesteps, 4, wip.sampling_period
    ts = np.zeros((N + 1) * nx)
    for k in range(N + 1):
        ts[k * nx] = x0[0] + k * T * v_target
        ts[k * nx + 2] = v_target
    return dict(A=prob.transition_state_matrix, B=prob.transition_input_matrix,
                C=None, D=prob.ineq_input_matrix, e=prob.ineq_vector, N=N,
                w_t=10.0, w_x=1.0, w_u=1e-3, x0=np.array(x0), goal=ts[-nx:],
                targets=ts[:-nx])


def random_ltv(seed, N, nx, nu, nc, with_C=True, with_D=True, w_t=0.7, w_x=0.3,
               w_u=1e-2):
    """Fully time-varying random problem (lists for A, B, C, D, e)."""
    rng = np.random.default_rng(seed)
    A = [np.eye(nx) + 0.2 * rng.standard_normal((nx, nx)) for _ in range(N)]
    B = [rng.standard_normal((nx, nu)) for _ in range(N)]
    C = [rng.standard_normal((nc, nx)) for _ in range(N)] if with_C else None
    D = [rng.standard_normal((nc, nu)) for _ in range(N)] if with_D else None
    e = [1.0 + rng.random(nc) for _ in range(N)]
    return dict(A=A, B=B, C=C, D=D, e=e, N=N, w_t=w_t, w_x=w_x, w_u=w_u,
                x0=0.1 * rng.standard_normal(nx), goal=rng.standard_normal(nx),
                targets=rng.standard_normal(N * nx))


def to_reference_problem(qpmpc, d):
    prob = qpmpc.MPCProblem(
        transition_state_matrix=d["A"], transition_input_matrix=d["B"],
        ineq_state_matrix=d["C"], ineq_input_matrix=d["D"], ineq_vector=d["e"],
        nb_timesteps=d["N"], terminal_cost_weight=d["w_t"],
        stage_state_cost_weight=d["w_x"], stage_input_cost_weight=d["w_u"],
        initial_state=d["x0"], goal_state=d["goal"],
    )
    if d["targets"] is not None:
        prob.update_target_states(d["targets"])
    return prob


def stack(op, N):
    """Canonical storage: None -> empty, LTI -> (r, c), LTV list -> (N, r, c)."""
    if op is None:
        return np.zeros(0)
    return np.stack(op) if isinstance(op, list) else np.asarray(op, dtype=float)


def dump(name, qpmpc, d):
    prob = to_reference_problem(qpmpc, d)
    ref = qpmpc.MPCQP(prob)
    h_ctor = ref.h.copy()
    if d["C"] is not None:  # update_constraint_vector needs every C_k (quirk Q4)
        ref.update_constraint_vector(prob)
    N = d["N"]
    np.savez_compressed(
        os.path.join(HERE, f"condense_{name}.npz"),
        A=stack(d["A"], N), B=stack(d["B"], N), C=stack(d["C"], N),
        D=stack(d["D"], N), e=stack(d["e"], N), N=N,
        A_ltv=isinstance(d["A"], list), B_ltv=isinstance(d["B"], list),
        C_ltv=isinstance(d["C"], list), D_ltv=isinstance(d["D"], list),
        e_ltv=isinstance(d["e"], list),
        w_t=np.nan if d["w_t"] is None else d["w_t"],
        w_x=np.nan if d["w_x"] is None else d["w_x"], w_u=d["w_u"],
        x0=d["x0"], goal=d["goal"] if d["goal"] is not None else np.zeros(0),
        targets=d["targets"] if d["targets"] is not None else np.zeros(0),
        ref_P=ref.P, ref_q=ref.q, ref_G=ref.G, ref_h=h_ctor, ref_h_update=ref.h,
        ref_Phi=ref.Phi, ref_Psi=ref.Psi, ref_phi_last=ref.phi_last,
        ref_psi_last=ref.psi_last, ref_e=ref.e,
    )
    print(f"condense_{name}.npz  n={ref.q.size} m={ref.h.size}")


def main():
    qpmpc = import_reference()
    dump("triple_integrator", qpmpc, triple_integrator())
    dump("triple_integrator_stage", qpmpc,
         {**triple_integrator(w_x=0.5), "targets": np.linspace(0, 1, 48)})
    dump("triple_integrator_tiny_wt", qpmpc, triple_integrator(w_t=1e-12))  # quirk Q1
    dump("humanoid", qpmpc, humanoid())
    dump("pendulum", qpmpc, pendulum(qpmpc))
    dump("random_ltv_cd", qpmpc, random_ltv(0, N=6, nx=3, nu=2, nc=3))
    dump("random_ltv_c", qpmpc, random_ltv(1, N=5, nx=4, nu=1, nc=2, with_D=False, w_x=None))
    dump("random_ltv_d", qpmpc, random_ltv(2, N=7, nx=2, nu=2, nc=4, with_C=False, w_t=None))
    for N in (8, 32, 64):
        dump(f"triple_integrator_N{N}", qpmpc, triple_integrator(N=N))


if __name__ == "__main__":
    main()
