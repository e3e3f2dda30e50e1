"""The reference's own test cases (tests/test_humanoid_one_step.py:72-85,
tests/test_update_constraint_vector.py:71-80,
tests/test_wheeled_inverted_pendulum.py:19-41 of stephane-caron/qpmpc), restated
against this package's mirror of the reference surface with the CUDA engine as
the backend -- same problems, same assertions -- plus the API edge cases the
reference's conventions define (Plan on an unsolved problem, solver names)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _humanoid_problem():
    """The fixture both humanoid test files of the reference build (N=16,
    T=2.5/16, com height 0.8, dsp 0.1 / ssp 0.7, step 0 -> 0.3, foot 0.1)."""
    from qpmpc_b200 import MPCProblem

    N, horizon, h, g = 16, 2.5, 0.8, 9.81
    T = horizon / N
    A = np.array([[1.0, T, T**2 / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]])
    B = np.array([T**3 / 6.0, T**2 / 2.0, T]).reshape((3, 1))
    zmp = np.array([1.0, 0.0, -h / g])
    C = np.array([+zmp, -zmp])
    n0, n1, n2 = int(round(0.1 / T)), int(round(0.7 / T)), int(round(0.1 / T))
    start, end, foot = 0.0, 0.3, 0.1
    e = []
    for i in range(N):
        if i < n0 or (i - n0 > n1 and i - n0 - n1 < n2):
            e.append(np.array([1000.0, 1000.0]))
        elif i - n0 <= n1:
            e.append(np.array([start + 0.5 * foot, -(start - 0.5 * foot)]))
        else:
            e.append(np.array([end + 0.5 * foot, -(end - 0.5 * foot)]))
    return MPCProblem(
        transition_state_matrix=A, transition_input_matrix=B, ineq_state_matrix=C,
        ineq_input_matrix=None, ineq_vector=e, initial_state=np.array([start, 0.0, 0.0]),
        goal_state=np.array([end, 0.0, 0.0]), nb_timesteps=N, terminal_cost_weight=1.0,
        stage_state_cost_weight=None, stage_input_cost_weight=1e-3)


def test_mpc_qp_first_constraints():
    """Without D the k = 0 rows of G do not depend on the inputs."""
    from qpmpc_b200.solve_mpc import MPCQP

    mpc_qp = MPCQP(_humanoid_problem())
    assert np.linalg.norm(mpc_qp.G[0:2]) < 1e-10


def test_solve_mpc_shapes():
    from qpmpc import solve_mpc  # the compat alias opts in to serving qpsolvers names

    problem = _humanoid_problem()
    plan = solve_mpc(problem, solver="proxqp")  # what tests/test_humanoid_one_step.py:78 passes
    assert plan.qpsol.extras["method"] == "active_set"
    N = problem.nb_timesteps
    assert plan.inputs.flatten().shape == (N * problem.input_dim,)
    assert plan.states.flatten().shape == ((N + 1) * problem.state_dim,)


def test_update_constraint_vector_is_bit_identical():
    from qpmpc_b200 import MPCQP

    problem = _humanoid_problem()
    mpc_qp = MPCQP(problem)
    h_before = mpc_qp.h.copy()
    mpc_qp.update_constraint_vector(problem)
    np.testing.assert_array_equal(h_before, mpc_qp.h)
    problem.update_initial_state(np.array([0.01, 0.1, 0.0]))
    mpc_qp.update_constraint_vector(problem)
    assert np.abs(mpc_qp.h - MPCQP(problem).h).max() == 0.0
    mpc_qp.update_cost_vector(problem)
    assert np.abs(mpc_qp.q - MPCQP(problem).q).max() == 0.0


def test_pendulum_properties_and_zero_problem():
    from qpmpc_b200 import solve_mpc
    from qpmpc_b200.systems import WheeledInvertedPendulum

    pendulum = WheeledInvertedPendulum()
    assert pendulum.horizon_duration > 0.1 and pendulum.omega > 0.1
    problem = pendulum.build_mpc_problem(terminal_cost_weight=10.0, stage_state_cost_weight=1.0,
                                         stage_input_cost_weight=1e-3)
    x0 = np.zeros(pendulum.STATE_DIM)
    problem.update_initial_state(x0)
    problem.update_goal_state(x0.copy())
    problem.update_target_states(np.zeros(pendulum.nb_timesteps * pendulum.STATE_DIM))
    plan = solve_mpc(problem, solver="b200")
    assert plan is not None and plan.first_input is not None
    state = pendulum.integrate(x0, plan.first_input, pendulum.sampling_period)
    assert np.allclose(state, x0)


def test_unsolved_problem_gives_an_empty_plan():
    """qpsol.found False -> Plan.is_empty, getters None (plan.py:36,45-62)."""
    from qpmpc_b200 import solve_mpc

    problem = _humanoid_problem()
    # the k = 0 rows do not depend on the inputs (no D): a ZMP outside them cannot be repaired
    problem.ineq_vector[0] = np.array([0.05, 0.05])
    problem.update_initial_state(np.array([5.0, 0.0, 0.0]))
    plan = solve_mpc(problem, solver="b200")
    assert plan.is_empty and plan.first_input is None and plan.inputs is None


def test_single_step_and_unconstrained_problems():
    """N = 1 and a problem whose constraints never bind (U = -P^-1 q)."""
    import oracle
    from qpmpc_b200 import MPCProblem, MPCQP, solve_mpc

    A = np.array([[1.0, 0.1], [0.0, 1.0]])
    B = np.array([[0.005], [0.1]])
    for N in (1, 5):
        problem = MPCProblem(A, B, None, np.array([[1.0], [-1.0]]), np.array([1e3, 1e3]), N,
                             terminal_cost_weight=1.0, stage_state_cost_weight=0.1,
                             stage_input_cost_weight=1e-2, initial_state=np.array([1.0, 0.0]),
                             goal_state=np.zeros(2), target_states=np.zeros(2 * N))
        plan = solve_mpc(problem, solver="b200")
        qp = MPCQP(problem)
        expect = -np.linalg.solve(qp.P, qp.q)
        assert np.abs(plan.inputs.flatten() - expect).max() <= 1e-9
        st, x, _, _ = oracle.qp_gi(qp.P, qp.q, qp.G, qp.h)
        assert st == 0 and np.abs(plan.inputs.flatten() - x).max() <= 1e-9
        assert plan.qpsol.extras["iters"] == 0


def test_solver_name_policy():
    """``solver=`` keeps its reference meaning (solve_mpc.py:43): native names run the CUDA
    engine, qpsolvers names need qpsolvers or the documented opt-in, anything else raises;
    keywords that do not apply are dropped with a warning, never silently."""
    import importlib.util

    from qpmpc_b200 import BackendError, solve_mpc
    from qpmpc_b200.solve_mpc import serve_qpsolvers_names

    problem = _humanoid_problem()
    with pytest.raises(BackendError, match="unknown solver"):
        solve_mpc(problem, solver="no_such_backend")
    try:
        serve_qpsolvers_names(False)
        if importlib.util.find_spec("qpsolvers") is None:
            with pytest.raises(BackendError, match="qpsolvers"):
                solve_mpc(problem, solver="osqp")
        serve_qpsolvers_names(True)
        with pytest.warns(UserWarning, match="initvals"):
            exact = solve_mpc(problem, solver="quadprog", initvals=np.zeros(16))
        ipm = solve_mpc(problem, solver="clarabel", eps_abs=1e-9)  # interior-point name -> pdip kernel
        assert ipm.qpsol.extras["method"] == "pdip" and exact.qpsol.extras["method"] == "active_set"
        assert np.abs(ipm.inputs - exact.inputs).max() <= 1e-6
        assert np.abs(solve_mpc(problem, solver="b200_pdip").inputs - exact.inputs).max() <= 1e-6
    finally:
        serve_qpsolvers_names(True)


def test_mpcqp_problem_record_and_sparse():
    """``MPCQP.problem`` (mpc_qp.py:124-127) and ``sparse=True`` (mpc_qp.py:108-109): P and G
    become csc matrices with the dense entries, q and h stay arrays, the record has no
    equalities or bounds, and solving the record's QP reproduces the plan."""
    import oracle
    import scipy.sparse

    from qpmpc_b200 import MPCQP, solve_mpc

    problem = _humanoid_problem()
    dense, sparse = MPCQP(problem), MPCQP(problem, sparse=True)
    assert scipy.sparse.isspmatrix_csc(sparse.P) and scipy.sparse.isspmatrix_csc(sparse.G)
    assert isinstance(dense.P, np.ndarray) and isinstance(dense.G, np.ndarray)
    np.testing.assert_array_equal(sparse.P.toarray(), dense.P)
    np.testing.assert_array_equal(sparse.G.toarray(), dense.G)
    np.testing.assert_array_equal(sparse.q, dense.q)
    np.testing.assert_array_equal(sparse.h, dense.h)
    qp = dense.problem
    assert qp.P is dense.P and qp.q is dense.q and qp.G is dense.G and qp.h is dense.h
    assert qp.A is None and qp.b is None and qp.lb is None and qp.ub is None
    assert sparse.problem.P is sparse.P and sparse.problem.G is sparse.G
    st, x, _, _ = oracle.qp_gi(qp.P, qp.q, qp.G, qp.h)
    assert st == 0
    assert np.abs(solve_mpc(problem, solver="b200", sparse=True).inputs.flatten() - x).max() <= 1e-6
