"""Host-side logic that needs no GPU: the MPCProblem / Plan mirror, problem
packing, workload byte counts, the C-ABI library (loads, exports every symbol
the header declares) and the world-size-2 gloo path of the sharding helpers."""

import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_problem, load_golden
from qpmpc_b200 import MPCProblem, Plan, ProblemDefinitionError, Solution, StateError
from qpmpc_b200.systems import WheeledInvertedPendulum


def _ti_problem(**kw):
    g = load_golden("triple_integrator")
    args = dict(transition_state_matrix=g["A"], transition_input_matrix=g["B"],
                ineq_state_matrix=g["C"], ineq_input_matrix=None, ineq_vector=g["e"],
                nb_timesteps=16, terminal_cost_weight=1.0, stage_state_cost_weight=None,
                stage_input_cost_weight=1e-6)
    args.update(kw)
    return MPCProblem(**args)


# -- MPCProblem (qpmpc/mpc_problem.py) ------------------------------------------

def test_weight_validation():
    with pytest.raises(ProblemDefinitionError):
        _ti_problem(stage_input_cost_weight=0.0)
    with pytest.raises(ProblemDefinitionError):
        _ti_problem(terminal_cost_weight=None, stage_state_cost_weight=None)


def test_dimensions_and_state_setters():
    p = _ti_problem()
    assert (p.state_dim, p.input_dim, p.nb_timesteps) == (3, 1, 16)
    p.update_initial_state(np.zeros((3, 1)))
    assert p.initial_state.shape == (3,)
    with pytest.raises(StateError):
        p.update_initial_state(np.zeros(4))
    with pytest.raises(StateError):
        p.update_goal_state(np.zeros(2))
    with pytest.raises(StateError):
        p.update_target_states(np.zeros(3 * 15))
    p.update_target_states(np.zeros((16, 3)))
    assert p.target_states.shape == (48,)


def test_cost_predicates_follow_reference_thresholds():
    p = _ti_problem(terminal_cost_weight=1e-12, stage_state_cost_weight=0.5)
    assert p.has_terminal_cost is False  # below 1e-10: no q term (quirk Q1)
    with pytest.raises(ProblemDefinitionError):
        p.has_stage_state_cost  # targets missing
    p2 = _ti_problem()
    with pytest.raises(ProblemDefinitionError):
        p2.has_terminal_cost  # goal missing


def test_ltv_accessors_use_lists_only():
    g = load_golden("random_ltv_cd")
    p = golden_problem(g)
    assert p.get_transition_state_matrix(3) is p.transition_state_matrix[3]
    lti = _ti_problem()
    assert lti.get_ineq_state_matrix(5) is lti.ineq_state_matrix
    assert lti.get_ineq_input_matrix(0) is None


def test_integrate_and_plan():
    g = load_golden("triple_integrator")
    p = golden_problem(g)
    U = np.linspace(-1, 1, 16).reshape(16, 1)
    X = p.integrate(g["x0"], U)
    assert X.shape == (17, 3)
    x = g["x0"].copy()
    for k in range(16):
        x = g["A"] @ x + g["B"] @ U[k]
    assert np.allclose(X[-1], x, rtol=0, atol=1e-15)
    plan = Plan(p, Solution(found=True, x=U.flatten()))
    assert not plan.is_empty and plan.inputs.shape == (16, 1)
    assert plan.first_input.shape == (1,) and plan.states is plan.states
    assert np.array_equal(plan.states, X)
    empty = Plan(p, Solution(found=False))
    assert empty.is_empty and empty.inputs is None and empty.first_input is None
    assert empty.states is None


def test_pendulum_system_properties():
    """tests/test_wheeled_inverted_pendulum.py:19-21 of the reference."""
    wip = WheeledInvertedPendulum()
    assert wip.horizon_duration > 0.1 and wip.omega > 0.1
    prob = wip.build_mpc_problem(terminal_cost_weight=10.0, stage_state_cost_weight=1.0,
                                 stage_input_cost_weight=1e-3)
    g = load_golden("pendulum")
    assert np.allclose(prob.transition_state_matrix, g["A"], rtol=0, atol=1e-15)
    assert np.allclose(prob.transition_input_matrix, g["B"], rtol=0, atol=1e-15)
    s = wip.integrate(np.zeros(4), 0.0, 0.01)
    assert np.allclose(s, np.zeros(4))


# -- packing ---------------------------------------------------------------------

def test_pack_problem_layouts():
    from qpmpc_b200.batched import pack_problem

    lti = pack_problem(golden_problem(load_golden("triple_integrator")))
    assert lti["A"].shape == (3, 3) and lti["C"].shape == (2, 3) and lti["D"] is None
    assert lti["e"].shape == (2,) and lti["row_map"] == list(range(32))
    ltv = pack_problem(golden_problem(load_golden("random_ltv_cd")))
    assert ltv["A"].shape == (6, 3, 3) and ltv["D"].shape == (6, 3, 2) and ltv["e"].shape == (6, 3)
    hum = pack_problem(golden_problem(load_golden("humanoid")))
    assert hum["e"].shape == (16, 2) and hum["C"].shape == (2, 3)


def test_pack_problem_pads_ragged_rows():
    from qpmpc_b200.batched import pack_problem

    A, B = np.eye(2), np.ones((2, 1))
    C = [np.ones((1, 2)), np.ones((3, 2)), None]
    e = [np.ones(1), np.ones(3), np.ones(2)]
    D = [None, None, np.ones((2, 1))]
    p = MPCProblem(A, B, C, D, e, 3, 1.0, None, 1e-2, initial_state=np.zeros(2),
                   goal_state=np.ones(2))
    pk = pack_problem(p)
    assert pk["nc"] == 3 and pk["C"].shape == (3, 3, 2) and pk["e"].shape == (3, 3)
    assert pk["row_map"] == [0, 3, 4, 5, 6, 7]
    assert pk["e"][0, 1] == 1.0 and pk["e"][2, 2] == 1.0  # padding rows: 0 . u <= 1
    assert np.all(pk["C"][0, 1:] == 0.0) and np.all(pk["C"][2] == 0.0)


def test_workload_algorithmic_bytes_match_survey():
    from qpmpc_b200.workloads import (algorithmic_bytes_per_solve, humanoid_batch,
                                      pendulum_batch, triple_integrator_batch)

    assert algorithmic_bytes_per_solve(triple_integrator_batch(4)) == 340       # config 2
    assert algorithmic_bytes_per_solve(pendulum_batch(4)) == 548                # config 3
    for N, b in ((8, 276), (32, 468), (64, 724)):                               # config 5
        assert algorithmic_bytes_per_solve(triple_integrator_batch(4, N=N)) == b
    assert algorithmic_bytes_per_solve(humanoid_batch(4), itemsize=4) == 4 * (32 + 3 + 3) + 64 + 4


def test_humanoid_workload_reproduces_reference_bounds():
    """e_k pattern of tests/test_humanoid_one_step.py:49-60 (golden fixture)."""
    from qpmpc_b200.workloads import humanoid_batch

    g = load_golden("humanoid")
    w = humanoid_batch(3)
    big = g["e"] > 999
    assert np.array_equal(w["e"][0] > 999, big)
    assert np.allclose(w["e"][1, ~big[:, 0], 0] - 0.05, np.where(
        np.arange(16)[~big[:, 0]] <= 5, w["x0"][1, 0], w["goal"][1, 0]))


# -- C ABI --------------------------------------------------------------------------

def test_library_exports_every_declared_symbol():
    from qpmpc_b200 import _capi
    from qpmpc_b200.build import build_library

    build_library()
    lib = _capi.load()
    header = open(os.path.join(ROOT, "include", "qpmpc_b200.h")).read()
    declared = set(re.findall(r"\b(qpmpc_b200_[a-z0-9_]+)\s*\(", header))
    assert declared and declared == set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.qpmpc_b200_version() == 100
    assert b"weights" in lib.qpmpc_b200_strerror(-3)
    assert lib.qpmpc_b200_max_vars(0) >= 32 and lib.qpmpc_b200_max_rows(0, 16) >= 32


def test_descriptor_struct_matches_header_size():
    from qpmpc_b200 import _capi

    # 16 int32 + 3 double + 2 int32 + double + 2 int32, no padding surprises
    assert ctypes.sizeof(_capi.Desc) == 16 * 4 + 3 * 8 + 2 * 4 + 8 + 2 * 4
    assert ctypes.sizeof(_capi.Operands) == 8 * 8 and ctypes.sizeof(_capi.Outputs) == 4 * 8


def test_c_abi_rejects_bad_arguments_without_a_device():
    from qpmpc_b200 import _capi

    lib = _capi.load()
    d = _capi.Desc()
    ops, outs = _capi.Operands(), _capi.Outputs()
    assert lib.qpmpc_b200_solve(None, None, None, None) == -1
    d.batch, d.N, d.nx, d.nu, d.nc = 4, 16, 3, 1, 2
    d.w_u = 0.0
    assert lib.qpmpc_b200_solve(ctypes.byref(d), ctypes.byref(ops), ctypes.byref(outs), None) == -3
    d.w_u, d.has_wt, d.w_t = 1e-3, 1, 1.0
    assert lib.qpmpc_b200_solve(ctypes.byref(d), ctypes.byref(ops), ctypes.byref(outs), None) == -1
    # the other entry points validate before touching a device, too
    peers, loop = _capi.Peers(), _capi.ClosedLoop()
    assert lib.qpmpc_b200_solve_scatter(ctypes.byref(d), ctypes.byref(ops), ctypes.byref(outs), None, None) == -1
    peers.count = 9
    assert lib.qpmpc_b200_solve_scatter(ctypes.byref(d), ctypes.byref(ops), ctypes.byref(outs),
                                        ctypes.byref(peers), None) == -1
    assert lib.qpmpc_b200_pendulum_closed_loop(ctypes.byref(d), ctypes.byref(ops), ctypes.byref(outs),
                                               ctypes.byref(loop), None) == -2  # needs nx = 4, nu = 1
    assert lib.qpmpc_b200_condense(ctypes.byref(d), ctypes.byref(ops), None, None) == -1
    assert lib.qpmpc_b200_integrate(ctypes.byref(d), ctypes.byref(ops), None, None, None) == -1
    assert lib.qpmpc_b200_fp64_peak(0, None) == -1
    assert b"invalid argument" in lib.qpmpc_b200_strerror(-1)
    assert lib.qpmpc_b200_max_vars(0) >= 64 and lib.qpmpc_b200_max_rows(0, 64) >= 128  # N = 64 sweep point


def test_product_path_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from qpmpc_b200 import BackendError, solve_mpc

    with pytest.raises(BackendError):
        solve_mpc(golden_problem(load_golden("triple_integrator")), solver="b200")


def test_qpmpc_alias_exposes_the_reference_import_paths():
    """The imports of the reference's tests (tests/test_humanoid_one_step.py:12-13,
    tests/test_wheeled_inverted_pendulum.py:11-12) resolve to this package."""
    import qpmpc
    import qpmpc_b200
    from qpmpc import MPCProblem, MPCQP, Plan, solve_mpc
    from qpmpc.exceptions import ProblemDefinitionError, StateError
    from qpmpc.mpc_problem import MPCProblem as P2
    from qpmpc.solve_mpc import MPCQP as Q2
    from qpmpc.systems import WheeledInvertedPendulum

    assert MPCProblem is qpmpc_b200.MPCProblem is P2 and MPCQP is qpmpc_b200.MPCQP is Q2
    assert Plan is qpmpc_b200.Plan and solve_mpc is qpmpc_b200.solve_mpc and callable(qpmpc.solve_mpc)
    assert WheeledInvertedPendulum is qpmpc_b200.systems.WheeledInvertedPendulum
    assert issubclass(ProblemDefinitionError, qpmpc_b200.QPMPCException) and StateError is qpmpc_b200.StateError
    assert qpmpc.__all__ == ["MPCProblem", "MPCQP", "Plan", "solve_mpc"]  # qpmpc/__init__.py:14-19


def test_solver_names_are_checked_before_any_device_work():
    from qpmpc_b200 import BackendError, solve_mpc

    with pytest.raises(BackendError, match="unknown solver"):
        solve_mpc(golden_problem(load_golden("triple_integrator")), solver="not_a_backend")


# -- sharding over ranks (gloo, world size 2) -----------------------------------------

def test_shard_bounds_cover_the_batch():
    from qpmpc_b200.distributed import shard_bounds

    for batch, world in ((65536, 8), (10, 4), (3, 8), (1, 1)):
        spans = [shard_bounds(batch, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, batch, out):
    import torch
    import torch.distributed as dist

    from qpmpc_b200.distributed import gather_plans, shard_bounds

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(batch, rank, world)
    full = torch.arange(batch * 4, dtype=torch.float64).reshape(batch, 4)
    status = (torch.arange(batch) % 3 == 0).to(torch.int32)
    U, st = gather_plans(full[lo:hi].clone(), status[lo:hi].clone(), batch)
    ok = bool(torch.equal(U, full) and torch.equal(st, status))
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [64, 7])
def test_gather_plans_world_size_two(batch):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() + batch) % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gloo_worker, args=(2, port, batch, out), nprocs=2, join=True)
        assert out[0] and out[1]


def test_product_library_is_not_a_host_emulation_build():
    """QPMPC_HOST_EMU (host bodies for the inline-PTX helpers, used by tests/emu
    only) never reaches the product: the build has no such flag, the library
    exports nothing of the emulator and still contains sm_100a device code."""
    import subprocess

    from qpmpc_b200 import build

    assert not any("HOST_EMU" in f for f in build.NVCC_FLAGS)
    syms = subprocess.run(["nm", "-D", "--defined-only", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "emu_" not in syms and "pdip_emu" not in syms
    elf = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf


def test_batched_front_end_validates_its_arguments(emulated_engine):
    """``out=`` must match the problem (shape, dtype, device, contiguity); ``from_problems``
    refuses problems that differ in weights or row pattern (they would silently be solved with
    the first problem's)."""
    import torch

    from qpmpc_b200 import BatchedMPCProblem, ProblemDefinitionError, solve_mpc_batch
    from qpmpc_b200.workloads import to_batched, triple_integrator_batch

    prob = to_batched(triple_integrator_batch(3, seed=2))
    good = torch.empty((3, 16), dtype=torch.float64)
    assert solve_mpc_batch(prob, out=good).inputs.data_ptr() == good.data_ptr()
    for bad in (torch.empty((3, 16), dtype=torch.float32), torch.empty((4, 16), dtype=torch.float64),
                torch.empty((16, 3), dtype=torch.float64).t()):
        with pytest.raises(ProblemDefinitionError):
            solve_mpc_batch(prob, out=bad)
    a = golden_problem(load_golden("triple_integrator"))
    b = golden_problem(load_golden("triple_integrator"))
    b.stage_input_cost_weight = 1e-3
    with pytest.raises(ProblemDefinitionError):
        BatchedMPCProblem.from_problems([a, b])
    both = BatchedMPCProblem.from_problems([a, golden_problem(load_golden("triple_integrator"))])
    plan = solve_mpc_batch(both)
    assert torch.equal(plan.inputs[0], plan.inputs[1]) and int(plan.status.sum()) == 0


def _ragged_problem():
    A, B = np.array([[1.0, 0.1], [0.0, 1.0]]), np.array([[0.005], [0.1]])
    C = [np.array([[0.0, 1.0]]), np.array([[0.0, 1.0], [0.0, -1.0], [1.0, 0.0]]), None, np.array([[1.0, 0.0]])]
    e = [np.array([0.3]), np.array([0.3, 0.3, 2.0]), np.array([1.0, 1.0]), np.array([0.8])]
    D = [None, None, np.array([[1.0], [-1.0]]), None]
    return MPCProblem(A, B, C, D, e, 4, 1.0, 0.1, 1e-2, initial_state=np.array([0.0, 0.0]),
                      goal_state=np.array([1.0, 0.0]), target_states=np.zeros(8))


def single_path_matches_batch_path(monkeypatch):
    """``solve_mpc`` through the page-locked staging block + host entry (single.py) and through
    device tensors as a batch of one (QPMPC_B200_SINGLE_ZEROCOPY=0): same plan, multipliers,
    status and iteration count -- time-invariant, time-varying and ragged problems, both methods,
    both precisions, an infeasible problem, and a second solve that reuses the cached block."""
    import torch

    from qpmpc_b200 import solve_mpc

    problems = [golden_problem(load_golden(name)) for name in ("triple_integrator", "humanoid", "pendulum", "random_ltv_cd")]
    problems.append(_ragged_problem())
    infeasible = golden_problem(load_golden("triple_integrator"))
    infeasible.update_initial_state(np.array([0.0, 0.0, 5.0]))
    problems.append(infeasible)
    for problem in problems + problems[:2]:
        for kw in ({}, {"method": "pdip"}, {"dtype": torch.float32}):
            if kw.get("dtype") is torch.float32 and problem is infeasible:
                continue
            monkeypatch.setenv("QPMPC_B200_SINGLE_ZEROCOPY", "1")
            a = solve_mpc(problem, "b200", **kw)
            monkeypatch.setenv("QPMPC_B200_SINGLE_ZEROCOPY", "0")
            b = solve_mpc(problem, "b200", **kw)
            assert a.is_empty == b.is_empty
            assert a.qpsol.extras == b.qpsol.extras
            if not a.is_empty:
                assert np.array_equal(a.qpsol.x, b.qpsol.x)
                assert (a.qpsol.z is None) == (b.qpsol.z is None)
                if a.qpsol.z is not None:
                    assert np.array_equal(a.qpsol.z, b.qpsol.z)
    assert problems[-1] is infeasible and solve_mpc(infeasible, "b200").is_empty


def test_solve_mpc_single_path_matches_batch_path_on_the_emulator(emulated_engine, monkeypatch):
    single_path_matches_batch_path(monkeypatch)
    assert emulated_engine.calls > 0
