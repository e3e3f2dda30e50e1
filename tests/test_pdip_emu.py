"""The interior-point phases of the CUDA kernel, run on the host.

``qpmpc_b200/csrc/mpc_pdip.cuh:pdip_core`` (the device source behind
``desc.method = QPMPC_B200_PDIP``) is compiled for the host against
``tests/emu/warp_emu.h`` -- one fiber per CUDA thread, every ``__syncwarp`` /
shuffle / vote a scheduling point -- and run on explicit QPs ``(P, q, G, h)``,
checked against the NumPy statement of the same iteration
(``oracle/pdip_np.py``) and the exact active-set oracle.  This is the CPU-side
evidence for the arithmetic, indexing and lock-step logic of the
interior-point phases; ``tests/test_kernel_emu.py`` runs the whole kernels
(staging, condensing, outputs) the same way and ``tests/test_gpu_pdip.py``
repeats the comparison on the device through the C ABI.  (It replaces the third-party solve at ``qpmpc/solve_mpc.py:43`` of the
reference, like every other solver test here.)
"""

import numpy as np
import pytest

import oracle
from emu import pdip_core
from oracle.pdip_np import pdip_batch
from qpmpc_b200.workloads import (humanoid_batch, oracle_ops, pendulum_batch, random_batch,
                                  triple_integrator_batch)


@pytest.fixture(scope="module")
def emu():
    """``emu(P, q, G, h, NP, MR, ...)``: one emulated warp; ``emu.all`` splits a batch into warps."""
    def solve_all(P, q, G, h, np_, mr, **kw):
        per = 32 // np_
        parts = [pdip_core(P[b:b + per], q[b:b + per], G[b:b + per], h[b:b + per], np_, mr, **kw)
                 for b in range(0, len(q), per)]
        return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}

    def solve(*a, **kw):
        return pdip_core(*a, **kw)

    solve.all = solve_all
    return solve


def condensed(w):
    """(P, q, G, h) stacks of a workload dict through the pinned C condensing."""
    ops = oracle_ops(w)
    out = []
    for b in range(w["batch"]):
        def pick(name):
            arr, *flags = ops[name]
            return None if arr is None else (arr[b] if flags[0] else arr)
        c = oracle.condense(w["N"], w["nx"], w["nu"], w["nc"], pick("A"), pick("B"), pick("C"), pick("D"),
                            pick("e"), pick("x0"), pick("goal"), pick("targets"), w["w_t"], w["w_x"], w["w_u"])
        out.append((c["P"], c["q"], c["G"], c["h"]))
    return tuple(np.stack(a) for a in zip(*out))


def exact(w):
    ref = oracle.solve_batch(w["batch"], w["N"], w["nx"], w["nu"], w["nc"], oracle_ops(w), w["w_t"], w["w_x"],
                             w["w_u"])
    assert ref["bad"] == 0
    return ref["U"]


def objective(P, q, U):
    return 0.5 * np.einsum("bi,bij,bj->b", U, P, U) + np.einsum("bi,bi->b", q, U)


def test_emulated_kernel_follows_the_numpy_model_iteration_for_iteration(emu):
    """Triple integrator N = 16 (BASELINE config 2): two instances per warp."""
    w = triple_integrator_batch(12, seed=4)
    P, q, G, h = condensed(w)
    got = emu.all(P, q, G, h, 16, 2, tol=1e-10, polish=False)
    model = pdip_batch(P, q, G, h, tol=1e-10, polish=False, max_iter=50)
    assert (got["status"] == 0).all() and (model["status"] == 0).all()
    assert (got["iters"] == model["iters"]).all()
    assert np.abs(got["U"] - model["U"]).max() <= 1e-6
    # against the exact optimum: same objective, primal feasible, multipliers >= 0
    Uref = exact(w)
    assert np.abs(objective(P, q, got["U"]) - objective(P, q, Uref)).max() <= 1e-8
    assert (np.einsum("bmn,bn->bm", G, got["U"]) - h).max() <= 1e-8
    assert got["z"].min() >= 0.0
    # the interior point alone does not pin U at w_u = 1e-6 ...
    assert np.abs(got["U"] - Uref).max() > 1e-4
    # ... the primal-dual active-set polish does: the north star's |dU| <= 1e-6
    pol = emu.all(P, q, G, h, 16, 2, tol=1e-9)
    assert (pol["status"] == 0).all()
    assert np.abs(pol["U"] - Uref).max() <= 1e-6
    model = pdip_batch(P, q, G, h, tol=1e-9, max_iter=50)
    assert model["polished"].all() and np.abs(pol["U"] - model["U"]).max() <= 1e-6


def test_well_conditioned_problems_match_the_exact_solution(emu):
    """Wheeled inverted pendulum (config 3 shape, w_u = 1e-3): |dU| <= 1e-6 after the polish."""
    w = pendulum_batch(8, seed=1)
    P, q, G, h = condensed(w)
    got = emu.all(P, q, G, h, 16, 2, tol=1e-10)
    assert (got["status"] == 0).all()
    assert np.abs(got["U"] - exact(w)).max() <= 1e-6


@pytest.mark.parametrize("N,np_,count", [(8, 8, 3), (8, 8, 4), (32, 32, 1)])
def test_lane_group_widths(emu, N, np_, count):
    """NP = 8 (four instances per warp, one tail slot empty) and NP = 32."""
    w = triple_integrator_batch(count, N=N, seed=3)
    P, q, G, h = condensed(w)
    got = emu(P, q, G, h, np_, 2, tol=1e-10, polish=False)
    model = pdip_batch(P, q, G, h, tol=1e-10, polish=False, max_iter=50)
    assert (got["status"] == 0).all()
    assert (np.abs(got["iters"] - model["iters"]) <= 1).all()
    # stopped at mu <= tol (1 + |obj|): the duality gap, hence the objective error, is <= m mu
    best = objective(P, q, exact(w))
    assert (np.abs(objective(P, q, got["U"]) - best) <= h.shape[1] * 1e-10 * (1.0 + np.abs(best))).all()


@pytest.mark.parametrize("N,nx,nu,nc,np_,mr", [(4, 3, 2, 4, 8, 2), (7, 5, 1, 4, 8, 4), (5, 2, 2, 5, 16, 2),
                                                 (6, 3, 2, 7, 16, 4), (9, 4, 3, 5, 32, 2)])
def test_random_ltv_shapes(emu, N, nx, nu, nc, np_, mr):
    """Per-step C_k, D_k, e_k, ragged against the lane-group width (n < NP, m < MR NP)."""
    w = random_batch(32 // np_, N, nx, nu, nc, seed=N)
    P, q, G, h = condensed(w)
    got = emu(P, q, G, h, np_, mr, tol=1e-10)
    ref = oracle.solve_batch(w["batch"], N, nx, nu, nc, oracle_ops(w), w["w_t"], w["w_x"], w["w_u"])
    ok = ref["status"] == 0
    assert ok.any()
    assert (got["status"][ok] == 0).all()
    assert np.abs(got["U"][ok] - ref["U"][ok]).max() <= 1e-6
    assert (got["status"][~ok] != 0).all()  # infeasible instances are never reported solved


def test_lock_step_neighbours_do_not_interact(emu):
    """An instance gives bit-identical results alone and next to a harder one."""
    w = triple_integrator_batch(6, seed=9)
    P, q, G, h = condensed(w)
    alone = [emu(P[b:b + 1], q[b:b + 1], G[b:b + 1], h[b:b + 1], 16, 2) for b in range(6)]
    order = np.argsort([a["iters"][0] for a in alone])
    lo, hi = order[0], order[-1]
    assert alone[lo]["iters"][0] < alone[hi]["iters"][0]
    pair = emu(P[[lo, hi]], q[[lo, hi]], G[[lo, hi]], h[[lo, hi]], 16, 2)
    assert np.array_equal(pair["U"][0], alone[lo]["U"][0]) and np.array_equal(pair["U"][1], alone[hi]["U"][0])
    assert pair["iters"].tolist() == [alone[lo]["iters"][0], alone[hi]["iters"][0]]


def test_infeasible_and_unconstrained(emu):
    n = 6
    P = np.eye(n)[None] * 2.0
    q = np.arange(1.0, n + 1)[None]
    # x_0 <= -1 and -x_0 <= -1 cannot both hold
    G = np.zeros((1, 2, n)); G[0, 0, 0] = 1.0; G[0, 1, 0] = -1.0
    h = np.array([[-1.0, -1.0]])
    got = emu(P, q, G, h, 8, 2, max_iter=30)
    # the iterates diverge: either the cap is hit or H loses positive definiteness on the way
    assert got["status"][0] in (1, 3) and got["iters"][0] <= 30
    # no constraint rows at all: one Newton step to -P^-1 q
    got = emu(P, q, np.zeros((1, 0, n)), np.zeros((1, 0)), 8, 2)
    assert got["status"][0] == 0 and got["iters"][0] == 1
    assert np.abs(got["U"][0] + q[0] / 2.0).max() <= 1e-14


def test_single_precision(emu):
    """The float instantiation of the core on the humanoid data (config 4 shape).
    Not offered through the ABI (qpmpc_b200.cu:solve_impl refuses it): at larger
    sample sizes 7-15 % of the instances stall above tol = 1e-6."""
    w = humanoid_batch(4, seed=2)
    P, q, G, h = condensed(w)
    got = emu.all(P, q, G, h, 16, 2, dtype=1, tol=1e-6)
    assert (got["status"] == 0).all()
    Uref = exact(w)
    assert np.abs(got["U"] - Uref).max() <= 1e-3 * max(1.0, np.abs(Uref).max())


@pytest.mark.parametrize("bound", [1.0, 1e20, 1e30])
def test_padding_rows_and_no_bound_constants_do_not_relax_the_tests(emu, bound):
    """Ragged per-step constraints are padded with all-zero rows (``pack_problem``), and callers
    write "no bound" as a huge constant: with a batch-wide scale max|h| such a row made every
    primal test pass (a violated point was reported solved).  The tests are per row now."""
    rng = np.random.default_rng(3)
    N, nx, nu, nc = 8, 3, 1, 3
    ncs = [2, 1, 3, 2, 1, 2, 3, 1]
    w = random_batch(2, N, nx, nu, nc, seed=3, w_x=None)
    for k in range(N):
        w["C"][:, k, ncs[k]:] = 0.0
        w["D"][:, k, ncs[k]:] = 0.0
        w["e"][:, k, ncs[k]:] = bound
    w["x0"] = 0.6 * rng.standard_normal((2, nx))
    P, q, G, h = condensed(w)
    ref = [oracle.qp_gi(P[b], q[b], G[b], h[b]) for b in range(2)]
    got = emu(P, q, G, h, 8, 4, tol=1e-9, max_iter=50)
    for b in range(2):
        if ref[b][0] == 0:
            assert got["status"][b] == 0
            assert np.abs(got["U"][b] - ref[b][1]).max() <= 1e-6
            assert (G[b] @ got["U"][b] - h[b]).max() <= 1e-8
        else:
            assert got["status"][b] != 0
    model = pdip_batch(P, q, G, h, tol=1e-9, max_iter=50)
    assert (model["status"] == got["status"]).all()
