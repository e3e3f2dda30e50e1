"""GPU parity of the interior-point method (``method="pdip"``,
``qpmpc_b200/csrc/mpc_pdip.cuh``) through the C ABI.

The kernel is checked (a) against the NumPy statement of the same iteration
(``oracle/pdip_np.py``: iteration counts and iterates of the interior-point
phase, polish switched off) and (b) against the exact active-set oracle.  Bar,
fp64, with the primal-dual active-set polish (the default): ``|dU|_inf <= 1e-6``
on every workload, the same bar as the default method -- including the triple
integrator, where ``lambda_min(P) = w_u = 1e-6`` and the interior point alone
only determines the objective (checked too: within 1e-8 of the optimum, primal
feasible to 1e-8).  It stands in for the interior-point / augmented-Lagrangian
backends a ``solver=`` string selects at ``qpmpc/solve_mpc.py:43`` of the
reference.
"""

U_TOL = 1e-6  # |du|_inf, fp64 (BASELINE.json north_star)

import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(60)]


def _solve(w, dtype=None, **kw):
    import torch

    from qpmpc_b200 import solve_mpc_batch
    from qpmpc_b200.workloads import to_batched

    plan = solve_mpc_batch(to_batched(w, dtype=dtype), method="pdip", return_multipliers=True, **kw)
    torch.cuda.synchronize()
    B = w["batch"]
    return dict(U=plan.inputs.reshape(B, -1).double().cpu().numpy(), status=plan.status.cpu().numpy(),
                iters=plan.iters.cpu().numpy(), z=plan.multipliers.double().cpu().numpy())


def _oracle(w):
    import oracle
    from qpmpc_b200.workloads import oracle_ops

    return oracle.solve_batch(w["batch"], w["N"], w["nx"], w["nu"], w["nc"], oracle_ops(w),
                              w["w_t"], w["w_x"], w["w_u"])


def _condensed(w, count):
    import oracle
    from qpmpc_b200.workloads import oracle_ops

    ops = oracle_ops(w)
    out = []
    for b in range(count):
        def pick(name):
            arr, *flags = ops[name]
            return None if arr is None else (arr[b] if flags[0] else arr)
        c = oracle.condense(w["N"], w["nx"], w["nu"], w["nc"], pick("A"), pick("B"), pick("C"), pick("D"),
                            pick("e"), pick("x0"), pick("goal"), pick("targets"), w["w_t"], w["w_x"], w["w_u"])
        out.append((c["P"], c["q"], c["G"], c["h"]))
    return tuple(np.stack(a) for a in zip(*out))


def _objective(P, q, U):
    return 0.5 * np.einsum("bi,bij,bj->b", U, P, U) + np.einsum("bi,bi->b", q, U)


@pytest.mark.parametrize("N,batch", [(16, 2048), (8, 1030), (32, 257)])
def test_triple_integrator_objective_and_model(N, batch):
    """BASELINE config 2 / 5 shapes: every instance converges; on a subsample the
    kernel follows the NumPy model iteration for iteration and reaches the
    exact optimum's objective."""
    from oracle.pdip_np import pdip_batch
    from qpmpc_b200.workloads import triple_integrator_batch

    w = triple_integrator_batch(batch, N=N, seed=N)
    got = _solve(w, tol=1e-11, polish=False)
    assert (got["status"] == 0).all()
    assert got["iters"].max() <= 30
    k = 48
    P, q, G, h = _condensed(w, k)
    model = pdip_batch(P, q, G, h, tol=1e-11, polish=False, max_iter=50)
    assert (np.abs(got["iters"][:k] - model["iters"]) <= 1).all()
    same = got["iters"][:k] == model["iters"]
    assert same.mean() >= 0.9
    assert np.abs(got["U"][:k][same] - model["U"][same]).max() <= 1e-5
    ref = _oracle(w)
    assert np.abs(_objective(P, q, got["U"][:k]) - _objective(P, q, ref["U"][:k])).max() <= 1e-8
    assert (np.einsum("bmn,bn->bm", G, got["U"][:k]) - h).max() <= 1e-8
    assert got["z"].min() >= 0.0
    # with the polish (default) the result is the exact solution
    pol = _solve(w, tol=1e-9)
    assert (pol["status"] == 0).all()
    assert np.abs(pol["U"] - ref["U"]).max() <= U_TOL
    assert pol["z"].min() >= -1e-9 * max(1.0, pol["z"].max())


@pytest.mark.parametrize("kind", ["pendulum", "pendulum_ltv", "humanoid"])
def test_well_conditioned_workloads_match_the_exact_solution(kind):
    """Configs 3 and 4 data in fp64: |dU|_inf <= 1e-6 against the exact oracle."""
    from qpmpc_b200.workloads import humanoid_batch, pendulum_batch

    w = humanoid_batch(1024) if kind == "humanoid" else pendulum_batch(1024, ltv_model=kind.endswith("ltv"))
    got = _solve(w, tol=1e-10)
    ref = _oracle(w)
    assert (got["status"] == 0).all() and (ref["status"] == 0).all()
    assert np.abs(got["U"] - ref["U"]).max() <= U_TOL


@pytest.mark.parametrize("shape", [(4, 3, 2, 4), (7, 5, 1, 4), (5, 2, 2, 5), (6, 3, 2, 7), (9, 4, 3, 5), (3, 6, 2, 0)])
@pytest.mark.parametrize("ltv", [False, True])
def test_random_shapes(shape, ltv):
    """Every (NP, MR) variant, generic-nx condensing, C and D present, ragged batch."""
    from qpmpc_b200.workloads import random_batch

    N, nx, nu, nc = shape
    w = random_batch(131, N, nx, nu, nc, seed=sum(shape), ltv=ltv)
    got = _solve(w, tol=1e-10)
    ref = _oracle(w)
    ok = ref["status"] == 0
    assert ok.any()
    assert (got["status"][ok] == 0).all()
    assert np.abs(got["U"][ok] - ref["U"][ok]).max() <= U_TOL
    assert (got["status"][~ok] != 0).all() and np.isnan(got["U"][~ok]).all()


def test_host_buffer_entry_accepts_the_method():
    """qpmpc_b200_solve_host with desc.method = PDIP (NumPy buffers in and out)."""
    import ctypes

    from qpmpc_b200 import _capi
    from qpmpc_b200.workloads import to_batched, triple_integrator_batch

    w = triple_integrator_batch(1000, seed=9)
    desc = to_batched(w).desc(_capi.PDIP, 0, 1e-9)
    arrs = {k: np.ascontiguousarray(w[k]) for k in ("A", "B", "C", "e", "x0", "goal")}
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    ops = _capi.Operands(ptr(arrs["A"]), ptr(arrs["B"]), ptr(arrs["C"]), None, ptr(arrs["e"]),
                         ptr(arrs["x0"]), ptr(arrs["goal"]), None)
    U = np.zeros((1000, 16))
    st = np.zeros(1000, dtype=np.int32)
    outs = _capi.Outputs(ptr(U), ptr(st), None, None)
    rc = _capi.load().qpmpc_b200_solve_host(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs), 0)
    assert rc == 0
    assert (st == 0).all()
    assert np.abs(U - _oracle(w)["U"]).max() <= U_TOL


def test_unsupported_combinations_are_refused():
    """N = 64 (beyond the warp kernel) and single precision: the interior-point
    method says so (QPMPC_B200_EUNSUPPORTED) instead of falling back."""
    import torch

    from qpmpc_b200.exceptions import BackendError
    from qpmpc_b200.workloads import humanoid_batch, triple_integrator_batch

    with pytest.raises(BackendError):
        _solve(triple_integrator_batch(4, N=64))
    with pytest.raises(BackendError):
        _solve(humanoid_batch(4), dtype=torch.float32)
