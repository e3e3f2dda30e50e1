"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the
golden fixtures generated from the reference's own condensing.

Bars: condensing within 1e-12 relative of the reference fields (summation
order differs from ndarray.dot); U within 1e-6 absolute of the exact
active-set oracle in fp64 (BASELINE north_star), plus KKT certificates.
"""

import numpy as np
import pytest

from conftest import GOLDEN_NAMES, golden_problem, load_golden

pytestmark = pytest.mark.gpu

U_TOL = 1e-6  # |du|_inf, fp64 (BASELINE.json north_star)


def _solve(w, **kw):
    import torch

    from qpmpc_b200 import solve_mpc_batch
    from qpmpc_b200.workloads import to_batched

    prob = to_batched(w)
    plan = solve_mpc_batch(prob, return_multipliers=True, **kw)
    torch.cuda.synchronize()
    return prob, plan


def _oracle(w):
    import oracle
    from qpmpc_b200.workloads import oracle_ops

    return oracle.solve_batch(w["batch"], w["N"], w["nx"], w["nu"], w["nc"], oracle_ops(w),
                              w["w_t"], w["w_x"], w["w_u"], want_kkt=True)


def _check_against_oracle(w, tol=U_TOL):
    prob, plan = _solve(w)
    ref = _oracle(w)
    U = plan.inputs.reshape(w["batch"], -1).cpu().numpy()
    st = plan.status.cpu().numpy()
    assert np.array_equal(st == 0, ref["status"] == 0), (st[:16], ref["status"][:16])
    ok = st == 0
    assert ok.any()
    err = np.abs(U[ok] - ref["U"][ok]).max()
    assert err <= tol, f"|dU|_inf = {err:.3e}"
    assert np.isnan(U[~ok]).all()
    return prob, plan, ref


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_condense_matches_reference_fields(name):
    """mpc_condense_kernel reproduces the reference MPCQP fields (golden)."""
    g = load_golden(name)
    from qpmpc_b200 import MPCQP

    qp = MPCQP(golden_problem(g))
    for field in ("P", "q", "G", "h", "Phi", "Psi", "phi_last", "psi_last", "e"):
        ref = g[f"ref_{field}"]
        got = getattr(qp, field)
        assert got.shape == ref.shape, field
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(got - ref).max() <= 1e-12 * scale, field


@pytest.mark.parametrize("name", ["triple_integrator", "humanoid", "pendulum",
                                  "random_ltv_cd", "random_ltv_c", "random_ltv_d",
                                  "triple_integrator_N8", "triple_integrator_N32",
                                  "triple_integrator_N64", "triple_integrator_stage",
                                  "triple_integrator_tiny_wt"])
def test_solve_mpc_single_matches_oracle(name):
    """solve_mpc(problem, "b200") on the golden problems vs the exact QP oracle
    applied to the REFERENCE's condensed matrices."""
    import oracle
    from qpmpc_b200 import solve_mpc

    g = load_golden(name)
    st, x, z, _ = oracle.qp_gi(g["ref_P"], g["ref_q"], g["ref_G"], g["ref_h"])
    plan = solve_mpc(golden_problem(g), solver="b200")
    if st != 0:
        assert plan.is_empty
        return
    assert not plan.is_empty
    assert np.abs(plan.inputs.reshape(-1) - x).max() <= U_TOL
    kkt = oracle.kkt(g["ref_P"], g["ref_q"], g["ref_G"], g["ref_h"],
                     plan.inputs.reshape(-1), plan.qpsol.z)
    scale = max(1.0, np.abs(g["ref_q"]).max())
    assert kkt[0] <= 1e-9 * scale and kkt[1] <= 1e-9 and kkt[2] == 0.0 and kkt[3] <= 1e-9


def test_triple_integrator_known_answer():
    """The one non-trivial QP answer recorded in SURVEY.md 8(c) (KKT-certified)."""
    from qpmpc_b200 import solve_mpc

    g = load_golden("triple_integrator")
    plan = solve_mpc(golden_problem(g), solver="b200")
    U = plan.inputs.reshape(-1)
    expect = {0: 48.0, 7: -27.51976087031062, 8: -54.819558402499155,
              9: -13.660680727190227, 15: 47.92404503861714}
    for i in range(16):
        assert abs(U[i] - expect.get(i, 0.0)) <= U_TOL


@pytest.mark.parametrize("N,batch", [(16, 4096), (8, 2048), (32, 1024), (64, 512)])
def test_triple_integrator_batch(N, batch):
    """BASELINE config 2 / 5 shapes against the oracle."""
    from qpmpc_b200.workloads import triple_integrator_batch

    _check_against_oracle(triple_integrator_batch(batch, N=N, seed=N))


def test_triple_integrator_shared_model_and_jitter():
    from qpmpc_b200.workloads import triple_integrator_batch

    _check_against_oracle(triple_integrator_batch(1024, per_instance_model=False))
    _check_against_oracle(triple_integrator_batch(1024, jitter=0.1, seed=5))


@pytest.mark.parametrize("ltv_model", [False, True])
def test_pendulum_batch(ltv_model):
    """BASELINE config 3, one cycle: box bounds on u through D only."""
    from qpmpc_b200.workloads import pendulum_batch

    _check_against_oracle(pendulum_batch(2048, ltv_model=ltv_model))


def test_humanoid_batch():
    """BASELINE config 4 data in fp64: per-instance per-step e_k."""
    from qpmpc_b200.workloads import humanoid_batch

    _check_against_oracle(humanoid_batch(2048))


@pytest.mark.parametrize("shape", [(6, 3, 2, 3), (5, 4, 1, 2), (7, 2, 2, 4), (4, 5, 3, 6),
                                   (10, 2, 1, 1), (3, 6, 2, 0)])
@pytest.mark.parametrize("ltv", [False, True])
def test_random_shapes(shape, ltv):
    """Random per-instance problems over (N, nx, nu, nc), C and D both present."""
    from qpmpc_b200.workloads import random_batch

    N, nx, nu, nc = shape
    _check_against_oracle(random_batch(257, N, nx, nu, nc, seed=sum(shape), ltv=ltv))


@pytest.mark.parametrize("with_C,with_D,w_t,w_x", [(True, False, 0.7, None), (False, True, None, 0.3),
                                                   (True, True, 1e-12, 0.5)])
def test_random_operand_patterns(with_C, with_D, w_t, w_x):
    from qpmpc_b200.workloads import random_batch

    _check_against_oracle(random_batch(130, 6, 3, 2, 2, seed=7, with_C=with_C, with_D=with_D,
                                       w_t=w_t, w_x=w_x))


@pytest.mark.parametrize("kind", ["ti16", "ti8", "pendulum", "pendulum_ltv", "humanoid", "random",
                                  "random_lti", "infeasible"])
def test_cta_kernel_matches_oracle(kind, monkeypatch):
    """The CTA-per-instance kernel (n > 32 path) forced onto small shapes."""
    from qpmpc_b200.workloads import (humanoid_batch, pendulum_batch, random_batch,
                                      triple_integrator_batch)

    monkeypatch.setenv("QPMPC_B200_FORCE_CTA", "1")
    if kind == "ti16":
        w = triple_integrator_batch(700, N=16, seed=21)
    elif kind == "ti8":
        w = triple_integrator_batch(300, N=8, seed=22)
    elif kind == "pendulum":
        w = pendulum_batch(400)
    elif kind == "pendulum_ltv":
        w = pendulum_batch(200, ltv_model=True)
    elif kind == "humanoid":
        w = humanoid_batch(400)
    elif kind == "random":
        w = random_batch(200, 7, 5, 2, 3, seed=23, ltv=True)
    elif kind == "random_lti":
        w = random_batch(200, 9, 2, 3, 5, seed=24, ltv=False)
    else:
        w = triple_integrator_batch(64, seed=11)
        w["x0"][::4, 2] = 5.0
    _check_against_oracle(w)


@pytest.mark.parametrize("kind", ["forced_ti16", "forced_pendulum", "forced_random", "humanoid_N100", "ti_stage_N96",
                                  "random_n80_nu2", "ti_N128", "ti_N256"])
def test_cta_kernel_beyond_shared_memory(kind, monkeypatch):
    """Horizons whose matrices do not fit in 227 KB (n > 72 with m = 2 n in fp64): the CTA kernel
    with its matrices in the stream-ordered global-memory workspace (qpmpc_b200.cu: plan_cta), a
    bounded grid striding over the batch -- forced onto small shapes with more instances than
    CTAs, and on horizons that need it, up to N = 256 (n = 256, m = 512)."""
    import time

    import torch

    from qpmpc_b200 import solve_mpc_batch
    from qpmpc_b200.workloads import (humanoid_batch, pendulum_batch, random_batch, to_batched,
                                      triple_integrator_batch)

    if kind.startswith("forced"):
        monkeypatch.setenv("QPMPC_B200_FORCE_CTA", "1")
        monkeypatch.setenv("QPMPC_B200_CTA_WORKSPACE", "1")
    w = {"forced_ti16": lambda: triple_integrator_batch(3000, N=16, seed=61),
         "forced_pendulum": lambda: pendulum_batch(2000, seed=62),
         "forced_random": lambda: random_batch(1500, 7, 5, 2, 3, seed=63, ltv=True),
         "humanoid_N100": lambda: humanoid_batch(600, N=100, seed=64),
         "ti_stage_N96": lambda: triple_integrator_batch(300, N=96, seed=65),
         "random_n80_nu2": lambda: random_batch(300, 40, 4, 2, 3, seed=66, ltv=True, w_u=1.0),
         "ti_N128": lambda: triple_integrator_batch(64, N=128, seed=67),
         "ti_N256": lambda: triple_integrator_batch(8, N=256, seed=68)}[kind]()
    if kind == "ti_stage_N96":
        w["w_x"], w["targets"] = 0.5, np.zeros((w["batch"], 96 * 3))
    prob = to_batched(w)
    t0 = time.perf_counter()
    plan = solve_mpc_batch(prob)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    import oracle
    from qpmpc_b200.workloads import oracle_ops

    ref = oracle.solve_batch(w["batch"], w["N"], w["nx"], w["nu"], w["nc"], oracle_ops(w), w["w_t"], w["w_x"], w["w_u"])
    st = plan.status.cpu().numpy()
    U = plan.inputs.reshape(w["batch"], -1).cpu().numpy()
    assert np.array_equal(st == 0, ref["status"] == 0), (st[:16], ref["status"][:16])
    ok = st == 0
    assert ok.mean() > 0.5
    assert np.array_equal(plan.iters.cpu().numpy()[ok], ref["iters"][ok])
    err = np.abs(U[ok] - ref["U"][ok]).max()
    assert err <= U_TOL, f"|dU|_inf = {err:.3e}"
    print(f"{kind}: {w['batch']} instances in {dt * 1e3:.1f} ms, |dU| {err:.2e}")
    if kind == "ti_N128":  # the MPCQP fields of a horizon this long, against the oracle's condensing
        from qpmpc_b200 import condense_batch

        fields = condense_batch(prob, ("P", "q", "G", "h"))
        ops = oracle_ops(w)
        pick = lambda a, flag: None if a is None else (a[3] if flag else a)  # noqa: E731
        c = oracle.condense(w["N"], w["nx"], w["nu"], w["nc"], *[pick(*ops[k][:2]) for k in ("A", "B", "C", "D", "e")],
                            w["x0"][3], w["goal"][3], None, w["w_t"], w["w_x"], w["w_u"])
        for field in ("P", "q", "G", "h"):
            ref_f = np.asarray(c[field])
            got = fields[field][3].cpu().numpy().reshape(ref_f.shape)
            assert np.abs(got - ref_f).max() <= 1e-12 * max(1.0, np.abs(ref_f).max()), field


def test_cta_kernel_condense_fields(monkeypatch):
    """mpc_condense_cta_kernel reproduces the reference MPCQP fields."""
    from qpmpc_b200 import MPCQP

    monkeypatch.setenv("QPMPC_B200_FORCE_CTA", "1")
    for name in ("humanoid", "pendulum", "random_ltv_cd", "triple_integrator_stage"):
        g = load_golden(name)
        qp = MPCQP(golden_problem(g))
        for field in ("P", "q", "G", "h", "Phi", "Psi", "phi_last", "psi_last"):
            ref = g[f"ref_{field}"]
            got = getattr(qp, field)
            scale = max(1.0, np.abs(ref).max())
            assert np.abs(got - ref).max() <= 1e-12 * scale, (name, field)


def test_ragged_tail_and_tiny_batches():
    """Batches that do not fill a CTA / warp (bulk-copy fallback path)."""
    from qpmpc_b200.workloads import triple_integrator_batch

    for batch in (1, 2, 3, 5, 7, 31):
        _check_against_oracle(triple_integrator_batch(batch, seed=batch))


def test_infeasible_instances_are_flagged():
    """Q3 / H4: an infeasible instance gets status 2 and NaN inputs; its warp
    neighbour is unaffected."""
    from qpmpc_b200.workloads import triple_integrator_batch

    w = triple_integrator_batch(64, seed=11)
    w["x0"][::4, 2] = 5.0  # |accel_0| > 3: row k=0 violated, inputs cannot repair it
    prob, plan, ref = _check_against_oracle(w)
    st = plan.status.cpu().numpy()
    assert (st[::4] == 2).all() and (st[1::4] == 0).all()


def test_multipliers_certify_kkt():
    """Z >= 0 and KKT residuals of (U, Z) on the oracle's condensed matrices."""
    import oracle
    from qpmpc_b200.workloads import triple_integrator_batch

    w = triple_integrator_batch(64, seed=3)
    prob, plan = _solve(w)
    U = plan.inputs.reshape(64, -1).cpu().numpy()
    Z = plan.multipliers.cpu().numpy()
    assert (Z >= 0).all()
    for b in range(64):
        c = oracle.condense(16, 3, 1, 2, w["A"][b], w["B"][b], w["C"][b], None, w["e"][b],
                            w["x0"][b], w["goal"][b], None, 1.0, None, 1e-6)
        k = oracle.kkt(c["P"], c["q"], c["G"], c["h"], U[b], Z[b])
        assert k[0] <= 1e-10 and k[1] <= 1e-9 and k[3] <= 1e-9, k


def test_states_match_host_integrate():
    from qpmpc_b200.workloads import pendulum_batch

    w = pendulum_batch(33)
    prob, plan = _solve(w)
    X = plan.states.cpu().numpy()
    U = plan.inputs.cpu().numpy()
    for b in (0, 17, 32):
        x = w["x0"][b].copy()
        for k in range(w["N"]):
            assert np.abs(X[b, k] - x).max() <= 1e-12
            x = w["A"] @ x + w["B"] @ U[b, k]
        assert np.abs(X[b, -1] - x).max() <= 1e-12


def test_solve_host_entry_matches_device_entry():
    """qpmpc_b200_solve_host (host buffers, the e2e call) == device entry."""
    import ctypes

    from qpmpc_b200 import _capi
    from qpmpc_b200.workloads import triple_integrator_batch

    w = triple_integrator_batch(1000, seed=9)
    prob, plan = _solve(w)
    lib = _capi.load()
    desc = prob.desc()
    arrs = {k: np.ascontiguousarray(w[k]) for k in ("A", "B", "C", "e", "x0", "goal")}
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    ops = _capi.Operands(ptr(arrs["A"]), ptr(arrs["B"]), ptr(arrs["C"]), None, ptr(arrs["e"]),
                         ptr(arrs["x0"]), ptr(arrs["goal"]), None)
    U = np.zeros((1000, 16))
    st = np.zeros(1000, dtype=np.int32)
    outs = _capi.Outputs(ptr(U), ptr(st), None, None)
    rc = lib.qpmpc_b200_solve_host(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs), 0)
    assert rc == 0
    assert np.array_equal(U, plan.inputs.reshape(1000, 16).cpu().numpy())
    assert (st == 0).all()


def test_fp32_humanoid_within_stated_tolerance():
    """BASELINE config 4 runs in fp32: |du|_inf <= 1e-4 * max(1, |u|_inf) vs the
    fp64 oracle (measured 1e-5; the reference itself is float64 only,
    mpc_qp.py:93-98)."""
    import torch

    from qpmpc_b200 import solve_mpc_batch
    from qpmpc_b200.workloads import humanoid_batch, to_batched

    w = humanoid_batch(1024)
    ref = _oracle(w)
    plan = solve_mpc_batch(to_batched(w, dtype=torch.float32))
    U = plan.inputs.reshape(1024, -1).double().cpu().numpy()
    st = plan.status.cpu().numpy()
    ok = (st == 0) & (ref["status"] == 0)
    assert ok.mean() > 0.99
    scale = np.maximum(1.0, np.abs(ref["U"][ok]).max(axis=1))
    err = np.abs(U[ok] - ref["U"][ok]).max(axis=1) / scale
    assert err.max() <= 1e-4, err.max()


def test_fp32_shared_model_path_and_walking_loop():
    """Single precision through the shared-model path: the humanoid pattern (BASELINE config 4)
    with the model factored once holds the fp32 bar against the fp64 oracle, and the walking
    loop of the same model (the closed-loop form of config 4), all 300 cycles in one fp32 launch,
    stays within 2e-3 of the fp64 CPU loop with every cycle solved."""
    import torch

    from qpmpc_b200 import factor_model, lipm_walking_closed_loop, solve_mpc_batch
    from qpmpc_b200.workloads import (humanoid_batch, lipm_advance, lipm_phase_vectors, lipm_walking_batch,
                                      to_batched)

    w = humanoid_batch(1024)
    ref = _oracle(w)
    prob = to_batched(w, dtype=torch.float32)
    plan = solve_mpc_batch(prob, factored=factor_model(prob))
    U = plan.inputs.reshape(1024, -1).double().cpu().numpy()
    ok = (plan.status.cpu().numpy() == 0) & (ref["status"] == 0)
    assert ok.mean() > 0.99
    scale = np.maximum(1.0, np.abs(ref["U"][ok]).max(axis=1))
    err = (np.abs(U[ok] - ref["U"][ok]).max(axis=1) / scale).max()
    assert err <= 1e-4, err

    import oracle
    from qpmpc_b200.workloads import oracle_ops

    B, cycles = 64, 300
    w = lipm_walking_batch(B, seed=4)
    x, foot = w["x0"].copy(), w["support_foot"].copy()
    pidx, sidx = w["phase_index"].copy(), w["stride_index"].copy()
    traj = [x.copy()]
    for _ in range(cycles):
        e, goal = lipm_phase_vectors(w, foot, pidx, sidx)
        wc = dict(w, e=e, goal=goal, x0=x)
        r = oracle.solve_batch(B, w["N"], w["nx"], w["nu"], w["nc"], oracle_ops(wc), w["w_t"], w["w_x"], w["w_u"])
        assert (r["status"] == 0).all()
        x, foot, pidx, sidx = lipm_advance(w, x, r["U"][:, 0], foot, pidx, sidx)
        traj.append(x.copy())
    ref_traj = np.stack(traj)
    prob = to_batched(w, dtype=torch.float32)
    plan, got, unsolved, phase = lipm_walking_closed_loop(prob, w["support_foot"], w["strides"], w["phase_index"],
                                                          w["stride_index"], cycles, record=True,
                                                          factored=factor_model(prob))
    torch.cuda.synchronize()
    assert int(unsolved.item()) == 0
    assert np.abs(got.double().cpu().numpy() - ref_traj).max() <= 2e-3
    assert np.array_equal(phase["phase_index"].cpu().numpy(), pidx)


def _cpu_closed_loop(w, cycles, substeps=15, return_iters=False):
    """The loop of examples/wheeled_inverted_pendulum.py:99-118 on the CPU: the
    oracle solves, the host mirror of the plant integrates."""
    import oracle
    from qpmpc_b200.systems import WheeledInvertedPendulum
    from qpmpc_b200.workloads import oracle_ops, pendulum_targets

    pend = WheeledInvertedPendulum()
    B, N, T = w["batch"], w["N"], w["T"]
    state = w["x0"].copy()
    traj, iters = [state.copy()], []
    w = dict(w)
    for _ in range(cycles):
        w["x0"] = state
        w["targets"], w["goal"] = pendulum_targets(state, w["v_target"], N, T)
        ref = oracle.solve_batch(B, N, 4, 1, 2, oracle_ops(w), w["w_t"], w["w_x"], w["w_u"])
        assert (ref["status"] == 0).all()
        if B <= 64:
            new = np.empty_like(state)
            for b in range(B):
                x = state[b]
                for _s in range(substeps):
                    x = pend.integrate(x, ref["U"][b, 0], T / substeps)
                new[b] = x
            state = new
        else:
            # the same Taylor step (systems/wheeled_inverted_pendulum.py:127-160), vectorised
            u, dt, x = ref["U"][:, 0], T / substeps, state
            for _s in range(substeps):
                pa = pend.omega**2 * (np.sin(x[:, 1]) - (u / pend.GRAVITY) * np.cos(x[:, 1]))
                x = np.stack([x[:, 0] + dt * x[:, 2] + 0.5 * dt * dt * u, x[:, 1] + dt * x[:, 3] + 0.5 * dt * dt * pa,
                              x[:, 2] + dt * u, x[:, 3] + dt * pa], axis=1)
            state = x
        iters.append(ref["iters"].copy())
        traj.append(state.copy())
    if return_iters:
        return np.stack(traj), np.stack(iters)
    return np.stack(traj)


def test_pendulum_closed_loop_matches_cpu_loop():
    """BASELINE config 3 (receding horizon): device loop == CPU loop, 40 cycles."""
    import torch

    from qpmpc_b200 import pendulum_closed_loop
    from qpmpc_b200.workloads import pendulum_batch, to_batched

    w = pendulum_batch(48, seed=1)
    ref = _cpu_closed_loop(w, 40)
    prob = to_batched(w)
    plan, traj, unsolved = pendulum_closed_loop(prob, w["v_target"], 40, record=True, factored=False)
    torch.cuda.synchronize()
    assert int(unsolved.item()) == 0
    got = traj.cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-6, np.abs(got - ref).max()
    # the bound |u| <= 10 is active early on for some instances (non-degenerate benchmark)
    assert np.abs(prob.x0.cpu().numpy() - ref[-1]).max() <= 1e-6


def test_pendulum_closed_loop_200_cycles_stays_upright():
    """Full-length loop: every instance has a plan in every cycle; the instances
    the bounded input can recover (most of this distribution) stay upright and
    their ground velocity converges to its target (size-independent property)."""
    import torch

    from qpmpc_b200 import pendulum_closed_loop
    from qpmpc_b200.workloads import pendulum_batch, to_batched

    w = pendulum_batch(2048, seed=1)
    prob = to_batched(w)
    plan, traj, unsolved = pendulum_closed_loop(prob, w["v_target"], 200, record=True)
    torch.cuda.synchronize()
    assert int(unsolved.item()) == 0
    X = traj.cpu().numpy()
    assert np.isfinite(X).all()
    upright = np.abs(X[:, :, 1]).max(axis=0) < 1.2
    assert upright.mean() > 0.9, upright.mean()
    assert np.abs(X[-1, upright, 2] - w["v_target"][upright]).max() < 5e-2


def test_solve_scatter_entry_writes_every_destination():
    """qpmpc_b200_solve_scatter (fused gather): rows land at row_offset in all
    destination buffers and equal the plain entry's output (single GPU: the
    'peers' are two local buffers)."""
    import ctypes

    import torch

    from qpmpc_b200 import _capi
    from qpmpc_b200.workloads import to_batched, triple_integrator_batch

    w = triple_integrator_batch(333, seed=4)
    prob, plan = _solve(w)
    ref = plan.inputs.reshape(333, 16)
    lib = _capi.load()
    desc = prob.desc()
    bufs = [torch.full((1000, 16), -7.0, dtype=torch.float64, device="cuda") for _ in range(2)]
    sts = [torch.full((1000,), -7, dtype=torch.int32, device="cuda") for _ in range(2)]
    peers = _capi.Peers()
    peers.count, peers.row_offset = 2, 500
    for r in range(2):
        peers.U[r] = bufs[r].data_ptr()
        peers.status[r] = sts[r].data_ptr()
    outs = _capi.Outputs(None, None, None, None)
    ops = prob.operands()
    rc = lib.qpmpc_b200_solve_scatter(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs),
                                      ctypes.byref(peers), None)
    assert rc == 0
    torch.cuda.synchronize()
    for r in range(2):
        assert torch.equal(bufs[r][500:833], ref)
        assert (bufs[r][:500] == -7).all() and (bufs[r][833:] == -7).all()
        assert (sts[r][500:833] == 0).all() and (sts[r][:500] == -7).all()


def test_non_finite_operands_are_flagged_not_returned():
    """NaN / inf in an instance's data gives status 3 and NaN inputs for that
    instance only; its warp neighbours are solved as usual."""
    from qpmpc_b200.workloads import triple_integrator_batch

    for force_cta in (False, True):
        w = triple_integrator_batch(40, seed=13)
        w["x0"][3, 1] = np.nan
        w["goal"][8, 0] = np.inf
        w["A"][21, 0, 1] = np.nan
        import os

        if force_cta:
            os.environ["QPMPC_B200_FORCE_CTA"] = "1"
        try:
            prob, plan = _solve(w)
        finally:
            os.environ.pop("QPMPC_B200_FORCE_CTA", None)
        st = plan.status.cpu().numpy()
        U = plan.inputs.reshape(40, -1).cpu().numpy()
        bad = np.array([3, 8, 21])
        assert (st[bad] != 0).all(), st[bad]
        assert np.isnan(U[bad]).all()
        good = np.setdiff1d(np.arange(40), bad)
        assert (st[good] == 0).all() and np.isfinite(U[good]).all()


# -- published known-answer QPs ---------------------------------------------------------------

@pytest.mark.parametrize("method", ["active_set", "pdip"])
def test_published_optima_on_the_device(method):
    """The solver half pinned to answers published by others: quadprog's documented example,
    the qpsolvers README example and Hock-Schittkowski 21 / 35 / 76 / 118 / 268 (citations in
    tests/published_qps.py), written as one-step MPC problems and solved through the C ABI by
    both kernels -- x* to the printed digits, the printed multipliers, and the oracle on the side."""
    import published_qps

    for make in published_qps.ALL:
        qp = make()
        w = published_qps.as_one_step_mpc(qp)
        _, plan = _solve(w, method=method, tol=1e-9)
        assert int(plan.status[0]) == 0, qp["name"]
        x = plan.inputs.reshape(-1).cpu().numpy()
        bar = max(qp["x_tol"], 1e-7) * max(1.0, np.abs(qp["x"]).max())
        assert np.abs(x - qp["x"]).max() <= bar, (qp["name"], x, qp["x"])
        if qp["z"] is not None:
            assert np.abs(plan.multipliers[0].cpu().numpy() - qp["z"]).max() <= 1e-6
        ref = _oracle(w)
        assert ref["status"][0] == 0 and np.abs(x - ref["U"][0]).max() <= U_TOL


@pytest.mark.parametrize("path", ["zero_copy", "staged_graph", "staged_streams", "pageable"])
@pytest.mark.parametrize("kind", ["ti", "pendulum"])
def test_solve_host_paths_are_bit_identical(path, kind, monkeypatch):
    """qpmpc_b200_solve_host with page-locked buffers (zero-copy: the kernel's bulk-TMA staging
    reads the host buffers itself; or staged in chunks, as a replayed CUDA graph or on streams)
    and with pageable ones returns exactly the rows of the device entry -- U, status, iterations,
    multipliers -- also when called again (graph replay) and with operands shared by the batch."""
    import ctypes

    import torch

    from qpmpc_b200 import _capi
    from qpmpc_b200.workloads import pendulum_batch, triple_integrator_batch

    B = 20000  # more than one chunk
    w = triple_integrator_batch(B, seed=19) if kind == "ti" else pendulum_batch(B, seed=20)
    prob, plan = _solve(w)
    if path != "zero_copy":
        monkeypatch.setenv("QPMPC_B200_HOST_ZEROCOPY", "0")
    if path == "staged_streams":
        monkeypatch.setenv("QPMPC_B200_HOST_GRAPH", "0")
    monkeypatch.setenv("QPMPC_B200_HOST_CHUNK", "4096")
    names = [k for k in ("A", "B", "C", "D", "e", "x0", "goal", "targets") if w[k] is not None]

    def buf(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t if path == "pageable" else t.pin_memory()

    host = {k: buf(w[k]) for k in names}
    n, m = w["N"] * w["nu"], w["N"] * w["nc"]
    U, Z = buf(np.zeros((B, n))), buf(np.zeros((B, m)))
    st, it = buf(np.full(B, -1, dtype=np.int32)), buf(np.zeros(B, dtype=np.int32))
    vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    ops = _capi.Operands(*[vp(host[k]) if k in host else None for k in ("A", "B", "C", "D", "e", "x0", "goal", "targets")])
    outs = _capi.Outputs(vp(U), vp(st), vp(it), vp(Z))
    desc = prob.desc()
    lib = _capi.load()
    for rep in range(3):
        U.zero_()
        assert lib.qpmpc_b200_solve_host(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs), 0) == 0
        assert torch.equal(U, plan.inputs.reshape(B, n).cpu())
        assert torch.equal(st, plan.status.cpu()) and torch.equal(it, plan.iters.cpu())
        assert torch.equal(Z, plan.multipliers.cpu())


@pytest.mark.parametrize("factored", [False, True, "two_launches_per_cycle"])
def test_pendulum_closed_loop_1024_instances_200_cycles_against_the_cpu_loop(factored, monkeypatch):
    """BASELINE config 3 at full length: 1 024 instances x 200 receding-horizon cycles on the
    device (re-condensing every cycle, or with the model factored once: qpmpc_b200_factor +
    qpmpc_b200_solve_factored) against the CPU loop (oracle + host plant), state by state for
    every instance that stays upright; the per-cycle iteration counts (summed over the batch) and
    the number of (instance, cycle) pairs solved from an upright state agree too."""
    import torch

    from qpmpc_b200 import factor_model, pendulum_closed_loop
    from qpmpc_b200.workloads import pendulum_batch, pendulum_targets, to_batched

    if factored == "two_launches_per_cycle":  # (True: the whole loop inside ONE launch of the shared-model kernel)
        monkeypatch.setenv("QPMPC_B200_LOOP_FUSED", "0")
    B, cycles = 1024, 200
    w = pendulum_batch(B, seed=1)
    ref, ref_iters = _cpu_closed_loop(w, cycles, return_iters=True)
    prob = to_batched(w)
    model = None
    if factored:
        tg, goal = pendulum_targets(w["x0"], w["v_target"], w["N"], w["T"])
        prob.update_goal_state(goal)
        prob.update_target_states(tg)
        model = factor_model(prob)
    plan, traj, unsolved, stats = pendulum_closed_loop(prob, w["v_target"], cycles, record=True, stats=True,
                                                      factored=model if model is not None else False)
    torch.cuda.synchronize()
    assert int(unsolved.item()) == 0
    got = traj.cpu().numpy()
    upright = np.abs(ref[:, :, 1]).max(axis=0) < 1.2
    assert upright.mean() > 0.98
    assert np.array_equal(np.abs(got[:, :, 1]).max(axis=0) < 1.2, upright)
    assert np.abs(got[:, upright] - ref[:, upright]).max() <= 1e-6
    assert int(stats["upright"].item()) == int((np.abs(ref[:-1, :, 1]) <= 1.2).sum())
    it = stats["iterations"].cpu().numpy()
    # the iteration histogram: identical while every instance is upright, and the loop is mostly
    # unconstrained once it has converged (the bound is active in the first cycles only)
    first_fall = int(np.argmax((np.abs(ref[:, :, 1]) > 1.2).any(axis=1))) if (~upright).any() else cycles
    assert np.array_equal(it[:first_fall], ref_iters.sum(axis=1)[:first_fall])
    assert it[0] > it[-1]


def test_factored_model_matches_the_full_path_on_the_device():
    """qpmpc_b200_factor + qpmpc_b200_solve_factored against qpmpc_b200_solve and the oracle:
    pendulum (stage cost, D rows), humanoid (per-instance per-step e_k), shared triple integrator
    at N = 8 / 32."""
    import torch

    from qpmpc_b200 import factor_model, solve_mpc_batch
    from qpmpc_b200.workloads import humanoid_batch, pendulum_batch, to_batched, triple_integrator_batch

    ti32 = triple_integrator_batch(300, N=32, seed=5, per_instance_model=False)
    for w in (pendulum_batch(4099, seed=3), pendulum_batch(515, seed=4, ltv_model=True), humanoid_batch(2050, seed=6),
              triple_integrator_batch(1000, N=8, seed=7, per_instance_model=False), ti32):
        prob = to_batched(w)
        full = solve_mpc_batch(prob, return_multipliers=True)
        fast = solve_mpc_batch(prob, return_multipliers=True, factored=factor_model(prob))
        torch.cuda.synchronize()
        ref = _oracle(w)
        ok = ref["status"] == 0
        assert np.array_equal(fast.status.cpu().numpy() == 0, ok) and torch.equal(fast.status, full.status)
        assert torch.equal(fast.iters, full.iters)
        U = fast.inputs.reshape(w["batch"], -1).cpu().numpy()
        assert np.abs(U[ok] - ref["U"][ok]).max() <= U_TOL


@pytest.mark.parametrize("factored", [False, True, "two_launches_per_cycle"])
def test_lipm_walking_closed_loop_300_cycles_against_the_cpu_loop(factored, monkeypatch):
    """The walking controller of examples/lipm_walking_controller.py:307-335 (LTV constraint
    vector rewritten every cycle by the phase machine, goal update, 300 cycles), 256 instances
    on the device against the CPU loop (numpy phase machine + oracle): states of every cycle,
    support foot and phase at the end."""
    import torch

    from qpmpc_b200 import factor_model, lipm_walking_closed_loop
    from qpmpc_b200.workloads import lipm_advance, lipm_phase_vectors, lipm_walking_batch, to_batched

    if factored == "two_launches_per_cycle":  # (True: the whole loop inside ONE launch of the shared-model kernel)
        monkeypatch.setenv("QPMPC_B200_LOOP_FUSED", "0")
    B, cycles = 256, 300
    w = lipm_walking_batch(B, seed=4)
    x, foot = w["x0"].copy(), w["support_foot"].copy()
    pidx, sidx = w["phase_index"].copy(), w["stride_index"].copy()
    ref, wc, active = [x.copy()], dict(w), 0
    for _ in range(cycles):
        wc["x0"] = x
        wc["e"], wc["goal"] = lipm_phase_vectors(wc, foot, pidx, sidx)
        sol = _oracle(wc)
        assert (sol["status"] == 0).all()
        active += int((sol["iters"] > 0).sum())
        x, foot, pidx, sidx = lipm_advance(wc, x, sol["U"][:, 0], foot, pidx, sidx)
        ref.append(x.copy())
    ref = np.stack(ref)
    prob = to_batched(w)
    model = factor_model(prob) if factored else None
    plan, traj, unsolved, phase = lipm_walking_closed_loop(prob, w["support_foot"], w["strides"], w["phase_index"],
                                                           w["stride_index"], cycles, record=True,
                                                           factored=model if model is not None else False)
    torch.cuda.synchronize()
    assert int(unsolved.item()) == 0
    assert np.abs(traj.cpu().numpy() - ref).max() <= 1e-6
    assert np.abs(phase["support_foot"].cpu().numpy() - foot).max() <= 1e-12
    assert np.array_equal(phase["phase_index"].cpu().numpy(), pidx)
    assert np.array_equal(phase["stride_index"].cpu().numpy(), sidx)
    assert active > 0.5 * B * cycles  # the ZMP bounds bind in most cycles: not a degenerate workload


@pytest.mark.parametrize("N,ltv", [(64, False), (64, True), (40, True), (33, False)])
def test_tensor_core_hessian_matches_the_simt_sum_and_the_oracle(N, ltv, monkeypatch):
    """Single precision, 32 < n <= 64, stage cost: P = w_u I + w_x Psi'Psi + w_t psi_N'psi_N of
    ``condense_batch`` comes from tcgen05.mma.kind::tf32 with a hi/lo split of the operands
    (mpc_hessian_tc.cuh).  Bar: 2e-6 relative to max|P| against the fp64 oracle -- the SIMT fp32
    sum (QPMPC_B200_HESSIAN_TC=0) is held to the same bar -- i.e. the split recovers full fp32
    accuracy from TF32 products; the other fields are untouched."""
    import torch

    import oracle
    from qpmpc_b200 import _capi, condense_batch
    from qpmpc_b200.workloads import oracle_ops, to_batched, triple_integrator_batch

    B = 70
    # a model whose powers stay bounded over the horizon: triple integrator, 1 s horizon, stage cost
    w = triple_integrator_batch(B, N=N, seed=N, per_instance_model=False)
    w["w_x"], w["w_t"], w["w_u"] = 0.5, 2.0, 1e-3
    w["targets"] = np.random.default_rng(N).standard_normal((B, N * 3))
    if ltv:
        w["A"], w["B"] = np.tile(w["A"], (N, 1, 1)), np.tile(w["B"], (N, 1, 1))
        w["ltv"] = ("A", "B")
    fields = ("P", "q", "G", "h", "Phi", "Psi", "phi_last", "psi_last")
    prob = to_batched(w, dtype=torch.float32)
    before = _capi.launch_count()
    tc = condense_batch(prob, fields)
    torch.cuda.synchronize()
    launches_tc = _capi.launch_count() - before
    monkeypatch.setenv("QPMPC_B200_HESSIAN_TC", "0")
    before = _capi.launch_count()
    simt = condense_batch(prob, fields)
    torch.cuda.synchronize()
    assert launches_tc == (_capi.launch_count() - before) + 1  # the extra launch is the tensor-core kernel
    ops = oracle_ops(w)
    worst_tc = worst_simt = 0.0
    for b in (0, 1, B - 1):
        pick = lambda k: None if ops[k][0] is None else (ops[k][0][b] if ops[k][1] else ops[k][0])  # noqa: E731
        ref = oracle.condense(w["N"], 3, 1, 2, pick("A"), pick("B"), pick("C"), None, pick("e"), w["x0"][b], w["goal"][b],
                              w["targets"][b], w["w_t"], w["w_x"], w["w_u"])
        scale = np.abs(ref["P"]).max()
        worst_tc = max(worst_tc, np.abs(tc["P"][b].double().cpu().numpy() - ref["P"]).max() / scale)
        worst_simt = max(worst_simt, np.abs(simt["P"][b].double().cpu().numpy() - ref["P"]).max() / scale)
    assert worst_tc <= 2e-6 and worst_simt <= 1e-5, (worst_tc, worst_simt)
    for k in fields[1:]:
        assert torch.equal(tc[k], simt[k]), k


def test_solve_mpc_single_path_matches_batch_path_on_the_device(monkeypatch):
    """``solve_mpc`` of ONE problem: the zero-copy host entry (qpmpc_b200/single.py) against the
    device-tensor path, bit for bit; and what a control loop pays per cycle."""
    import time

    from qpmpc_b200 import solve_mpc
    from test_host import single_path_matches_batch_path

    single_path_matches_batch_path(monkeypatch)
    problem = golden_problem(load_golden("triple_integrator"))
    for mode in ("1", "0"):
        monkeypatch.setenv("QPMPC_B200_SINGLE_ZEROCOPY", mode)
        for _ in range(20):
            solve_mpc(problem, "b200")
        t0 = time.perf_counter()
        for _ in range(200):
            solve_mpc(problem, "b200")
        print(f"solve_mpc latency, QPMPC_B200_SINGLE_ZEROCOPY={mode}: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
