"""Synthetic workloads of the BASELINE configs (host-side NumPy, seeded).

Problem data follow the reference's examples:
  * triple integrator -- ``examples/triple_integrator.py:15-42``
  * humanoid / LIPM with per-step ZMP bounds -- ``examples/humanoid_one_step.py:32-80``
  * wheeled inverted pendulum -- ``qpmpc/systems/wheeled_inverted_pendulum.py:64-125``
    with the targets of ``examples/wheeled_inverted_pendulum.py:65-83``
Distributions and seeds are the ones fixed in SURVEY.md section 8(d).  Each
generator returns a dict of arrays in the canonical operand layouts of
``include/qpmpc_b200.h`` (per-instance operands carry a leading batch axis)
plus the scalar weights; :func:`to_batched` turns it into a
``BatchedMPCProblem`` and :func:`oracle_ops` into the oracle's argument map.
"""

from typing import Dict, Optional

import numpy as np

GRAVITY = 9.81


def triple_integrator_matrices(N: int, horizon: float = 1.0):
    """A, B, C, e of the triple integrator with sampling period horizon / N."""
    T = horizon / N
    A = np.array([[1.0, T, T**2 / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]])
    B = np.array([[T**3 / 6.0], [T**2 / 2.0], [T]])
    C = np.array([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0]])
    e = np.array([3.0, 3.0])
    return A, B, C, e


def triple_integrator_batch(batch: int, N: int = 16, seed: int = 0, per_instance_model: bool = True,
                            jitter: float = 0.0) -> Dict:
    """BASELINE config 2 / 5: random initial and goal states, |accel| <= 3.

    ``per_instance_model`` replicates A, B, C, e for every instance (the
    "per-instance LTI" record the metric is quoted on); ``jitter`` perturbs the
    sampling period per instance by +-jitter so the models really differ.
    """
    rng = np.random.default_rng(seed)
    x0 = np.stack([rng.uniform(-1, 1, batch), rng.uniform(-1, 1, batch),
                   rng.uniform(-2.5, 2.5, batch)], axis=1)
    goal = np.stack([rng.uniform(-1, 1, batch), np.zeros(batch), np.zeros(batch)], axis=1)
    A, B, C, e = triple_integrator_matrices(N)
    if per_instance_model:
        if jitter > 0.0:
            scale = 1.0 + rng.uniform(-jitter, jitter, batch)
            T = scale / N
            Ab = np.tile(np.eye(3), (batch, 1, 1))
            Ab[:, 0, 1] = T
            Ab[:, 0, 2] = T**2 / 2.0
            Ab[:, 1, 2] = T
            Bb = np.stack([T**3 / 6.0, T**2 / 2.0, T], axis=1)[:, :, None]
        else:
            Ab = np.tile(A, (batch, 1, 1))
            Bb = np.tile(B, (batch, 1, 1))
        A, B = Ab, Bb
        C = np.tile(C, (batch, 1, 1))
        e = np.tile(e, (batch, 1))
    return dict(name=f"triple_integrator_N{N}", batch=batch, N=N, nx=3, nu=1, nc=2,
                A=A, B=B, C=C, D=None, e=e, x0=x0, goal=goal, targets=None,
                w_t=1.0, w_x=None, w_u=1e-6, ltv=())


def humanoid_batch(batch: int, N: int = 16, seed: int = 2, com_height: float = 0.8,
                   horizon: float = 2.5, dsp: float = 0.1, ssp: float = 0.7,
                   foot_length: float = 0.1, big: float = 1000.0) -> Dict:
    """BASELINE config 4: LIPM stepping, per-instance per-step ZMP bounds e_k."""
    rng = np.random.default_rng(seed)
    T = horizon / N
    A = np.array([[1.0, T, T**2 / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]])
    B = np.array([[T**3 / 6.0], [T**2 / 2.0], [T]])
    zmp = np.array([1.0, 0.0, -com_height / GRAVITY])
    C = np.array([+zmp, -zmp])
    start = rng.uniform(-0.1, 0.1, batch)
    end = start + rng.uniform(0.2, 0.4, batch)
    n0, n1, n2 = (int(round(v / T)) for v in (dsp, ssp, dsp))
    e = np.empty((batch, N, 2))
    for i in range(N):
        if i < n0 or (i - n0 > n1 and i - n0 - n1 < n2):
            e[:, i, :] = big
        elif i - n0 <= n1:
            e[:, i, 0] = start + 0.5 * foot_length
            e[:, i, 1] = -(start - 0.5 * foot_length)
        else:
            e[:, i, 0] = end + 0.5 * foot_length
            e[:, i, 1] = -(end - 0.5 * foot_length)
    x0 = np.stack([start, np.zeros(batch), np.zeros(batch)], axis=1)
    goal = np.stack([end, np.zeros(batch), np.zeros(batch)], axis=1)
    return dict(name="humanoid", batch=batch, N=N, nx=3, nu=1, nc=2, A=A, B=B, C=C, D=None,
                e=e, x0=x0, goal=goal, targets=None, w_t=1.0, w_x=None, w_u=1e-3, ltv=())


def pendulum_matrices(length: float = 0.6, T: float = 0.1):
    """ZOH-discretised wheeled-inverted-pendulum model (systems/...:86-102)."""
    w = np.sqrt(GRAVITY / length)
    ch, sh = np.cosh(T * w), np.sinh(T * w)
    A = np.array([[1.0, 0.0, T, 0.0], [0.0, ch, 0.0, sh / w],
                  [0.0, 0.0, 1.0, 0.0], [0.0, w * sh, 0.0, ch]])
    B = np.array([[T**2 / 2.0], [-ch / GRAVITY + 1.0 / GRAVITY], [T], [-w * sh / GRAVITY]])
    return A, B


def pendulum_targets(x0: np.ndarray, v_target: np.ndarray, N: int, T: float):
    """(targets [B, N*4], goal [B, 4]) as get_target_states builds them
    (examples/wheeled_inverted_pendulum.py:65-83)."""
    batch = x0.shape[0]
    ts = np.zeros((batch, N + 1, 4))
    k = np.arange(N + 1)[None, :]
    ts[:, :, 0] = x0[:, 0:1] + k * T * v_target[:, None]
    ts[:, :, 2] = v_target[:, None]
    return ts[:, :N].reshape(batch, N * 4), ts[:, N]


def pendulum_batch(batch: int, N: int = 12, seed: int = 1, T: float = 0.1, max_accel: float = 10.0,
                   ltv_model: bool = False) -> Dict:
    """BASELINE config 3 (one cycle): |u| <= max_accel via D, terminal + stage cost."""
    rng = np.random.default_rng(seed)
    A, B = pendulum_matrices(T=T)
    x0 = np.stack([np.zeros(batch), rng.uniform(-0.4, 0.4, batch), np.zeros(batch),
                   rng.uniform(-1.5, 1.5, batch)], axis=1)
    v = rng.uniform(0.0, 1.0, batch)
    targets, goal = pendulum_targets(x0, v, N, T)
    D = np.array([[1.0], [-1.0]])
    e = np.array([max_accel, max_accel])
    ltv = ()
    if ltv_model:
        A, B = np.tile(A, (N, 1, 1)), np.tile(B, (N, 1, 1))
        ltv = ("A", "B")
    return dict(name="pendulum", batch=batch, N=N, nx=4, nu=1, nc=2, A=A, B=B, C=None, D=D, e=e,
                x0=x0, goal=goal, targets=targets, w_t=10.0, w_x=1.0, w_u=1e-3, ltv=ltv,
                v_target=v, T=T)


def lipm_walking_batch(batch: int, N: int = 16, seed: int = 4, T: float = 0.1, com_height: float = 0.84) -> Dict:
    """The walking controller of ``examples/lipm_walking_controller.py`` (reference tree) for a
    batch: shared triple-integrator model with T = 0.1 s and ZMP rows C = [+zmp; -zmp]
    (:62-101), per-instance strides, initial support foot and initial phase.  e (per instance
    and step) and the goal are rewritten every cycle by the phase machine; the arrays here are
    those of the first cycle (:func:`lipm_phase_vectors`).  Parameters of :31-59."""
    rng = np.random.default_rng(seed)
    A = np.array([[1.0, T, T**2 / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]])
    B = np.array([[T**3 / 6.0], [T**2 / 2.0], [T]])
    omega2 = GRAVITY / com_height
    zmp = np.array([1.0, 0.0, -1.0 / omega2])
    C = np.array([+zmp, -zmp])
    stride = rng.uniform(0.12, 0.24, batch)
    strides = np.stack([-stride, stride], axis=1)           # :47  strides = [-0.18, 0.18]
    foot = rng.uniform(0.07, 0.11, batch)                   # :42  init_support_foot_pos = 0.09
    phase_index = rng.integers(3, 7, batch).astype(np.int32)  # :116 initial_index = 5
    stride_index = np.zeros(batch, dtype=np.int32)
    # initial ZMP at the centre of the initial foothold, DCM halfway (:296-299)
    omega = np.sqrt(omega2)
    x0 = np.stack([np.zeros(batch), 0.5 * omega * foot, -omega2 * foot], axis=1)
    w = dict(name="lipm_walking", batch=batch, N=N, nx=3, nu=1, nc=2, A=A, B=B, C=C, D=None, e=None, x0=x0,
             goal=None, targets=None, w_t=1.0, w_x=None, w_u=1e-3, ltv=(), T=T, strides=strides,
             support_foot=foot, phase_index=phase_index, stride_index=stride_index, foot_size=0.065,
             nb_dsp=int(round(0.1 / T)), nb_ssp=int(round(0.7 / T)), max_zmp=100.0)
    w["e"], w["goal"] = lipm_phase_vectors(w, foot, phase_index, stride_index)
    return w


def lipm_phase_vectors(w: Dict, foot, phase_index, stride_index):
    """(e [B, N, 2], goal [B, 3]) of one control cycle: ``update_goal_and_constraints``
    (``examples/lipm_walking_controller.py:175-205``) with ``PhaseStepper.get_nb_steps``
    (:132-163), vectorised over the instances."""
    N, nd, ns, B = w["N"], w["nb_dsp"], w["nb_ssp"], len(foot)
    off = phase_index.astype(np.int64)
    n0 = np.maximum(0, nd - off)
    off = np.maximum(0, off - nd)
    n1 = np.maximum(0, ns - off)
    rem = N - n0 - n1
    n2 = np.minimum(nd, rem)
    rem = np.maximum(0, rem - nd)
    n3 = np.minimum(ns, rem)
    rem = np.maximum(0, rem - ns)
    n4 = np.minimum(nd, rem)
    rows = np.arange(B)
    nxt = foot + w["strides"][rows, stride_index]
    last = nxt + w["strides"][rows, (stride_index + 1) % 2]
    hf, big = 0.5 * w["foot_size"], w["max_zmp"]
    k = np.arange(N)[None, :]
    c0, c1, c2, c3, c4 = (np.cumsum([n0, n1, n2, n3, n4], axis=0)[i][:, None] for i in range(5))
    centre = np.where(k < c1, foot[:, None], np.where(k < c3, nxt[:, None], last[:, None]))
    free = (k < c0) | ((k >= c1) & (k < c2)) | ((k >= c3) & (k < c4))
    e = np.stack([np.where(free, big, centre + hf), np.where(free, big, -(centre - hf))], axis=2)
    goal = np.stack([np.where(n4 > 0, last, nxt), np.zeros(B), np.zeros(B)], axis=1)
    return e, goal


def lipm_advance(w: Dict, x, u, foot, phase_index, stride_index, substeps: int = 15):
    """State, support foot and phase after one control cycle under jerk u: ``integrate``
    (``examples/lipm_walking_controller.py:208-226``), ``PhaseStepper.advance`` (:124-130) and
    the foot switch (:331-334), vectorised."""
    dt = w["T"] / substeps
    p, v, a = x[:, 0].copy(), x[:, 1].copy(), x[:, 2].copy()
    for _ in range(substeps):
        p, v, a = (p + dt * (v + dt * (a / 2 + dt * u / 6)), v + dt * (a + dt * (u / 2)), a + dt * u)
    index = np.where(phase_index + 1 >= w["nb_dsp"] + w["nb_ssp"], 0, phase_index + 1).astype(np.int32)
    switch = index == 0
    rows = np.arange(len(foot))
    foot = np.where(switch, foot + w["strides"][rows, stride_index], foot)
    stride_index = np.where(switch, (stride_index + 1) % 2, stride_index).astype(np.int32)
    return np.stack([p, v, a], axis=1), foot, index, stride_index


def random_batch(batch: int, N: int, nx: int, nu: int, nc: int, seed: int = 0, with_C: bool = True,
                 with_D: bool = True, w_t: Optional[float] = 0.7, w_x: Optional[float] = 0.3,
                 w_u: float = 1e-2, ltv: bool = True) -> Dict:
    """Fully per-instance random problems (LTV when ``ltv``), always feasible at U = 0
    when x0 is small."""
    rng = np.random.default_rng(seed)
    lead = (batch, N) if ltv else (batch,)
    A = np.eye(nx) + 0.2 * rng.standard_normal(lead + (nx, nx))
    B = rng.standard_normal(lead + (nx, nu))
    C = rng.standard_normal(lead + (nc, nx)) if with_C else None
    D = rng.standard_normal(lead + (nc, nu)) if with_D else None
    e = 1.0 + rng.random(lead + (nc,))
    return dict(name=f"random_{N}_{nx}_{nu}_{nc}", batch=batch, N=N, nx=nx, nu=nu, nc=nc, A=A, B=B,
                C=C, D=D, e=e, x0=0.1 * rng.standard_normal((batch, nx)),
                goal=rng.standard_normal((batch, nx)),
                targets=rng.standard_normal((batch, N * nx)), w_t=w_t, w_x=w_x, w_u=w_u, ltv=())


# -- adapters ---------------------------------------------------------------


def to_batched(w: Dict, dtype=None, device=None):
    """Workload dict -> BatchedMPCProblem on the CUDA device."""
    import torch

    from .batched import BatchedMPCProblem

    return BatchedMPCProblem(
        w["A"], w["B"], w["C"], w["D"], w["e"], w["N"], w["w_t"], w["w_x"], w["w_u"],
        initial_state=w["x0"], goal_state=w["goal"], target_states=w["targets"],
        ltv=w.get("ltv", ()), batch_size=w["batch"],
        dtype=dtype or torch.float64, device=device,
    )


def operand_layout(w: Dict, name: str):
    """(per_instance, per_step) of a matrix operand of a workload dict."""
    arr = w[name]
    if arr is None:
        return False, False
    base = 1 if name == "e" else 2
    extra = arr.ndim - base
    if extra == 0:
        return False, False
    if extra == 2:
        return True, True
    return (False, True) if name in w.get("ltv", ()) else (True, False)


def oracle_ops(w: Dict) -> Dict:
    """Argument map of ``oracle.solve_batch`` for a workload dict."""
    ops = {}
    for name in ("A", "B", "C", "D", "e"):
        pi, ps = operand_layout(w, name)
        ops[name] = (w[name], pi, ps)
    for name in ("x0", "goal", "targets"):
        arr = w[name]
        ops[name] = (arr, arr is not None and arr.ndim == 2)
    return ops


def rows_are_paired(w: Dict) -> bool:
    """True when every C_k, D_k of the workload has the form [M; -M] (two-sided bounds): what
    ``qpmpc_b200_desc.paired`` promises to the kernels."""
    nc = w["nc"]
    if nc == 0 or nc % 2:
        return False
    h = nc // 2
    for name in ("C", "D"):
        arr = w[name]
        if arr is not None and not np.array_equal(arr[..., :h, :], -arr[..., h:, :]):
            return False
    return w["C"] is not None or w["D"] is not None


def slice_workload(w: Dict, lo: int, hi: int) -> Dict:
    """Instances [lo, hi) of a workload (per-instance operands sliced)."""
    out = dict(w)
    out["batch"] = hi - lo
    for name in ("A", "B", "C", "D", "e"):
        if w[name] is not None and operand_layout(w, name)[0]:
            out[name] = w[name][lo:hi]
    for name in ("x0", "goal", "targets", "v_target"):
        arr = w.get(name)
        if arr is not None and getattr(arr, "ndim", 0) >= 1 and arr.shape[0] == w["batch"]:
            out[name] = arr[lo:hi]
    return out


def algorithmic_bytes_per_solve(w: Dict, itemsize: int = 8) -> int:
    """Compulsory HBM bytes of one solve: per-instance inputs
    read once, U written once, plus the 4-byte status."""
    total = 0
    sizes = dict(A=w["nx"] ** 2, B=w["nx"] * w["nu"], C=w["nc"] * w["nx"],
                 D=w["nc"] * w["nu"], e=w["nc"])
    for name, item in sizes.items():
        pi, ps = operand_layout(w, name)
        if w[name] is not None and pi:
            total += item * (w["N"] if ps else 1)
    for name, size in (("x0", w["nx"]), ("goal", w["nx"]), ("targets", w["N"] * w["nx"])):
        arr = w[name]
        if arr is not None and arr.ndim == 2:
            total += size
    return itemsize * (total + w["N"] * w["nu"]) + 4
