#!/usr/bin/env python3
"""bench.py -- MPC solves/sec (batched), the BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|3|4|5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs; SURVEY.md 8(d) for shapes, distributions, seeds):
  --config 2 (default, the configuration the metric is quoted on): triple integrator
             (nx=3, nu=1, nc=2, N=16), fp64, 65 536 instances per GPU, per-instance
             A, B, C, e, x0, goal records.  A step condenses and solves the batch once.
  --config 3 wheeled inverted pendulum (nx=4, nu=1, N=12), 16 384 instances, a step is one
             200-cycle receding-horizon closed loop (targets -> solve -> 15 plant substeps).
  --config 4 humanoid / LIPM with per-instance per-step ZMP bounds e_k, fp32, 8 192 instances.
  --config 5 one point of the horizon sweep: triple integrator with T = 1/N,
             --horizon N in {8,16,32,64}, --batch B (default N=64, B=262 144).
Weak scaling: every rank owns its own batch; for N > 1 (config 2) every rank ends each step
with the stacked U of ALL ranks -- the all-gather the north star names, fused into the solve
kernel (its epilogue stores each row into every rank's symmetric buffer over NVLink; NCCL
all-gather is the fallback) -- and the gathered buffer is checked, after the timed region,
against a local re-solve of the other ranks' shards.

`value`     device-resident throughput (inputs already in HBM), CUDA events, max over ranks.
`e2e`       the same batch through qpmpc_b200_solve_host with pinned HOST buffers: H2D of
            every operand + kernel + D2H of U/status per step; at N > 1 the ranks write their
            rows into ONE host array shared by the ranks (the gather), barrier per step.
`roofline`  algorithmic bytes per launch / measured kernel time vs the HBM peak -- by
            construction tiny: the path is FP64-issue bound, so two FP64 fractions are
            reported next to it (`fp64`): SURVEY 8(d)'s structure-agnostic flop count F and
            the flops this algorithm executes, both against a measured DFMA peak.
`cpu_baseline`  the C oracle (oracle/mpc_oracle.c: condensing + Goldfarb-Idnani, OpenMP)
            on a bounded sample of the same instances, all host threads.

`--impl reference` times that CPU port alone on all host threads, on the whole job's
batch (world x batch) -- the reference itself is pure Python over qpsolvers wheels that
cannot be installed offline; see DESIGN.md.
"""

import argparse
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ROTATE_BYTES = 160e6  # distinct input sets must exceed the 126 MB L2
METRIC = "MPC solves/sec (batched)"
UNIT = "solves/s"
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback
DEFAULTS = {2: (65536, 16), 3: (16384, 12), 4: (8192, 16), 5: (262144, 64), 6: (8192, 16)}  # config -> (batch/GPU, N)
CYCLES = 200  # config 3
WALK_CYCLES = 300  # config 6 (examples/lipm_walking_controller.py:306)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5, 6])
    ap.add_argument("--f32", action="store_true",
                    help="config 6 only: run the walking loop in single precision (the precision BASELINE configs[3] names)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--horizon", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--method", default="active_set", choices=["active_set", "pdip"],
                    help="solver kernel; the headline and the default is the exact active set")
    args = ap.parse_args()
    b, n = DEFAULTS[args.config]
    args.batch = args.batch or b
    args.horizon = args.horizon or n
    if args.steps is None:
        args.steps = {2: 1024, 3: 8, 4: 1024, 5: 64, 6: 8}[args.config] if args.impl == "b200" else 8
    if args.warmup is None:
        args.warmup = {2: 16, 3: 3, 4: 16, 5: 4, 6: 3}[args.config] if args.impl == "b200" else 3
    return args


# ---------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------
def make_workload(args, seed):
    from qpmpc_b200.workloads import humanoid_batch, lipm_walking_batch, pendulum_batch, triple_integrator_batch

    if args.config == 6:
        return lipm_walking_batch(args.batch, N=args.horizon, seed=4 + seed)
    if args.config == 3:
        return pendulum_batch(args.batch, N=args.horizon, seed=1 + seed)
    if args.config == 4:
        return humanoid_batch(args.batch, N=args.horizon, seed=2 + seed)
    return triple_integrator_batch(args.batch, N=args.horizon, seed=(0 if args.config == 2 else 3) + seed)


def is_f32(args):
    return args.config == 4 or (args.config == 6 and getattr(args, "f32", False))


def dtype_of(args):
    return "f32" if is_f32(args) else "f64"


def solves_per_step(args):
    return args.batch * (CYCLES if args.config == 3 else WALK_CYCLES if args.config == 6 else 1)


def config_dict(args, world, rotate=None):
    B, N = args.batch, args.horizon
    names = {
        2: f"triple_integrator fp64 nx=3 nu=1 nc=2 N={N} batch={B}/GPU per-instance A,B,C,e,x0,goal "
           "(BASELINE configs[1])",
        3: f"wheeled_inverted_pendulum fp64 nx=4 nu=1 nc=2 N={N} batch={B}/GPU, {CYCLES}-cycle receding horizon "
           "(targets -> condense+solve -> 15 plant substeps per cycle), shared model (BASELINE configs[2])",
        4: f"humanoid/LIPM fp32 nx=3 nu=1 nc=2 N={N} batch={B}/GPU per-instance per-step e_k (BASELINE configs[3])",
        5: f"triple_integrator T=1/N fp64 nx=3 nu=1 nc=2 N={N} batch={B}/GPU (BASELINE configs[4], one sweep point)",
        6: f"lipm_walking_controller {'fp32' if is_f32(args) else 'fp64'} nx=3 nu=1 nc=2 N={N} batch={B}/GPU, {WALK_CYCLES}-cycle walking loop: per "
           "cycle the phase machine rewrites the per-step ZMP bounds e_k and the goal, condense+solve, 15 "
           "integration substeps (the closed-loop form of BASELINE configs[3]'s LIPM LTV constraints; "
           "examples/lipm_walking_controller.py)",
    }
    cfg = {
        "workload": names[args.config],
        "batch_per_gpu": B, "global_batch": B * world, "horizon": N,
        "method": "dual active set (Goldfarb-Idnani), exact" if args.method == "active_set"
        else "interior point (Mehrotra) + primal-dual active-set polish, tol 1e-9",
        "parallelism": f"batch-sharded x{world}, U gathered on every rank each step" if world > 1 else "single GPU",
    }
    if rotate is not None:
        cfg["l2"] = rotate
    return cfg


# ---------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ---------------------------------------------------------------------------
def host_threads():
    """Every core this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its
    workers; the CPU arm ignores it and asks the oracle for this many threads."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_solve(workload, threads):
    import oracle
    from qpmpc_b200.workloads import oracle_ops

    return oracle.solve_batch(workload["batch"], workload["N"], workload["nx"], workload["nu"], workload["nc"],
                              oracle_ops(workload), workload["w_t"], workload["w_x"], workload["w_u"],
                              nthreads=threads)


def cpu_closed_loop(workload, cycles, threads, substeps=15):
    """Config 3 on the host: the closed loop of examples/wheeled_inverted_pendulum.py:99-118
    with the oracle as the solver and the host mirror of the plant."""
    from qpmpc_b200.systems import WheeledInvertedPendulum
    from qpmpc_b200.workloads import pendulum_targets

    pend = WheeledInvertedPendulum()
    w = dict(workload)
    x = w["x0"].copy()
    T = w["T"]
    for _ in range(cycles):
        w["targets"], w["goal"] = pendulum_targets(x, w["v_target"], w["N"], T)
        w["x0"] = x
        ref = cpu_solve(w, threads)
        u0 = np.where(ref["status"] == 0, ref["U"][:, 0], 0.0)
        # WheeledInvertedPendulum.integrate (systems/wheeled_inverted_pendulum.py:127-160 of the
        # reference), vectorised over the instances
        dt = T / substeps
        for _ in range(substeps):
            pa = pend.omega**2 * (np.sin(x[:, 1]) - (u0 / pend.GRAVITY) * np.cos(x[:, 1]))
            x = np.stack([x[:, 0] + dt * x[:, 2] + 0.5 * dt * dt * u0, x[:, 1] + dt * x[:, 3] + 0.5 * dt * dt * pa,
                          x[:, 2] + dt * u0, x[:, 3] + dt * pa], axis=1)
    return x


def cpu_walking_loop(workload, cycles, threads):
    """Config 6 on the host: examples/lipm_walking_controller.py:307-335 with the numpy phase
    machine of qpmpc_b200.workloads and the oracle as the solver."""
    from qpmpc_b200.workloads import lipm_advance, lipm_phase_vectors

    w = dict(workload)
    x, foot = w["x0"].copy(), w["support_foot"].copy()
    pidx, sidx = w["phase_index"].copy(), w["stride_index"].copy()
    for _ in range(cycles):
        w["x0"] = x
        w["e"], w["goal"] = lipm_phase_vectors(w, foot, pidx, sidx)
        ref = cpu_solve(w, threads)
        u0 = np.where(ref["status"] == 0, ref["U"][:, 0], 0.0)
        x, foot, pidx, sidx = lipm_advance(w, x, u0, foot, pidx, sidx)
    return x


def cpu_loop(args, workload, cycles, threads):
    return (cpu_closed_loop if args.config == 3 else cpu_walking_loop)(workload, cycles, threads)


def reference_python_condense(seconds=2.0):
    """the reference's OWN condensing (qpmpc.MPCQP, pure Python /
    NumPy) on one core, timed when the reference tree is importable (the development container);
    elsewhere (the GPU box has no copy of it) the container measurement of BASELINE.md is quoted.
    The solver half of the reference (qpsolvers wheels) is not installable offline."""
    ref_root = os.environ.get("QPMPC_REFERENCE", "/root/reference")
    quoted = {"mpcqp_build_us": 553.0, "builds_per_s_per_core": 1.0e6 / 553.0,
              "source": "quoted: BASELINE.md section 2 (measured in the development container; the reference "
                        "tree is not present on this machine)"}
    if not os.path.isdir(os.path.join(ref_root, "qpmpc")):
        return quoted
    try:
        import importlib
        import types

        stub = types.ModuleType("qpsolvers")
        stub.Problem = lambda *a, **k: None
        stub.Solution = object
        stub.solve_problem = lambda *a, **k: None
        saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "qpsolvers" or k.split(".")[0] == "qpmpc"}
        for k in list(saved):
            sys.modules.pop(k, None)
        sys.modules["qpsolvers"] = stub
        sys.path.insert(0, ref_root)
        try:
            ref = importlib.import_module("qpmpc")
            T = 1.0 / 16
            prob = ref.MPCProblem(
                transition_state_matrix=np.array([[1.0, T, T**2 / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]]),
                transition_input_matrix=np.array([T**3 / 6.0, T**2 / 2.0, T]).reshape((3, 1)),
                ineq_state_matrix=np.array([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0]]), ineq_input_matrix=None,
                ineq_vector=np.array([3.0, 3.0]), initial_state=np.zeros(3), goal_state=np.array([1.0, 0.0, 0.0]),
                nb_timesteps=16, terminal_cost_weight=1.0, stage_state_cost_weight=None, stage_input_cost_weight=1e-6)
            ref.MPCQP(prob)
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < seconds:
                ref.MPCQP(prob)
                n += 1
            us = (time.perf_counter() - t0) / n * 1e6
        finally:
            sys.path.remove(ref_root)
            for k in [k for k in sys.modules if k == "qpsolvers" or k.split(".")[0] == "qpmpc"]:
                sys.modules.pop(k, None)
            for k, v in saved.items():
                if v is not None:
                    sys.modules[k] = v
        return {"mpcqp_build_us": us, "builds_per_s_per_core": 1e6 / us,
                "source": f"measured here: {n} x qpmpc.MPCQP(triple integrator N=16) from {ref_root}, one core"}
    except Exception as exc:  # noqa: BLE001
        quoted["source"] += f" [import failed: {type(exc).__name__}]"
        return quoted


def cpu_arm(args, workload, seconds):
    """Time the oracle on a bounded sample of the workload for about `seconds`;
    returns (solves/s, threads, sample description)."""
    from qpmpc_b200.workloads import slice_workload

    threads = host_threads()
    if args.config in (3, 6):
        # the closed loop is sequential over cycles: a slice of the batch, fewer cycles
        k, cyc = min(workload["batch"], 16384), 20
        ws = slice_workload(workload, 0, k)
        t0 = time.perf_counter()
        cpu_loop(args, ws, cyc, threads)
        dt = time.perf_counter() - t0
        return k * cyc / dt, threads, (f"{k} instances x {cyc} closed-loop cycles (C oracle solve on {threads} "
                                       f"threads + NumPy plant step) in {dt:.1f} s")
    k = min(workload["batch"], 65536 if workload["N"] <= 16 else 16384 if workload["N"] <= 32 else 4096)
    ws = slice_workload(workload, 0, k)
    cpu_solve(ws, threads)  # warm-up (page in, thread pool)
    reps, t0 = 0, time.perf_counter()
    while True:
        cpu_solve(ws, threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return k * reps / dt, threads, f"{reps} x {k} instances of the bench workload (fp64) in {dt:.1f} s"


def run_reference(args, rank, world):
    """The reference arm: the C port of the path on ALL host threads, on the whole job's
    batch (world x batch instances per step), rank 0 only."""
    if rank != 0:
        return
    from qpmpc_b200.workloads import slice_workload

    threads = host_threads()
    if args.config in (3, 6):
        w = make_workload(args, 0)
        k, cyc = min(args.batch, 16384), 20
        ws = slice_workload(w, 0, k)
        for _ in range(min(args.warmup, 1)):
            cpu_loop(args, ws, 1, threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_loop(args, ws, cyc, threads)
        dt = time.perf_counter() - t0
        value, per_step = k * cyc * args.steps / dt, k * cyc
        sample = f"{args.steps} steps x ({k} instances x {cyc} closed-loop cycles): C oracle + NumPy plant"
    else:
        total = args.batch * world
        cap = 1 << 20 if args.horizon <= 16 else 1 << 17 if args.horizon <= 32 else 1 << 14
        k = min(total, cap)
        saved, args.batch = args.batch, k
        w = make_workload(args, 0)
        args.batch = saved
        for _ in range(max(args.warmup, 1)):
            cpu_solve(w, threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_solve(w, threads)
        dt = time.perf_counter() - t0
        value, per_step = k * args.steps / dt, k
        sample = (f"{args.steps} steps x {k} instances" + ("" if k == total else f" (of the job's {total})")
                  + ", C oracle (condense + Goldfarb-Idnani), OpenMP")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_of(args),
        "data": "synthetic", "config": config_dict(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "reference_python": reference_python_condense()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "instances_per_step": per_step,
        "note": "reference is pure Python over qpsolvers/proxqp wheels that are not installable "
                f"offline; this arm is the C port of its algorithm (oracle/), on {threads} host threads "
                "(OMP_NUM_THREADS exported by torchrun is ignored)",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# roofline helpers
# ---------------------------------------------------------------------------
def kernel_name(args):
    n = args.horizon  # nu = 1 in every config
    t = "float" if is_f32(args) else "double"
    if args.method == "pdip":
        return f"mpc_pdip_kernel<{t},NP={8 if n <= 8 else 16 if n <= 16 else 32}>"
    terminal_only = args.config != 3  # config 3 has a stage cost
    if args.config in (3, 6):
        return f"mpc_solve_kernel<{t},NP=16,paired rows,shared-model record>"
    if t == "double" and terminal_only and 16 < n <= 64:
        return f"mpc_solve_lr_kernel<double,NP={32 if n <= 32 else 64}>"  # structure-exploiting (rank-nx Hessian)
    if n > 32:
        return f"mpc_solve_cta_kernel<{t}>"
    return f"mpc_solve_kernel<{t},NP={8 if n <= 8 else 16 if n <= 16 else 32},paired rows>"


def measured_profile(args):
    """Figures of the newest committed ncu summary of this kernel and shape (profiles/r*_*.txt,
    written by tools/ncu_summary.py): dram bytes (read + write) of one launch, and the flops the
    launch executed (opcode counts of the profiler) per solve.  Missing entries are None --
    never a constant."""
    want = kernel_name(args).split("<")[0]
    tag = {2: "ti16", 3: "pend", 4: "hum", 5: f"ti{args.horizon}", 6: "walk"}[args.config]
    best = {"traffic": None, "flops_per_solve": None, "source": None}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_*.txt"))):
        base = os.path.basename(path)
        if tag not in base:
            continue
        try:
            text = open(path).read()
        except OSError:
            continue
        m = re.search(r"== kernel: void (?:qpmpc::)?(\w+)<(\w+),", text)
        if not m or m.group(1) != want or m.group(2) != ("float" if is_f32(args) else "double"):
            continue
        vals = {}
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mm = re.search(re.escape(key) + r"\s+([0-9.]+)\s+(\w?byte)", text)
            if mm:
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[mm.group(2)]
                vals[key] = float(mm.group(1)) * mult
        if len(vals) != 2:
            continue
        best = {"traffic": sum(vals.values()), "flops_per_solve": None, "source": base}
        mb = re.search(r"instances_per_launch (\d+)", text)
        mf = re.search(r"flops_fp(?:64|32)_per_launch (\d+)", text if args.config != 4 else text[text.find("fp32 warp"):])
        if mb and mf and int(mf.group(1)) > 0:
            best["flops_per_solve"] = int(mf.group(1)) / int(mb.group(1))
            best["profile_batch"] = int(mb.group(1))
    return best


def flop_models(args, iters_mean):
    """(F of SURVEY 8(d), executed flops of the active-set algorithm) per solve."""
    N, nu = args.horizon, 1
    nx = 4 if args.config == 3 else 3
    nc, n, m = 2, N * nu, 2 * N
    has_C = args.config != 3
    has_wx = args.config == 3
    if args.config in (3, 6):
        pass  # (the executed model below is replaced by the factored path's in run_b200)
    # dense, structure-agnostic condensing + K = 10 interior-point iterations
    f_cond = N * (2 * nx**3 + 2 * nx * nx * n + (2 * nc * nx * n + 2 * nc * nx * nx + 2 * nc * nx if has_C else 0)) \
        + 2 * nx * n * n + (2 * N * nx * n * n if has_wx else 0)
    f_iter = 2 * m * n * n + n**3 / 3 + 6 * n * n + 12 * m * n
    survey = f_cond + 10 * f_iter
    # executed by the active-set kernels at the measured mean iteration count (paired rows: one
    # stored row per [M; -M] pair, mp = m / 2 rows of M):
    mp = m // 2
    if 16 < n <= 64 and not has_wx and args.config != 4:
        # long-horizon kernel: matrix powers, 3 x 3 algebra, rows of M in O(n nx), one product with M
        # and one Householder update per iteration, recovery of x by Woodbury
        f_setup = 2 * nx**3 * 6 + 4 * nx * nx * 6 + mp * (4 * nx * n + 4 * n) + 8 * nx * n
        executed = f_setup + iters_mean * (4 * mp * n + n * n / 2)
    else:
        # warp kernel: matrix powers, Cholesky, forward substitutions for t and the rows of M, the
        # final two triangular solves; per iteration a product with M, a Householder update (two
        # passes) and the product with R^-1
        f_setup = 2 * nx**3 * 4 + 4 * nx * nx * 4 + n**3 / 3 + (mp + 1) * n * n + 2 * n * n
        executed = f_setup + iters_mean * (6 * mp * n + n * n / 2)
    return survey, executed


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU runs: keep this rank's threads -- and so, by first touch, the host buffers it
    allocates -- on the NUMA node its GPU hangs off (sysfs; None when the box does not say).
    QPMPC_B200_BENCH_NUMA=0 switches it off."""
    if os.environ.get("QPMPC_B200_BENCH_NUMA", "1") == "0":
        return None
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read())
        if node < 0:
            return None
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):  # no sysfs entry, no driver ...
        return None


def loop_kind(model):
    fused = model is not None and os.environ.get("QPMPC_B200_LOOP_FUSED", "1") != "0"
    return ("all cycles inside ONE launch of the shared-model solve kernel (each lane group solves, moves its plant "
            "and rewrites its vectors in shared memory)") if fused else "one solve launch and one plant launch per cycle"


def single_call_latency(calls=300):
    """BASELINE configs[0]: ONE triple-integrator problem through the reference-facing call,
    ``solve_mpc(problem)`` -> Plan (what a control loop pays per cycle; host objects in, host
    arrays out).  Outside the timed region of the headline."""
    from qpmpc_b200 import MPCProblem, solve_mpc

    T = 1.0 / 16
    problem = MPCProblem(
        transition_state_matrix=np.array([[1.0, T, T**2 / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]]),
        transition_input_matrix=np.array([T**3 / 6.0, T**2 / 2.0, T]).reshape((3, 1)),
        ineq_state_matrix=np.array([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0]]), ineq_input_matrix=None,
        ineq_vector=np.array([3.0, 3.0]), nb_timesteps=16, terminal_cost_weight=1.0,
        stage_state_cost_weight=None, stage_input_cost_weight=1e-6,
        initial_state=np.array([0.0, 0.0, 0.0]), goal_state=np.array([1.0, 0.0, 0.0]))
    for _ in range(20):
        plan = solve_mpc(problem, "b200")
    t0 = time.perf_counter()
    for i in range(calls):
        problem.update_initial_state(np.array([0.001 * i, 0.0, 0.0]))
        plan = solve_mpc(problem, "b200")
    us = (time.perf_counter() - t0) / calls * 1e6
    assert not plan.is_empty
    return {"solve_mpc_us": us, "calls": calls,
            "path": "qpmpc_b200.solve_mpc -> page-locked staging block -> qpmpc_b200_solve_host (zero-copy, one launch)",
            "workload": "triple_integrator N=16, one instance (BASELINE configs[0]), new initial state every call"}


def run_b200(args, rank, local_rank, world):
    import ctypes

    import torch

    from qpmpc_b200 import _capi, factor_model, lipm_walking_closed_loop, pendulum_closed_loop, solve_mpc_batch
    from qpmpc_b200.workloads import algorithmic_bytes_per_solve, to_batched

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    numa_node = None
    if world > 1:
        import torch.distributed as dist

        numa_node = bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=dev)

    B, N = args.batch, args.horizon
    n = N  # nu = 1
    tdtype = torch.float32 if is_f32(args) else torch.float64
    es = 4 if is_f32(args) else 8
    w0 = make_workload(args, 1000 * rank)
    bytes_per_solve = algorithmic_bytes_per_solve(w0, es)
    in_bytes = (bytes_per_solve - es * n - 4) * B
    loop_cfg = args.config in (3, 6)
    rotate = 1 if loop_cfg else max(2, min(16, int(np.ceil(ROTATE_BYTES / max(in_bytes, 1)))))
    sets = [w0] + [make_workload(args, 1000 * rank + s) for s in range(1, rotate)]
    problems = [to_batched(w, dtype=tdtype, device=dev) for w in sets]
    U_out = torch.empty((B, n), dtype=tdtype, device=dev)
    U_all = torch.empty((world * B, n), dtype=tdtype, device=dev) if world > 1 else None
    mkw = {} if args.method == "active_set" else {"method": args.method}
    gather, gather_kind = None, "none"
    if world > 1 and not loop_cfg:
        gather_kind = "nccl all_gather_into_tensor"
        # (the fused gather exists for the active-set kernel only)
        if os.environ.get("QPMPC_B200_GATHER", "peer") == "peer" and args.method == "active_set" \
                and tdtype == torch.float64:
            try:
                from qpmpc_b200.distributed import PeerGather

                gather = PeerGather(B, n)
                gather_kind = "peer stores from the solve kernel (fused) + barrier"
            except Exception as exc:  # noqa: BLE001
                gather_kind += f" (symmetric memory unavailable: {type(exc).__name__})"

    x0_init = torch.as_tensor(w0["x0"]).to(dev) if loop_cfg else None
    v_dev = torch.as_tensor(w0["v_target"]).to(dev) if args.config == 3 else None
    walk0 = {k: torch.as_tensor(w0[k]).to(dev) for k in ("support_foot", "strides", "phase_index", "stride_index")} \
        if args.config == 6 else None
    loop_info = {}
    # config 3: the model (A, B, D, weights) is shared by the batch and constant over the loop --
    # factored once (qpmpc_b200_factor), every cycle then only rebuilds q and h.
    # QPMPC_B200_BENCH_FACTORED=0 re-condenses every instance every cycle instead.
    model = None
    if loop_cfg and os.environ.get("QPMPC_B200_BENCH_FACTORED", "1") != "0":
        model = factor_model(problems[0])

    def walk(prob, x0_src, phase_src):
        prob.x0.copy_(x0_src, non_blocking=True)
        return lipm_walking_closed_loop(prob, phase_src["support_foot"].clone(), phase_src["strides"],
                                        phase_src["phase_index"].clone(), phase_src["stride_index"].clone(),
                                        WALK_CYCLES, factored=model if model is not None else False)

    def step(i):
        if args.config == 6:
            plan, _, unsolved, _ = walk(problems[0], x0_init, walk0)
            loop_info["unsolved"] = unsolved
            return plan
        if args.config == 3:
            problems[0].x0.copy_(x0_init)
            plan, _, unsolved, stats = pendulum_closed_loop(problems[0], v_dev, CYCLES, factored=model if model is not None else False, stats=True)
            loop_info["unsolved"], loop_info["stats"] = unsolved, stats
            return plan
        if gather is not None:
            U, status, iters = gather.solve(problems[i % rotate])
            return _StepResult(status[rank * B:(rank + 1) * B], iters, U)
        plan = solve_mpc_batch(problems[i % rotate], out=U_out, **mkw)
        if world > 1:
            dist.all_gather_into_tensor(U_all, U_out)
            plan.gathered = U_all
        return plan

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler comes up first (nvidia-smi takes a moment), so that nothing idles the GPU
    # between the W warm-up steps and the K timed ones
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    for i in range(args.warmup):
        plan = step(i)
    barrier()
    if not loop_cfg:
        assert int((plan.status != 0).sum().item()) == 0, "warm-up batch has unsolved instances"
    launches0 = _capi.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        plan = step(i)
        ev[i + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    launches = _capi.launch_count() - launches0
    per_step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    iters_mean = float(plan.iters.float().mean().item())
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- N > 1: the gathered U equals a local re-solve of every other rank's shard ----------
    gather_check = None
    if world > 1 and not loop_cfg:
        last = step(0)
        barrier()
        got = (last.gathered if hasattr(last, "gathered") else U_all).clone()
        worst = 0.0
        for r2 in range(world):
            w2 = make_workload(args, 1000 * r2)
            p2 = solve_mpc_batch(to_batched(w2, dtype=tdtype, device=dev), **mkw)
            torch.cuda.synchronize()
            diff = (got[r2 * B:(r2 + 1) * B] - p2.inputs.reshape(B, -1)).abs().max().item()
            worst = max(worst, float(diff))
        flag = torch.tensor([worst], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        gather_check = {"max_abs_diff_vs_local_resolve": float(flag.item()), "ranks_checked": world,
                        "rows_per_rank": B}
        assert gather_check["max_abs_diff_vs_local_resolve"] == 0.0, gather_check

    # ---- kernel-only duration (no collective): events around bare launches -----------------
    kernel_ms = None
    if not loop_cfg:
        ksteps = min(args.steps, 256)
        kev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * ksteps)]
        torch.cuda.synchronize()
        for i in range(ksteps):
            kev[2 * i].record()
            solve_mpc_batch(problems[i % rotate], out=U_out, **mkw)
            kev[2 * i + 1].record()
        torch.cuda.synchronize()
        kernel_ms = float(np.mean([kev[2 * i].elapsed_time(kev[2 * i + 1]) for i in range(ksteps)]))
    else:
        # the solve kernel's share of a cycle: the same problem solved outside the loop
        kev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        kev[0].record()
        for _ in range(50):
            solve_mpc_batch(problems[0], out=U_out, factored=model)
        kev[1].record()
        torch.cuda.synchronize()
        kernel_ms = kev[0].elapsed_time(kev[1]) / 50
    # units one launch of the dominant kernel processes: the batch -- or, when the whole closed
    # loop runs inside ONE launch of the shared-model kernel, the batch times the cycles (the
    # launch IS the step then)
    units_per_launch = B
    fused_loop = loop_cfg and model is not None and os.environ.get("QPMPC_B200_LOOP_FUSED", "1") != "0"
    if fused_loop:
        units_per_launch = solves_per_step(args)
        kernel_ms = total_ms / args.steps

    # ---- e2e: host buffers through the C ABI ----------------------------------------------
    lib = _capi.load()
    e2e = None
    if not loop_cfg:
        np_dtype = np.float32 if is_f32(args) else np.float64
        names = [k for k in ("A", "B", "C", "D", "e", "x0", "goal", "targets") if sets[0][k] is not None]
        host_sets = []
        for w in sets[:min(4, rotate)]:
            host_sets.append({k: torch.from_numpy(np.ascontiguousarray(w[k], dtype=np_dtype)).pin_memory()
                              for k in names})
        # one host array for the whole job: rank r's rows land in [r*B, (r+1)*B) -- the gather
        shm_path = f"/dev/shm/qpmpc_b200_bench_{os.environ.get('MASTER_PORT', '0')}_{world}"
        if world > 1:
            U_job = torch.from_file(shm_path + "_U", shared=True, size=world * B * n, dtype=tdtype)
            st_job = torch.from_file(shm_path + "_st", shared=True, size=world * B, dtype=torch.int32)
            # first touch: every rank faults its own rows in (on its NUMA node) before anybody pins
            U_job[rank * B * n:(rank + 1) * B * n].zero_()
            st_job[rank * B:(rank + 1) * B].zero_()
            dist.barrier()
            cudart = torch.cuda.cudart()
            for t in (U_job, st_job):
                rc = cudart.cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
                assert int(rc) == 0, rc
        else:
            U_job = torch.empty(B * n, dtype=tdtype).pin_memory()
            st_job = torch.empty(B, dtype=torch.int32).pin_memory()
        U_host = U_job[rank * B * n:(rank + 1) * B * n]
        st_host = st_job[rank * B:(rank + 1) * B]
        desc = problems[0].desc() if args.method == "active_set" else problems[0].desc(_capi.PDIP)
        h2d = sum(host_sets[0][k].numel() * es for k in names)
        d2h = U_host.numel() * es + st_host.numel() * 4
        vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

        def e2e_step(i):
            hs = host_sets[i % len(host_sets)]
            ops = _capi.Operands(*[vp(hs[k]) if k in hs else None
                                   for k in ("A", "B", "C", "D", "e", "x0", "goal", "targets")])
            outs = _capi.Outputs(vp(U_host), vp(st_host), None, None)
            rc = lib.qpmpc_b200_solve_host(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs), local_rank)
            assert rc == 0, rc
            if world > 1:
                dist.barrier()  # every rank's rows are in the shared host array: the gather is complete

        e2e_steps = max(4, min(args.steps, 64))
        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        assert int((st_job != 0).sum()) == 0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
            # the job-wide host array holds every rank's rows: compare with the device gather
            # (last step of the e2e loop used host set (e2e_steps-1) % 4 = device set of the same index)
            chk = step((e2e_steps - 1) % len(host_sets))
            barrier()
            ref_rows = (chk.gathered if hasattr(chk, "gathered") else U_all).cpu().reshape(-1)
            assert torch.equal(ref_rows, U_job), "shared host array differs from the device gather"
            dist.barrier()
            cudart = torch.cuda.cudart()
            for t in (U_job, st_job):
                cudart.cudaHostUnregister(t.data_ptr())
            if rank == 0:
                for suffix in ("_U", "_st"):
                    try:
                        os.unlink(shm_path + suffix)
                    except OSError:
                        pass
        e2e = {"value": world * B * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "call": "qpmpc_b200_solve_host (pinned host buffers, zero-copy: one kernel launch whose bulk-TMA staging "
                       "reads the operands from host memory over PCIe and whose epilogue stores U / status into host "
                       "memory; sync per step)" + ("; ranks write their rows into one host array shared by the "
                                                   "job (the gather), barrier per step" if world > 1 else "")}
    elif args.config == 6:
        # the walking loop: initial states and the phase machine's state come from pinned host
        # memory every step, the final states go back
        x0_host = torch.from_numpy(np.ascontiguousarray(w0["x0"])).pin_memory()
        ph_host = {k: torch.from_numpy(np.ascontiguousarray(w0[k])).pin_memory()
                   for k in ("support_foot", "strides", "phase_index", "stride_index")}
        xf_host = torch.empty_like(x0_host).pin_memory()
        e2e_steps = max(2, min(args.steps, 8))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ph_dev = {k: v.to(dev, non_blocking=True) for k, v in ph_host.items()}
            walk(problems[0], x0_host, ph_dev)
            xf_host.copy_(problems[0].x0, non_blocking=True)
            torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": world * solves_per_step(args) * e2e_steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": x0_host.numel() * 8 + sum(v.numel() * v.element_size() for v in ph_host.values()),
               "d2h_bytes_per_step": xf_host.numel() * 8, "steps": e2e_steps,
               "call": "lipm_walking_closed_loop -> qpmpc_b200_lipm_closed_loop: initial states, strides, support feet "
                       "and phases from pinned host memory, final states read back; sync per step"}
    else:
        # config 3: the state lives on the device for the whole loop; a step uploads the initial
        # states / target velocities from pinned host memory and reads the final states back
        x0_host = torch.from_numpy(np.ascontiguousarray(w0["x0"])).pin_memory()
        v_host = torch.from_numpy(np.ascontiguousarray(w0["v_target"])).pin_memory()
        xf_host = torch.empty_like(x0_host).pin_memory()
        vd = torch.empty_like(v_dev)
        e2e_steps = max(2, min(args.steps, 8))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            problems[0].x0.copy_(x0_host, non_blocking=True)
            vd.copy_(v_host, non_blocking=True)
            pendulum_closed_loop(problems[0], vd, CYCLES, factored=model if model is not None else False)
            xf_host.copy_(problems[0].x0, non_blocking=True)
            torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": world * solves_per_step(args) * e2e_steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": x0_host.numel() * 8 + v_host.numel() * 8,
               "d2h_bytes_per_step": xf_host.numel() * 8, "steps": e2e_steps,
               "call": "pendulum_closed_loop -> qpmpc_b200_pendulum_closed_loop: initial states and target "
                       "velocities from pinned host memory, final states read back; sync per step"}
    clocks = sampler.stop()

    # ---- roofline denominators ----------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            with open(peaks_path) as f:
                hbm_peak, peak_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    achieved_gbs = bytes_per_solve * units_per_launch / (kernel_ms * 1e-3) / 1e9
    tf = ctypes.c_double(0.0)
    lib.qpmpc_b200_fp64_peak(local_rank, ctypes.byref(tf))
    f_survey, f_exec = flop_models(args, iters_mean)
    tfl = lambda f: f * units_per_launch / (kernel_ms * 1e-3) / 1e12  # noqa: E731

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * solves_per_step(args) * args.steps / (total_ms * 1e-3)
    prof = measured_profile(args)
    traffic, traffic_src = prof["traffic"], prof["source"]
    if traffic is not None and prof.get("profile_batch") not in (None, units_per_launch):
        traffic = traffic * units_per_launch / prof["profile_batch"]  # the summary's launch had another batch
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype_of(args), "data": "synthetic",
        "config": config_dict(args, world,
                              f"inputs rotate over {rotate} distinct sets ({rotate}x{in_bytes / 1e6:.1f} MB > L2)"
                              if rotate > 1 else "state resident on the device across the cycles of the loop (the workload)"),
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "kernel": kernel_name(args), "kernel_ms": kernel_ms,
                     "bytes_per_solve": bytes_per_solve, "units_per_launch": units_per_launch,
                     "note": ("latency/FP64-issue bound by design (SURVEY 8d): HBM fraction is necessarily tiny; see fp64"
                              + ("; the launch is the whole closed loop: `achieved` counts the algorithmic bytes of every "
                                 "cycle, `traffic` is what the launch really moved -- the state never leaves the SM "
                                 "between cycles" if fused_loop else ""))},
        "iters_mean": iters_mean, "gather": gather_kind, "numa_node": numa_node,
        "step_ms_min_max": [min(per_step_ms), max(per_step_ms)],
        "clocks": clocks,
    }
    if not is_f32(args):
        line["fp64"] = {
            "peak_tflops": tf.value,
            "peak_source": "qpmpc_b200_fp64_peak: 16 independent DFMA chains per thread, 8 CTAs of 256 threads "
                           "per SM, 4096 iterations, best of 4 timed launches (CUDA events); not in "
                           "MEASURED_PEAKS.json",
            "survey_F": {"flops_per_solve": f_survey, "achieved_tflops": tfl(f_survey),
                         "frac": tfl(f_survey) / tf.value if tf.value > 0 else None,
                         "model": "SURVEY 8(d): dense structure-agnostic condensing + 10 interior-point iterations"},
            "executed": {"flops_per_solve": f_exec, "achieved_tflops": tfl(f_exec),
                         "frac": tfl(f_exec) / tf.value if tf.value > 0 else None,
                         "model": "flops the kernel's algorithm needs at the measured mean iteration count "
                                  "(paired rows, structure exploited; a lower bound of what is executed)"},
        }
        if args.method != "active_set":
            del line["fp64"]["executed"]  # (the model above is the active-set algorithm's)
        if prof["flops_per_solve"]:
            fm = prof["flops_per_solve"]
            line["fp64"]["executed_ncu"] = {
                "flops_per_solve": fm, "achieved_tflops": tfl(fm), "frac": tfl(fm) / tf.value if tf.value > 0 else None,
                "model": "DFMA/DMUL/DADD warp instructions the profiler counted for one launch x 32 lanes "
                         f"(profiles/{prof['source']}), at this run's kernel time"}
    if gather_check is not None:
        line["gather_check"] = gather_check
    if args.config == 6:
        line["closed_loop"] = {"cycles": WALK_CYCLES, "unsolved": int(loop_info["unsolved"].item()),
                               "model": "factored once (qpmpc_b200_factor), q and h per cycle" if model is not None
                               else "condensed and factored per instance and cycle",
                               "launches_per_step": launches / args.steps,
                               "loop": loop_kind(model)}
    if args.config == 3:
        hist = loop_info["stats"]["iterations"].cpu().numpy().astype(float) / B
        useful = int(loop_info["stats"]["upright"].item())
        line["closed_loop"] = {
            "cycles": CYCLES, "unsolved": int(loop_info["unsolved"].item()),
            "upright_frac_final": float((problems[0].x0[:, 1].abs() < 1.2).float().mean().item()),
            "useful_solves_per_step": useful, "useful_frac": useful / (B * CYCLES),
            "value_useful_only": value * useful / (B * CYCLES),
            "note": "useful = (instance, cycle) pairs solved from a state with |pitch| <= 1.2 rad; `value` "
                    "counts every solve, value_useful_only only those",
            "mean_iterations_per_cycle": {"cycle_0": hist[0], "cycle_1": hist[1], "cycle_2": hist[2],
                                          "cycles_3_9": float(hist[3:10].mean()), "cycles_10_49": float(hist[10:50].mean()),
                                          "cycles_50_199": float(hist[50:].mean())},
            "model": "factored once (qpmpc_b200_factor), q and h per cycle" if model is not None
                     else "condensed and factored per instance and cycle",
            "launches_per_step": launches / args.steps, "loop": loop_kind(model)}
    if not args.no_cpu_baseline:
        v, threads, sample = cpu_arm(args, sets[0], args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                                "reference_python": reference_python_condense()}
    if world == 1 and args.config == 2 and args.method == "active_set":
        line["single_call"] = single_call_latency()
    if world == 1 and args.config == 4 and args.method == "active_set":
        # the humanoid instances share one model (A, B, C; only e_k, x0 and the goal differ): the same
        # batches with the model factored ONCE outside the timed region (qpmpc_b200_factor), next to
        # the headline, which condenses and factors every instance
        models = [factor_model(pr) for pr in problems]
        for i in range(8):
            solve_mpc_batch(problems[i % rotate], out=U_out, factored=models[i % rotate])
        torch.cuda.synchronize()
        fe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = max(64, min(args.steps, 1024))
        fe[0].record()
        for i in range(reps):
            plan_f = solve_mpc_batch(problems[i % rotate], out=U_out, factored=models[i % rotate])
        fe[1].record()
        torch.cuda.synchronize()
        ms_f = fe[0].elapsed_time(fe[1]) / reps
        line["shared_model_factored"] = {
            "value": B / (ms_f * 1e-3), "unit": UNIT, "ms_per_step": ms_f, "steps": reps,
            "unsolved": int((plan_f.status != 0).sum().item()),
            "note": "same rotating batches through qpmpc_b200_solve_factored; the record of the shared model "
                    "is computed once per batch object outside the timed region"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _StepResult:
    def __init__(self, status, iters, gathered=None):
        self.status, self.iters, self.gathered = status, iters, gathered


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
