#!/usr/bin/env python3
"""bench.py -- MPC solves/sec (batched), the BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): triple integrator (nx=3, nu=1, nc=2, N=16),
fp64, 65 536 instances per GPU, per-instance A, B, C, e, x0, goal records
(seed 0).  A "step" condenses and solves the whole batch once.
Weak scaling: every rank owns its own 65 536 instances; for N > 1 every rank
ends each step with the stacked U of ALL ranks -- the all-gather the north star
names, fused into the solve kernel (its epilogue stores each row into every
rank's symmetric buffer over NVLink; NCCL all-gather is the fallback).

`value`     device-resident throughput (inputs already in HBM), CUDA events.
`e2e`       the same batch through qpmpc_b200_solve_host with pinned HOST
            buffers: H2D of every operand + kernel + D2H of U/status per step
            (pipelined over three streams in chunks, one sync per step).
`roofline`  algorithmic bytes per launch / measured kernel time vs the HBM
            peak -- by construction tiny: the path is FP64-issue bound, so the
            FP64 fraction is reported next to it (`fp64`).
`cpu_baseline`  the C oracle (oracle/mpc_oracle.c: condensing + Goldfarb-Idnani,
            OpenMP) on the same instances, all host threads.

`--impl reference` times that CPU port alone (the reference itself is pure
Python over qpsolvers wheels that cannot be installed offline).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BATCH_PER_GPU = 65536
HORIZON = 16
ROTATE = 12  # distinct input sets: 12 x 13.6 MB = 164 MB > 126 MB of L2
METRIC = "MPC solves/sec (batched)"
UNIT = "solves/s"
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--horizon", type=int, default=HORIZON)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--method", default="active_set", choices=["active_set", "pdip"],
                    help="solver kernel; the headline and the default is the exact active set "
                         "(pdip: experimental, needs QPMPC_B200_ENABLE_PDIP=1, DESIGN.md 2b)")
    return ap.parse_args()


def config_dict(args, world):
    return {
        "workload": f"triple_integrator fp64 nx=3 nu=1 nc=2 N={args.horizon} "
                    f"batch={args.batch}/GPU per-instance A,B,C,e,x0,goal (BASELINE configs[1])",
        "batch_per_gpu": args.batch, "global_batch": args.batch * world,
        "horizon": args.horizon,
        "method": "dual active set (Goldfarb-Idnani), exact" if args.method == "active_set"
        else "interior point (Mehrotra) + primal-dual active-set polish, tol 1e-9",
        "l2": f"inputs rotate over {ROTATE} distinct sets ({ROTATE}x{args.batch * 208 / 1e6:.1f} MB > L2)",
        "parallelism": f"batch-sharded x{world}, U gathered on every rank each step" if world > 1 else "single GPU",
    }


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
def cpu_arm(workload, seconds, min_reps=1):
    """Time oracle.solve_batch (C, OpenMP, all threads) on the workload for
    about `seconds`; returns (solves/s, threads, sample description)."""
    import oracle
    from qpmpc_b200.workloads import oracle_ops

    threads = oracle.num_threads()
    ops = oracle_ops(workload)
    B = workload["batch"]
    args = (B, workload["N"], workload["nx"], workload["nu"], workload["nc"], ops,
            workload["w_t"], workload["w_x"], workload["w_u"])
    oracle.solve_batch(*args)  # warm-up (page in, thread pool)
    reps, t0 = 0, time.perf_counter()
    while True:
        oracle.solve_batch(*args)
        reps += 1
        dt = time.perf_counter() - t0
        if reps >= min_reps and dt >= seconds:
            break
    return B * reps / dt, threads, f"{reps} x {B} instances of the bench workload in {dt:.1f} s"


class _StepResult:
    def __init__(self, status, iters):
        self.status, self.iters = status, iters


def run_reference(args, rank, world):
    if rank != 0:
        return
    from qpmpc_b200.workloads import triple_integrator_batch

    import oracle

    w = triple_integrator_batch(args.batch, N=args.horizon, seed=0)
    ops_args = None
    from qpmpc_b200.workloads import oracle_ops

    ops_args = (args.batch, w["N"], 3, 1, 2, oracle_ops(w), w["w_t"], w["w_x"], w["w_u"])
    threads = oracle.num_threads()
    for _ in range(max(args.warmup, 1)):
        oracle.solve_batch(*ops_args)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.solve_batch(*ops_args)
    dt = time.perf_counter() - t0
    value = args.batch * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x {args.batch} instances, C oracle "
                                   "(condense + Goldfarb-Idnani), OpenMP"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python over qpsolvers/proxqp wheels that are not installable "
                "offline; this arm is the C port of its algorithm (oracle/), on all host threads",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import ctypes

    import torch

    from qpmpc_b200 import _capi, solve_mpc_batch
    from qpmpc_b200.workloads import (algorithmic_bytes_per_solve, to_batched,
                                      triple_integrator_batch)

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    B, N = args.batch, args.horizon
    n = N  # nu = 1
    sets = [triple_integrator_batch(B, N=N, seed=1000 * rank + s) for s in range(ROTATE)]
    problems = [to_batched(w, device=dev) for w in sets]
    U_out = torch.empty((B, n), dtype=torch.float64, device=dev)
    U_all = torch.empty((world * B, n), dtype=torch.float64, device=dev) if world > 1 else None
    # N > 1: the kernel's epilogue stores every U row into all ranks' buffers
    # over NVLink (fused gather); NCCL all-gather only if symmetric memory is
    # unavailable on the box.
    mkw = {} if args.method == "active_set" else {"method": args.method}
    gather, gather_kind = None, "none"
    if world > 1:
        gather_kind = "nccl all_gather_into_tensor"
        # (the fused gather exists for the active-set kernel only)
        if os.environ.get("QPMPC_B200_GATHER", "peer") == "peer" and args.method == "active_set":
            try:
                from qpmpc_b200.distributed import PeerGather

                gather = PeerGather(B, n)
                gather_kind = "peer stores from the solve kernel (fused) + barrier"
            except Exception as exc:  # noqa: BLE001
                gather_kind += f" (symmetric memory unavailable: {type(exc).__name__})"

    def step(i):
        if gather is not None:
            _, status, iters = gather.solve(problems[i % ROTATE])
            return _StepResult(status[rank * B:(rank + 1) * B], iters)
        plan = solve_mpc_batch(problems[i % ROTATE], out=U_out, **mkw)
        if world > 1:
            dist.all_gather_into_tensor(U_all, U_out)
        return plan

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        plan = step(i)
    barrier()
    assert int((plan.status != 0).sum().item()) == 0, "warm-up batch has unsolved instances"

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
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

    # kernel-only duration (no collective): events around bare launches
    kev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    torch.cuda.synchronize()
    for i in range(args.steps):
        kev[2 * i].record()
        solve_mpc_batch(problems[i % ROTATE], out=U_out, **mkw)
        kev[2 * i + 1].record()
    torch.cuda.synchronize()
    kernel_ms = float(np.mean([kev[2 * i].elapsed_time(kev[2 * i + 1]) for i in range(args.steps)]))

    # ---- e2e: host buffers through the C ABI -----------------------------
    lib = _capi.load()
    host_sets = []
    for w in sets[:4]:
        hs = {k: torch.from_numpy(np.ascontiguousarray(w[k])).pin_memory()
              for k in ("A", "B", "C", "e", "x0", "goal")}
        host_sets.append(hs)
    U_host = torch.empty((B, n), dtype=torch.float64).pin_memory()
    st_host = torch.empty(B, dtype=torch.int32).pin_memory()
    desc = problems[0].desc() if args.method == "active_set" else problems[0].desc(_capi.PDIP)
    h2d = sum(t.numel() * 8 for t in host_sets[0].values())
    d2h = U_host.numel() * 8 + st_host.numel() * 4
    vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

    def e2e_step(i):
        hs = host_sets[i % len(host_sets)]
        ops = _capi.Operands(vp(hs["A"]), vp(hs["B"]), vp(hs["C"]), None, vp(hs["e"]),
                             vp(hs["x0"]), vp(hs["goal"]), None)
        outs = _capi.Outputs(vp(U_host), vp(st_host), None, None)
        rc = lib.qpmpc_b200_solve_host(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs),
                                       local_rank)
        assert rc == 0, rc

    e2e_steps = max(4, min(args.steps, 64))
    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert int((st_host != 0).sum()) == 0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop()

    # ---- roofline denominators ------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    if os.path.exists(peaks_path):
        try:
            with open(peaks_path) as f:
                hbm_peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
    bytes_per_solve = algorithmic_bytes_per_solve(sets[0])
    achieved_gbs = bytes_per_solve * B / (kernel_ms * 1e-3) / 1e9
    tf = ctypes.c_double(0.0)
    lib.qpmpc_b200_fp64_peak(local_rank, ctypes.byref(tf))
    # flops the kernel's algorithm needs per solve: condensing
    # recursions, Cholesky, forward substitutions for t and the m rows of M,
    # the final two triangular solves, and per active-set iteration one
    # matrix-vector product with M plus one Householder update of M.
    nn, mm = n, 2 * N
    f_setup = N * (4 * 9 + 2 * 2 * 3) * nn + nn**3 / 3 + (mm + 1) * nn * nn + 2 * nn * nn
    f_iter = 6 * mm * nn
    flops_per_solve = f_setup + iters_mean * f_iter
    achieved_tf = flops_per_solve * B / (kernel_ms * 1e-3) / 1e12

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * B * args.steps / (total_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world),
        "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "call": "qpmpc_b200_solve_host (pinned host buffers; H2D, kernel, D2H pipelined over "
                        "3 streams in 16384-instance chunks; sync per step)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak,
                     # dram__bytes_read + write of one launch, ncu --set full (profiles/r01r_solve_ti16_f64.txt)
                     "traffic": 13774336 if (B, N) == (65536, 16) else None, "peak_source": peak_src,
                     "kernel": "mpc_solve_kernel<double,NP=%d,MR=2>" % (8 if n <= 8 else 16 if n <= 16 else 32)
                     if n <= 32 else "mpc_solve_cta_kernel<double>", "kernel_ms": kernel_ms,
                     "bytes_per_solve": bytes_per_solve,
                     "note": "latency/FP64-issue bound by design (SURVEY 8d): HBM fraction is "
                             "necessarily tiny; see fp64"},
        "fp64": {"achieved_tflops": achieved_tf, "peak_tflops": tf.value,
                 "frac": achieved_tf / tf.value if tf.value > 0 else None,
                 "flops_per_solve": flops_per_solve, "peak_source": "qpmpc_b200_fp64_peak (DFMA probe)"},
        "iters_mean": iters_mean, "gather": gather_kind,
        "step_ms_min_max": [min(per_step_ms), max(per_step_ms)],
        "clocks": clocks,
    }
    if args.method != "active_set":
        # the flop model and the ncu traffic figure above belong to the active-set kernel
        line["roofline"]["kernel"] = line["roofline"]["kernel"].replace("mpc_solve_kernel", "mpc_pdip_kernel")
        line["roofline"]["traffic"] = None
        line["fp64"] = None
    if not args.no_cpu_baseline:
        v, threads, sample = cpu_arm(sets[0], args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
