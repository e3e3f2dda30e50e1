#!/usr/bin/env python3
"""Config-2 shape with ONE model shared by the batch (the usual situation: one robot, many
initial / goal states): full path against the shared-model path (factor once, then per solve
only q, h, t and the violations).  Kernel-only, CUDA events.

    python tools/shared_model_bench.py [--N 16] [--batch 65536]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from qpmpc_b200 import factor_model, solve_mpc_batch
from qpmpc_b200.workloads import to_batched, triple_integrator_batch

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, nargs="*", default=[16, 8, 32])
ap.add_argument("--batch", type=int, default=65536)
a = ap.parse_args()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for N in a.N:
    w = triple_integrator_batch(a.batch, N=N, seed=0, per_instance_model=False)
    prob = to_batched(w)
    full = solve_mpc_batch(prob)
    model = factor_model(prob)
    fast = solve_mpc_batch(prob, factored=model)
    torch.cuda.synchronize()
    ok = (full.status == 0) & (fast.status == 0)
    err = float((full.inputs - fast.inputs)[ok].abs().max())
    same = bool(torch.equal(full.status, fast.status) and torch.equal(full.iters[ok], fast.iters[ok]))
    t_full = timeit(lambda: solve_mpc_batch(prob))
    t_fast = timeit(lambda: solve_mpc_batch(prob, factored=model))
    print(f"triple integrator N={N} batch={a.batch} shared model: full path {t_full:.3f} ms "
          f"({a.batch / t_full / 1e3:.1f} M solves/s), factored {t_fast:.3f} ms ({a.batch / t_fast / 1e3:.1f} M solves/s); "
          f"|dU| {err:.1e}, same status and iterations: {same}, mean iterations {float(full.iters.float().mean()):.2f}",
          flush=True)
