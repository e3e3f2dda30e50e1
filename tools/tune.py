#!/usr/bin/env python3
"""Kernel-only timing sweep on one GPU: batch / horizon / warps-per-CTA.

    python tools/tune.py [--wpc 1 2 4] [--N 16] [--batch 65536] [--dtype f64]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from qpmpc_b200 import solve_mpc_batch
from qpmpc_b200.workloads import (humanoid_batch, pendulum_batch, to_batched,
                                  triple_integrator_batch)


def timeit(fn, reps=10):
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


ap = argparse.ArgumentParser()
ap.add_argument("--wpc", type=int, nargs="*", default=[1, 2, 4])
ap.add_argument("--N", type=int, nargs="*", default=[16])
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--workload", default="ti")
a = ap.parse_args()
dt = torch.float64 if a.dtype == "f64" else torch.float32
for N in a.N:
    if a.workload == "ti":
        w = triple_integrator_batch(a.batch, N=N, seed=0)
    elif a.workload == "hum":
        w = humanoid_batch(a.batch)
    else:
        w = pendulum_batch(a.batch)
    prob = to_batched(w, dtype=dt)
    for wpc in a.wpc:
        os.environ["QPMPC_B200_WPC"] = str(wpc)
        plan = solve_mpc_batch(prob)
        ms = timeit(lambda: solve_mpc_batch(prob))
        it = plan.iters.float()
        print(f"{w['name']} {a.dtype} batch={a.batch} wpc={wpc}: {ms:.3f} ms  "
              f"{a.batch / ms * 1e3 / 1e6:.2f} M solves/s  iters mean {it.mean().item():.2f} "
              f"max {int(it.max().item())} unsolved {int((plan.status != 0).sum().item())}", flush=True)
