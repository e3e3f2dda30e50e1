#!/usr/bin/env python3
"""Kernel time of config 2 against an iteration cap: separates set-up
(condense + factor + final solve) from the per-iteration cost."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from qpmpc_b200 import solve_mpc_batch
from qpmpc_b200.workloads import to_batched, triple_integrator_batch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prob = to_batched(triple_integrator_batch(65536, N=N, seed=0))
for cap in (1, 2, 3, 4, 6, 8, 12, 16, 64):
    for _ in range(3):
        plan = solve_mpc_batch(prob, max_iter=cap)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan = solve_mpc_batch(prob, max_iter=cap)
    e1.record()
    torch.cuda.synchronize()
    it = plan.iters.float().clamp(max=cap)
    print(f"cap {cap:3d}: {e0.elapsed_time(e1) / 10:.3f} ms  mean executed iterations {it.mean().item():.2f}  "
          f"unsolved {int((plan.status != 0).sum())}", flush=True)
