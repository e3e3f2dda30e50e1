#!/usr/bin/env python3
"""Small workload mix for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitize.py

Covers the warp kernel (Toeplitz and dense-G paths, register and shared-memory
M variants), the CTA kernel, the condense kernels and the closed loop."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from qpmpc_b200 import condense_batch, pendulum_closed_loop, solve_mpc_batch
from qpmpc_b200.workloads import (humanoid_batch, pendulum_batch, random_batch, to_batched,
                                  triple_integrator_batch)

cases = [
    triple_integrator_batch(37, N=16, seed=1),
    triple_integrator_batch(21, N=8, seed=2),
    triple_integrator_batch(9, N=32, seed=3),
    triple_integrator_batch(3, N=64, seed=4),
    humanoid_batch(19),
    pendulum_batch(23),
    pendulum_batch(11, ltv_model=True),
    random_batch(13, 6, 3, 2, 3, seed=5, ltv=True),
    random_batch(13, 5, 6, 1, 2, seed=6, ltv=False),
]
bad = 0
for w in cases:
    prob = to_batched(w)
    plan = solve_mpc_batch(prob, return_multipliers=True)
    condense_batch(prob, ("P", "q", "G", "h", "Phi", "Psi", "phi_last", "psi_last"))
    torch.cuda.synchronize()
    bad += int((plan.status != 0).sum())
    print(w["name"], w["batch"], "unsolved", int((plan.status != 0).sum()), flush=True)
w = pendulum_batch(16, seed=1)
plan, traj, unsolved = pendulum_closed_loop(to_batched(w), w["v_target"], 5, record=True)
torch.cuda.synchronize()
print("closed loop ok, unsolved", int(unsolved.item()))
