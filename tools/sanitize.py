#!/usr/bin/env python3
"""Small workload mix for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitize.py

Covers the warp kernel (paired and unpaired rows, Toeplitz and dense-G paths, register and
shared-memory M variants), the long-horizon kernel (one and two warps per instance), the CTA
kernel, the shared-model path (factor + factored solve), the interior point, the condense
kernels, the tcgen05 Hessian kernel, the closed loops and the zero-copy host entry."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ctypes

import numpy as np

from qpmpc_b200 import (_capi, condense_batch, factor_model, lipm_walking_closed_loop, pendulum_closed_loop,
                        solve_mpc_batch)
from qpmpc_b200.workloads import (humanoid_batch, lipm_walking_batch, pendulum_batch, random_batch, to_batched,
                                  triple_integrator_batch)

cases = [
    triple_integrator_batch(37, N=16, seed=1),
    triple_integrator_batch(21, N=8, seed=2),
    triple_integrator_batch(9, N=32, seed=3),
    triple_integrator_batch(3, N=64, seed=4),
    triple_integrator_batch(5, N=48, seed=7),
    triple_integrator_batch(7, N=24, seed=8),
    triple_integrator_batch(3, N=96, seed=9),   # CTA kernel out of its global-memory workspace
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
plan, traj, unsolved = pendulum_closed_loop(to_batched(w), w["v_target"], 5, record=True, factored=False)
torch.cuda.synchronize()
print("closed loop ok, unsolved", int(unsolved.item()))

# unpaired variants, interior point, shared-model path
os.environ["QPMPC_B200_NO_PAIRED"] = "1"
for w in (triple_integrator_batch(21, N=16, seed=11), triple_integrator_batch(5, N=32, seed=12), pendulum_batch(9)):
    plan = solve_mpc_batch(to_batched(w))
    torch.cuda.synchronize()
    print("unpaired", w["name"], "unsolved", int((plan.status != 0).sum()), flush=True)
del os.environ["QPMPC_B200_NO_PAIRED"]
for w in (triple_integrator_batch(33, seed=13), pendulum_batch(17, seed=14)):
    plan = solve_mpc_batch(to_batched(w), method="pdip")
    torch.cuda.synchronize()
    print("pdip", w["name"], "unsolved", int((plan.status != 0).sum()), flush=True)
for w in (pendulum_batch(37, seed=15), humanoid_batch(21, seed=16), triple_integrator_batch(9, N=32, seed=17, per_instance_model=False)):
    prob = to_batched(w)
    plan = solve_mpc_batch(prob, factored=factor_model(prob), return_multipliers=True)
    torch.cuda.synchronize()
    print("factored", w["name"], "unsolved", int((plan.status != 0).sum()), flush=True)
w = pendulum_batch(21, seed=20)   # fused loop (one launch), lane groups of the last CTA without an instance
prob = to_batched(w)
from qpmpc_b200.workloads import pendulum_targets
tg, goal = pendulum_targets(w["x0"], w["v_target"], w["N"], w["T"])
prob.update_goal_state(goal)
prob.update_target_states(tg)
plan, traj, unsolved, stats = pendulum_closed_loop(prob, w["v_target"], 7, record=True, stats=True, factored=factor_model(prob))
torch.cuda.synchronize()
print("fused pendulum loop ok, unsolved", int(unsolved.item()))
w = lipm_walking_batch(24)
prob = to_batched(w)
plan, traj, unsolved, _ = lipm_walking_closed_loop(prob, w["support_foot"], w["strides"], w["phase_index"],
                                                   w["stride_index"], 6, record=True, factored=factor_model(prob))
torch.cuda.synchronize()
print("walking loop ok, unsolved", int(unsolved.item()))
# tcgen05 Hessian (fp32, n = 64, stage cost)
w = triple_integrator_batch(6, N=64, seed=18, per_instance_model=False)
w["w_x"], w["targets"] = 0.5, np.zeros((6, 64 * 3))
out = condense_batch(to_batched(w, dtype=torch.float32), ("P", "Psi", "psi_last"))
torch.cuda.synchronize()
print("tensor-core Hessian ok", float(out["P"].abs().max()))
# zero-copy host entry
w = triple_integrator_batch(300, seed=19)
prob = to_batched(w)
host = {k: torch.from_numpy(np.ascontiguousarray(w[k])).pin_memory() for k in ("A", "B", "C", "e", "x0", "goal")}
U, st = torch.zeros((300, 16), dtype=torch.float64).pin_memory(), torch.zeros(300, dtype=torch.int32).pin_memory()
vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
ops = _capi.Operands(vp(host["A"]), vp(host["B"]), vp(host["C"]), None, vp(host["e"]), vp(host["x0"]), vp(host["goal"]), None)
outs = _capi.Outputs(vp(U), vp(st), None, None)
desc = prob.desc()
rc = _capi.load().qpmpc_b200_solve_host(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs), 0)
print("zero-copy host entry rc", rc, "unsolved", int((st != 0).sum()))
