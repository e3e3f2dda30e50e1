#!/usr/bin/env python3
"""BASELINE config 3: wheeled inverted pendulum, batch 16384, 200-cycle
receding horizon on one GPU (state resident, 2 launches per cycle).

    python tools/closed_loop.py [--batch 16384] [--cycles 200] [--ltv]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from qpmpc_b200 import pendulum_closed_loop
from qpmpc_b200.workloads import pendulum_batch, to_batched

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--cycles", type=int, default=200)
ap.add_argument("--ltv", action="store_true", help="pass A, B as per-step stacks (LTV code path)")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--graph", action="store_true", help="capture the 2*cycles+1 launches in one CUDA graph")
a = ap.parse_args()
w = pendulum_batch(a.batch, seed=1, ltv_model=a.ltv)
x0 = torch.as_tensor(w["x0"]).cuda()
v_dev = torch.as_tensor(w["v_target"]).cuda()  # on the device already: the loop is then capturable
prob = to_batched(w)
times = []
graph = None
if a.graph:
    # warm-up outside capture (module loading, function attributes), then capture once
    pendulum_closed_loop(prob, v_dev, 2)
    torch.cuda.synchronize()
    prob.x0.copy_(x0)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        plan, traj, unsolved = pendulum_closed_loop(prob, v_dev, a.cycles)
for rep in range(a.reps + 1):
    prob.x0.copy_(x0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        plan, traj, unsolved = pendulum_closed_loop(prob, v_dev, a.cycles)
    e1.record()
    torch.cuda.synchronize()
    if rep:
        times.append(e0.elapsed_time(e1))
ms = sum(times) / len(times)
print(json.dumps({"workload": "wheeled_inverted_pendulum closed loop" + (" (LTV stacks)" if a.ltv else "")
                  + (" [CUDA graph]" if a.graph else ""),
                  "batch": a.batch, "cycles": a.cycles, "ms": ms, "ms_per_cycle": ms / a.cycles,
                  "solves_per_s": a.batch * a.cycles / ms * 1e3, "unsolved": int(unsolved.item()),
                  "iters_mean_last_cycle": float(plan.iters.float().mean()),
                  "max_abs_pitch_final": float(prob.x0[:, 1].abs().max())}))
