#!/usr/bin/env python3
"""BASELINE config 4: humanoid / LIPM with per-instance per-step ZMP bounds
e_k (LTV constraints), batch 8192, fp32 (and fp64 for reference), one GPU.
Parity of the fp32 run is stated against the fp64 CPU oracle."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from qpmpc_b200 import solve_mpc_batch
from qpmpc_b200.workloads import humanoid_batch, oracle_ops, to_batched

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
w = humanoid_batch(B)
ref = oracle.solve_batch(B, w["N"], 3, 1, 2, oracle_ops(w), w["w_t"], w["w_x"], w["w_u"])
for dt in (torch.float32, torch.float64):
    prob = to_batched(w, dtype=dt)
    for _ in range(3):
        plan = solve_mpc_batch(prob)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        plan = solve_mpc_batch(prob)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    U = plan.inputs.reshape(B, -1).double().cpu().numpy()
    st = plan.status.cpu().numpy()
    ok = (st == 0) & (ref["status"] == 0)
    scale = np.maximum(1.0, np.abs(ref["U"][ok]).max(axis=1))
    err = np.abs(U[ok] - ref["U"][ok]).max(axis=1)
    print(json.dumps({"workload": "humanoid LIPM, LTV e_k", "dtype": str(dt).split(".")[-1], "batch": B, "ms": ms,
                      "solves_per_s": B / ms * 1e3, "iters_mean": float(plan.iters.float().mean()),
                      "solved_frac": float(ok.mean()), "max_abs_err": float(err.max()),
                      "max_rel_err": float((err / scale).max())}))
