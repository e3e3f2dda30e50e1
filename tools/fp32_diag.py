#!/usr/bin/env python3
"""fp32 vs fp64-oracle error distribution on the humanoid workload (config 4)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from qpmpc_b200 import solve_mpc_batch
from qpmpc_b200.workloads import humanoid_batch, oracle_ops, to_batched

w = humanoid_batch(8192)
ref = oracle.solve_batch(w["batch"], w["N"], 3, 1, 2, oracle_ops(w), w["w_t"], w["w_x"], w["w_u"])
for dt in (torch.float64, torch.float32):
    plan = solve_mpc_batch(to_batched(w, dtype=dt))
    U = plan.inputs.reshape(w["batch"], -1).double().cpu().numpy()
    st = plan.status.cpu().numpy()
    ok = (st == 0) & (ref["status"] == 0)
    scale = np.maximum(1.0, np.abs(ref["U"][ok]).max(axis=1))
    err = np.abs(U[ok] - ref["U"][ok]).max(axis=1) / scale
    print(dt, "solved", ok.mean(), "status counts", np.bincount(st, minlength=4),
          "rel err median %.2e p99 %.2e p999 %.2e max %.2e" %
          (np.median(err), np.percentile(err, 99), np.percentile(err, 99.9), err.max()),
          "n>1e-3:", int((err > 1e-3).sum()), "iters", plan.iters.float().mean().item())
