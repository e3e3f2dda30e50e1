#!/usr/bin/env python3
"""Times the tcgen05 Hessian contraction (mpc_hessian_tc.cuh) next to the SIMT accumulation of the
CTA condensing kernel: fp32 pendulum model with a stage cost, N = 64 (n = 64, K = 260)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from qpmpc_b200 import condense_batch
import numpy as np
from qpmpc_b200.workloads import to_batched, triple_integrator_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
w = triple_integrator_batch(B, N=64, seed=1, per_instance_model=False)
w["w_x"], w["targets"] = 0.5, np.zeros((B, 64 * 3))
w["A"], w["B"], w["ltv"] = np.tile(w["A"], (64, 1, 1)), np.tile(w["B"], (64, 1, 1)), ("A", "B")
prob = to_batched(w, dtype=torch.float32)
fields = ("P", "Psi", "psi_last")
for tc in ("1", "0"):
    os.environ["QPMPC_B200_HESSIAN_TC"] = tc
    for _ in range(3):
        out = condense_batch(prob, fields)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = condense_batch(prob, fields)
    e1.record()
    torch.cuda.synchronize()
    print(f"HESSIAN_TC={tc}: condense_batch(P, Psi, psi_last) of {B} instances: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
