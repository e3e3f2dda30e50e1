#!/usr/bin/env python3
"""Throughput of the config-2 solve kernel against resident warps per SM, by
padding the kernel's dynamic shared memory (QPMPC_B200_SMEM_PAD_KB)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from qpmpc_b200 import solve_mpc_batch
from qpmpc_b200.workloads import to_batched, triple_integrator_batch

prob = to_batched(triple_integrator_batch(65536, N=16, seed=0))
for wpc, pad in ((4, 160), (4, 50), (4, 0), (2, 190), (2, 80), (2, 42), (2, 23), (2, 12), (2, 5), (2, 0), (1, 0)):
    os.environ["QPMPC_B200_WPC"] = str(wpc)
    os.environ["QPMPC_B200_SMEM_PAD_KB"] = str(pad)
    for _ in range(3):
        solve_mpc_batch(prob)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        solve_mpc_batch(prob)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"wpc={wpc} pad={pad} KB: {ms:.3f} ms  {65536 / ms / 1e3:.1f} M solves/s", flush=True)
