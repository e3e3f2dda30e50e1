#!/usr/bin/env python3
"""BASELINE config 5: horizon sweep N in {8,16,32,64} x batch in {1k,16k,256k,1M},
triple integrator with T = 1/N, fp64, seed 3, one GPU (device-resident, CUDA
events, kernel only).  Prints one JSON line per point and a closing summary.

    python tools/sweep.py [--N 8 16 32 64] [--batch 1024 16384 262144 1048576]
                          [--cpu-seconds 2] [--out gpurun_out/sweep.json]

The CPU column is the C oracle (oracle/, OpenMP, all host threads) on a bounded
sample of the same instances; it is a reported baseline, not the product.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from qpmpc_b200 import solve_mpc_batch
from qpmpc_b200.workloads import (algorithmic_bytes_per_solve, oracle_ops, slice_workload,
                                  to_batched, triple_integrator_batch)

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, nargs="*", default=[8, 16, 32, 64])
ap.add_argument("--batch", type=int, nargs="*", default=[1024, 16384, 262144, 1048576])
ap.add_argument("--cpu-seconds", type=float, default=2.0)
ap.add_argument("--out", default=None)
ap.add_argument("--check", type=int, default=512, help="instances compared with the oracle per point")
a = ap.parse_args()

rows = []
for N in a.N:
    for B in a.batch:
        if N >= 64 and B > 262144:
            reps = 2
        else:
            reps = 5 if B >= 262144 else 20
        w = triple_integrator_batch(B, N=N, seed=3)
        prob = to_batched(w)
        for _ in range(3):
            plan = solve_mpc_batch(prob)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            plan = solve_mpc_batch(prob)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        it = plan.iters.float()
        row = dict(N=N, batch=B, ms=ms, solves_per_s=B / ms * 1e3, iters_mean=float(it.mean()),
                   iters_max=int(it.max()), unsolved=int((plan.status != 0).sum()),
                   bytes_per_solve=algorithmic_bytes_per_solve(w))
        row["hbm_gbs"] = row["solves_per_s"] * row["bytes_per_solve"] / 1e9
        if a.check or a.cpu_seconds > 0:
            import oracle

            k = min(B, max(a.check, 1))
            ws = slice_workload(w, 0, k)
            args = (k, N, 3, 1, 2, oracle_ops(ws), ws["w_t"], ws["w_x"], ws["w_u"])
            ref = oracle.solve_batch(*args)
            U = plan.inputs.reshape(B, -1)[:k].cpu().numpy()
            ok = ref["status"] == 0
            row["max_abs_err_vs_oracle"] = float(np.abs(U[ok] - ref["U"][ok]).max())
            if a.cpu_seconds > 0:
                kc = min(B, 16384)
                wc = slice_workload(w, 0, kc)
                cargs = (kc, N, 3, 1, 2, oracle_ops(wc), wc["w_t"], wc["w_x"], wc["w_u"])
                oracle.solve_batch(*cargs)
                r, t0 = 0, time.perf_counter()
                while time.perf_counter() - t0 < a.cpu_seconds:
                    oracle.solve_batch(*cargs)
                    r += 1
                row["cpu_solves_per_s"] = kc * r / (time.perf_counter() - t0)
                row["cpu_threads"] = oracle.num_threads()
        print(json.dumps(row), flush=True)
        rows.append(row)
        del prob, plan
        torch.cuda.empty_cache()
if a.out:
    with open(a.out, "w") as f:
        json.dump(rows, f, indent=1)
