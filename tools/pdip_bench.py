#!/usr/bin/env python
"""Interior-point method (``method="pdip"``) next to the default active-set
method on BASELINE config 2 (triple integrator, fp64, N = 16, B = 65 536):
throughput (CUDA events, inputs rotating over distinct sets), mean iteration
counts and the error against the CPU oracle on a subsample (the oracle is the
checker here, as in tests/).  One JSON line per variant."""

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--horizon", type=int, default=16)
ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--sets", type=int, default=3)
ap.add_argument("--check", type=int, default=2048)
ap.add_argument("--workload", default="triple_integrator", choices=["triple_integrator", "pendulum", "humanoid"])
args = ap.parse_args()

import torch  # noqa: E402

import oracle  # noqa: E402
from qpmpc_b200 import solve_mpc_batch  # noqa: E402
from qpmpc_b200.workloads import (humanoid_batch, oracle_ops, pendulum_batch, slice_workload,  # noqa: E402
                                  to_batched, triple_integrator_batch)

B, N = args.batch, args.horizon
if args.workload == "triple_integrator":
    sets = [triple_integrator_batch(B, N=N, seed=100 + s) for s in range(args.sets)]
elif args.workload == "pendulum":
    sets = [pendulum_batch(B, seed=100 + s) for s in range(args.sets)]
else:
    sets = [humanoid_batch(B, seed=100 + s) for s in range(args.sets)]
problems = [to_batched(w) for w in sets]
w0 = slice_workload(sets[0], 0, min(args.check, B))
ref = oracle.solve_batch(w0["batch"], w0["N"], w0["nx"], w0["nu"], w0["nc"], oracle_ops(w0), w0["w_t"], w0["w_x"],
                         w0["w_u"])

variants = [
    ("active_set", dict(method="active_set"), {}),
    ("pdip tol=1e-9 polish wpc=4", dict(method="pdip", tol=1e-9),
     {"QPMPC_B200_PDIP_WPC": "4", "QPMPC_B200_PDIP_SOLVE": "0"}),
    ("pdip tol=1e-9 polish wpc=4 solves through L^-1", dict(method="pdip", tol=1e-9),
     {"QPMPC_B200_PDIP_WPC": "4", "QPMPC_B200_PDIP_SOLVE": "1"}),
    ("pdip tol=1e-9 polish wpc=8", dict(method="pdip", tol=1e-9), {"QPMPC_B200_PDIP_WPC": "8", "QPMPC_B200_PDIP_SOLVE": "0"}),
    ("pdip tol=1e-9 polish wpc=2", dict(method="pdip", tol=1e-9), {"QPMPC_B200_PDIP_WPC": "2"}),
    ("pdip tol=1e-6 polish wpc=4", dict(method="pdip", tol=1e-6), {"QPMPC_B200_PDIP_WPC": "4"}),
    ("pdip tol=1e-9 no polish wpc=4", dict(method="pdip", tol=1e-9, polish=False), {"QPMPC_B200_PDIP_WPC": "4"}),
]
for name, kw, env in variants:
    os.environ.update(env)
    for i in range(3):
        plan = solve_mpc_batch(problems[i % len(problems)], **kw)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for i in range(args.steps):
        solve_mpc_batch(problems[i % len(problems)], **kw)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    plan = solve_mpc_batch(problems[0], **kw)
    torch.cuda.synchronize()
    k = w0["batch"]
    U = plan.inputs.reshape(B, -1)[:k].cpu().numpy()
    st = plan.status.cpu().numpy()
    ok = (st[:k] == 0) & (ref["status"] == 0)
    print(json.dumps({
        "workload": f"{args.workload} fp64 N={sets[0]['N']} batch={B}", "variant": name,
        "ms_per_launch": ms, "solves_per_s": B / (ms * 1e-3),
        "solved_frac": float((st == 0).mean()), "iters_mean": float(plan.iters.float().mean().item()),
        "iters_max": int(plan.iters.max().item()),
        "max_abs_err_vs_oracle": float(np.abs(U[ok] - ref["U"][ok]).max()), "checked": int(ok.sum()),
    }), flush=True)
