#!/usr/bin/env python3
"""Randomised differential soak on the device: random problem shapes and operand patterns
through every dispatch route (warp kernels paired / unpaired, long-horizon kernel, CTA kernel and
its workspace path, shared-model path, interior point) against the CPU oracle.

    python tools/soak.py [--seconds 120] [--seed 0]

Per case: statuses must agree, U within 1e-6 (the bar of the parity tests) of the oracle on the
instances both solve, relative to max(1, |U|); prints one line per mismatch and a closing
summary (JSON).  Open-loop unstable models (the pendulum, random A = I + 0.2 randn) are kept to
horizons where A^N has not made the QP ill conditioned (N <= 18 / 14: beyond, two correct fp64
algorithms differ by more than the bar -- a first run of this tool showed exactly that and
nothing else).  The oracle is the checker, never the product.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from qpmpc_b200 import BackendError, factor_model, solve_mpc_batch
from qpmpc_b200.workloads import (humanoid_batch, oracle_ops, pendulum_batch, random_batch, to_batched,
                                  triple_integrator_batch)

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=120.0)
ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
t_end = time.time() + a.seconds
stats = dict(cases=0, instances=0, mismatches=0, worst=0.0, routes={})


def check(w, tag, **kw):
    prob = to_batched(w)
    try:
        if kw.pop("factored", False):
            kw["factored"] = factor_model(prob)
        plan = solve_mpc_batch(prob, **kw)
    except BackendError as exc:  # (a shape the route does not take: not a mismatch)
        stats["routes"][tag + " refused"] = stats["routes"].get(tag + " refused", 0) + 1
        return
    torch.cuda.synchronize()
    ref = oracle.solve_batch(w["batch"], w["N"], w["nx"], w["nu"], w["nc"], oracle_ops(w), w["w_t"], w["w_x"], w["w_u"])
    st = plan.status.cpu().numpy()
    U = plan.inputs.reshape(w["batch"], -1).cpu().numpy()
    ok = (st == 0) & (ref["status"] == 0)
    bad_status = int(((st == 0) != (ref["status"] == 0)).sum())
    err = 0.0
    if ok.any():
        scale = np.maximum(1.0, np.abs(ref["U"][ok]).max(axis=1))
        err = float((np.abs(U[ok] - ref["U"][ok]).max(axis=1) / scale).max())
    stats["cases"] += 1
    stats["instances"] += w["batch"]
    stats["worst"] = max(stats["worst"], err)
    stats["routes"][tag] = stats["routes"].get(tag, 0) + 1
    tol = 1e-6 if kw.get("method") != "pdip" else 2e-6
    if bad_status or err > tol or not np.isnan(U[st != 0]).all():
        stats["mismatches"] += 1
        print(f"MISMATCH {tag} {w['name']} N={w['N']} nx={w['nx']} nu={w['nu']} nc={w['nc']} batch={w['batch']}: "
              f"status differs on {bad_status}, |dU| {err:.2e}", flush=True)


while time.time() < t_end:
    seed = int(rng.integers(1 << 30))
    kind = int(rng.integers(8))
    B = int(rng.integers(1, 200))
    if kind == 0:    # paired warp kernels / long-horizon kernel, any horizon
        N = int(rng.integers(1, 65))
        w = triple_integrator_batch(B if N <= 32 else min(B, 24), N=N, seed=seed, per_instance_model=bool(rng.integers(2)))
        check(w, f"ti N<={8 if N <= 8 else 16 if N <= 16 else 32 if N <= 32 else 64}")
    elif kind == 1:  # random dense shapes: unpaired rows, C and D, LTV
        nu, nx, nc = int(rng.integers(1, 4)), int(rng.integers(1, 7)), int(rng.integers(1, 6))
        N = int(rng.integers(1, max(2, min(14, 32 // nu))))
        w = random_batch(B, N, nx, nu, nc, seed=seed, ltv=bool(rng.integers(2)), with_C=bool(rng.integers(2)) or nc > 0,
                         with_D=bool(rng.integers(2)), w_t=None if rng.integers(4) == 0 else 0.7,
                         w_x=0.3 if rng.integers(2) else None, w_u=10.0 ** rng.uniform(-2, 0))
        if w["w_t"] is None and w["w_x"] is None:
            w["w_t"] = 0.5
        if w["C"] is None and w["D"] is None:
            continue
        check(w, "random warp")
    elif kind == 2:  # CTA kernel incl. its workspace path (n > 32)
        nu = int(rng.integers(1, 3))
        N = int(rng.integers(33 // nu + 1, 100 // nu))
        w = random_batch(min(B, 12), N, int(rng.integers(2, 5)), nu, int(rng.integers(1, 4)), seed=seed, ltv=True, w_u=1.0)
        check(w, "random cta" if N * nu <= 72 else "random cta workspace")
    elif kind == 3:
        check(pendulum_batch(B, N=int(rng.integers(2, 19)), seed=seed, ltv_model=bool(rng.integers(2))), "pendulum")
    elif kind == 4:
        check(humanoid_batch(B, N=int(rng.integers(2, 65)), seed=seed), "humanoid")
    elif kind == 5:  # shared-model path
        N = int(rng.integers(2, 33))
        w = pendulum_batch(B, N=min(N, 18), seed=seed) if rng.integers(2) else humanoid_batch(B, N=N, seed=seed)
        check(w, "factored", factored=True)
    elif kind == 6:  # interior point (n <= 32)
        N = int(rng.integers(2, 33))
        w = triple_integrator_batch(B, N=N, seed=seed) if rng.integers(2) else pendulum_batch(B, N=min(N, 16), seed=seed)
        check(w, "pdip", method="pdip")
    else:            # forced routes on small shapes
        os.environ["QPMPC_B200_FORCE_CTA"] = "1"
        if rng.integers(2):
            os.environ["QPMPC_B200_CTA_WORKSPACE"] = "1"
        try:
            check(random_batch(B, int(rng.integers(2, 12)), 3, int(rng.integers(1, 3)), int(rng.integers(1, 5)), seed=seed), "forced cta")
        finally:
            os.environ.pop("QPMPC_B200_FORCE_CTA", None)
            os.environ.pop("QPMPC_B200_CTA_WORKSPACE", None)
print(json.dumps(stats))
sys.exit(1 if stats["mismatches"] else 0)
