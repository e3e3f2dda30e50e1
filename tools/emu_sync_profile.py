#!/usr/bin/env python
"""Warp-level synchronisation points per warp (shuffles, votes, reductions,
__syncwarp) of the two solver kernels on samples of the BASELINE workloads,
counted on the host emulator (tests/emu).  No GPU needed.  The kernels are
latency-bound chains of such points, so the ratio is a first, hardware-free
estimate of the interior-point kernel's cost relative to the active-set kernel
(it ignores the arithmetic between the points, which is heavier for the
interior point: m rank-one updates of H per iteration)."""

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import emu  # noqa: E402
from qpmpc_b200.workloads import humanoid_batch, pendulum_batch, triple_integrator_batch  # noqa: E402

lib = emu.load()
B = 64

if "--lines" in sys.argv:
    # where the synchronisation points of one kernel are: per source line, config 2
    method = "pdip" if "pdip" in sys.argv else "active_set"
    src = os.path.join(ROOT, "qpmpc_b200", "csrc", "mpc_pdip.cuh" if method == "pdip" else "mpc_kernels.cuh")
    text = open(src).read().splitlines()
    w = triple_integrator_batch(B, seed=0)
    lib.emu_sync_points_at(0, 1)
    emu.solve(w, method=method)
    hist = sorted(((lib.emu_sync_points_at(i, 0), i) for i in range(1, 4096)), reverse=True)
    total = sum(c for c, _ in hist)
    warps = B * 16 / 32.0
    print(f"{method}: {total / 32.0 / warps:.1f} synchronisation points per warp; by source line of "
          f"{os.path.basename(src)} (helpers in other headers are listed by their own line numbers):")
    for c, i in hist[:18]:
        if c:
            print(f"  {c / 32.0 / warps:7.1f}  {100.0 * c / total:5.1f} %  line {i:4d}: {text[i - 1].strip()[:90] if i <= len(text) else ''}")
    sys.exit(0)

for name, w, np_ in (("config 2: triple integrator N=16", triple_integrator_batch(B, seed=0), 16),
                     ("config 3: pendulum N=12", pendulum_batch(B, seed=1), 16),
                     ("config 4: humanoid N=16", humanoid_batch(B, seed=2), 16),
                     ("config 5: triple integrator N=8", triple_integrator_batch(B, N=8, seed=3), 8),
                     ("config 5: triple integrator N=32", triple_integrator_batch(16, N=32, seed=3), 32)):
    row = {"workload": name, "instances": w["batch"]}
    for method, kw, ls in (("active_set", {}, "0"), ("pdip", {"tol": 1e-9}, "0"),
                           ("pdip_tol1e-6", {"method": "pdip", "tol": 1e-6}, "0"),
                           ("pdip_LS", {"method": "pdip", "tol": 1e-9}, "1")):
        os.environ["QPMPC_B200_PDIP_SOLVE"] = ls  # 1: solves through L^-1
        before = lib.emu_sync_points()
        got = emu.solve(w, **{"method": method, **kw})
        warps = w["batch"] * np_ / 32.0
        row[method] = {"sync_points_per_warp": round((lib.emu_sync_points() - before) / 32.0 / warps, 1),
                       "iters_mean": round(float(got["iters"].mean()), 2),
                       "solved": float((got["status"] == 0).mean())}
    row["pdip_over_active_set"] = round(row["pdip"]["sync_points_per_warp"] / row["active_set"]["sync_points_per_warp"], 2)
    print(json.dumps(row), flush=True)
