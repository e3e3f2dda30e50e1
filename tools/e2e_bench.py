#!/usr/bin/env python3
"""End-to-end timing of qpmpc_b200_solve_host (pinned host buffers) on config 2: sweep of the
pipelining chunk (QPMPC_B200_HOST_CHUNK) next to the device-resident kernel time."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from qpmpc_b200 import _capi, solve_mpc_batch
from qpmpc_b200.workloads import to_batched, triple_integrator_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = 16
lib = _capi.load()
sets = [triple_integrator_batch(B, N=N, seed=s) for s in range(4)]
prob = to_batched(sets[0])
host = [{k: torch.from_numpy(np.ascontiguousarray(w[k])).pin_memory() for k in ("A", "B", "C", "e", "x0", "goal")}
        for w in sets]
U = torch.empty((B, N), dtype=torch.float64).pin_memory()
st = torch.empty(B, dtype=torch.int32).pin_memory()
desc = prob.desc()
vp = lambda t: ctypes.c_void_p(t.data_ptr())


def step(i):
    hs = host[i % 4]
    ops = _capi.Operands(vp(hs["A"]), vp(hs["B"]), vp(hs["C"]), None, vp(hs["e"]), vp(hs["x0"]), vp(hs["goal"]), None)
    outs = _capi.Outputs(vp(U), vp(st), None, None)
    rc = lib.qpmpc_b200_solve_host(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs), 0)
    assert rc == 0, rc


for chunk in [int(c) for c in os.environ.get("CHUNKS", "65536 32768 16384 8192 4096").split()]:
    os.environ["QPMPC_B200_HOST_CHUNK"] = str(chunk)
    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 40
    for i in range(reps):
        step(i)
    dt = (time.perf_counter() - t0) / reps
    assert int((st != 0).sum()) == 0
    print(f"chunk {chunk:6d}: {dt * 1e3:.3f} ms/step  {B / dt / 1e6:.1f} M solves/s", flush=True)
for _ in range(3):
    solve_mpc_batch(prob)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    solve_mpc_batch(prob)
e1.record()
torch.cuda.synchronize()
print(f"device-resident kernel: {e0.elapsed_time(e1) / 20:.3f} ms")

# ---- zero-copy: the kernel's bulk-TMA staging reads the pinned host buffers over PCIe itself and
# the epilogue stores U rows into pinned host memory (UVA: the host pointer is the device pointer)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
st_dev = torch.empty(B, dtype=torch.int32, device="cuda")
for mode in ("U and status to host", "U to host, status via one D2H copy"):
    def zstep(i):
        hs = host[i % 4]
        ops = _capi.Operands(vp(hs["A"]), vp(hs["B"]), vp(hs["C"]), None, vp(hs["e"]), vp(hs["x0"]), vp(hs["goal"]), None)
        outs = _capi.Outputs(vp(U), vp(st) if mode.startswith("U and") else ctypes.c_void_p(st_dev.data_ptr()), None, None)
        rc = lib.qpmpc_b200_solve(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs), stream)
        assert rc == 0, rc
        if not mode.startswith("U and"):
            st.copy_(st_dev, non_blocking=True)
        torch.cuda.synchronize()
    U.zero_()
    for i in range(4):
        zstep(i)
    t0 = time.perf_counter()
    for i in range(40):
        zstep(i)
    dt = (time.perf_counter() - t0) / 40
    ref = solve_mpc_batch(to_batched(sets[39 % 4]))
    torch.cuda.synchronize()
    err = (ref.inputs.reshape(B, -1).cpu() - U).abs().max().item()
    print(f"zero-copy ({mode}): {dt * 1e3:.3f} ms/step  {B / dt / 1e6:.1f} M solves/s  |dU| vs device path {err:.1e} "
          f"unsolved {int((st != 0).sum())}", flush=True)
