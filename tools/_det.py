import sys, os, ctypes
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from qpmpc_b200 import _capi, solve_mpc_batch
from qpmpc_b200.workloads import to_batched, triple_integrator_batch
w = triple_integrator_batch(1000, seed=9)
prob = to_batched(w)
plan = solve_mpc_batch(prob, return_multipliers=True)
torch.cuda.synchronize()
lib = _capi.load()
desc = prob.desc()
arrs = {k: np.ascontiguousarray(w[k]) for k in ("A", "B", "C", "e", "x0", "goal")}
ptr = lambda a: ctypes.c_void_p(a.ctypes.data)
ops = _capi.Operands(ptr(arrs["A"]), ptr(arrs["B"]), ptr(arrs["C"]), None, ptr(arrs["e"]), ptr(arrs["x0"]), ptr(arrs["goal"]), None)
for trial in range(3):
    U = np.zeros((1000, 16)); st = np.zeros(1000, dtype=np.int32); it = np.zeros(1000, dtype=np.int32)
    outs = _capi.Outputs(ptr(U), ptr(st), ptr(it), None)
    rc = lib.qpmpc_b200_solve_host(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs), 0)
    V = plan.inputs.reshape(1000, 16).cpu().numpy()
    d = np.abs(U - V)
    bad = np.argwhere(d.max(axis=1) > 0).ravel()
    print(trial, rc, "maxdiff", d.max(), "nbad", bad.size, bad[:8], "iters host", it[bad[:8]], "dev", plan.iters.cpu().numpy()[bad[:8]], "st", st[bad[:8]])
