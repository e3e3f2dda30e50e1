#!/bin/bash
# First device session of the interior-point kernel.
# One gpurun call, every step under its own `timeout` so that a hang costs that
# step only; everything lands in gpurun_out/ with the given tag.
#
#   gpurun --timeout 420 -- 'bash tools/pdip_gpu_session.sh r02a'
set -u
TAG=${1:-pdip}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1

echo "== 1. two instances, smallest possible launch (60 s)"
timeout 60 python - > $OUT/${TAG}_pdip_tiny.txt 2>&1 <<'PY'
import numpy as np, torch, oracle
from qpmpc_b200 import solve_mpc_batch
from qpmpc_b200.workloads import oracle_ops, to_batched, triple_integrator_batch
for polish in (False, True):
    w = triple_integrator_batch(2, seed=0)
    plan = solve_mpc_batch(to_batched(w), method="pdip", polish=polish)
    torch.cuda.synchronize()
    ref = oracle.solve_batch(2, 16, 3, 1, 2, oracle_ops(w), 1.0, None, 1e-6)
    print("polish", polish, "status", plan.status.tolist(), "iters", plan.iters.tolist(),
          "err", float(np.abs(plan.inputs.reshape(2, -1).cpu().numpy() - ref["U"]).max()), flush=True)
PY
echo "rc=$?"; tail -3 $OUT/${TAG}_pdip_tiny.txt

echo "== 2. the GPU test module (120 s)"
timeout 120 python -m pytest tests/test_gpu_pdip.py -x -q -p no:cacheprovider > $OUT/${TAG}_pdip_pytest.txt 2>&1
echo "rc=$?"; tail -3 $OUT/${TAG}_pdip_pytest.txt

echo "== 3. throughput next to the active-set kernel, config 2 (90 s)"
timeout 90 python tools/pdip_bench.py > $OUT/${TAG}_pdip_bench.jsonl 2> $OUT/${TAG}_pdip_bench.err
echo "rc=$?"; cat $OUT/${TAG}_pdip_bench.jsonl

echo "== 3b. the bench contract with the interior-point kernel (120 s)"
timeout 120 python bench.py --method pdip --steps 64 --warmup 4 --cpu-seconds 2 > $OUT/${TAG}_bench_pdip.json 2> $OUT/${TAG}_bench_pdip.err
echo "rc=$?"; cat $OUT/${TAG}_bench_pdip.json

echo "== 4. compute-sanitizer synccheck + memcheck on a small launch (90 s)"
for tool in synccheck memcheck; do
  timeout 90 compute-sanitizer --tool $tool python - > $OUT/${TAG}_pdip_${tool}.txt 2>&1 <<'PY'
import torch
from qpmpc_b200 import solve_mpc_batch
from qpmpc_b200.workloads import pendulum_batch, to_batched, triple_integrator_batch
for w in (triple_integrator_batch(67, seed=1), pendulum_batch(35)):
    plan = solve_mpc_batch(to_batched(w), method="pdip")
    torch.cuda.synchronize()
    print(w["name"], "solved", int((plan.status == 0).sum()), "of", w["batch"], flush=True)
PY
  echo "$tool rc=$?"; tail -2 $OUT/${TAG}_pdip_${tool}.txt
done

echo "== 5. ncu --set full of one interior-point launch (120 s)"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:mpc_pdip_kernel -c 1 \
  -o $OUT/${TAG}_pdip_prof -f python tools/pdip_bench.py --steps 2 --sets 1 --check 64 > $OUT/${TAG}_pdip_ncu.log 2>&1
echo "rc=$?"
