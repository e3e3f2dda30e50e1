#!/bin/bash
# Multi-GPU session: gpurun --gpus N -- 'bash tools/gpu_session_multi.sh <tag> <N>'
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== multi-GPU tests"
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -p no:cacheprovider > $OUT/${TAG}_pytest_multi.txt 2>&1; tail -3 $OUT/${TAG}_pytest_multi.txt
echo "== bench --gpus $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 8 --cpu-seconds 2 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
echo "rc=$?"; cat $OUT/${TAG}_bench_${N}gpu.json; tail -5 $OUT/${TAG}_bench_${N}gpu.err
echo "== reference arm under torchrun"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 4 --warmup 1 > $OUT/${TAG}_bench_ref_${N}gpu.json 2>> $OUT/${TAG}_bench_${N}gpu.err
cat $OUT/${TAG}_bench_ref_${N}gpu.json | cut -c1-400
echo "== host<->device copy ceiling of the box, all ranks at once"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/pcie_probe.py 2>> $OUT/${TAG}_bench_${N}gpu.err | grep '^{' > $OUT/${TAG}_pcie_${N}gpu.json
cat $OUT/${TAG}_pcie_${N}gpu.json
