#!/bin/bash
# One gpurun call: GPU tests, then the bench contract on every config; everything lands in
# gpurun_out/ with the given tag.   gpurun --timeout 1200 -- 'bash tools/gpu_session.sh r2c'
set -u
TAG=${1:-sess}
WHAT=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
if [ "$WHAT" = all ] || [ "$WHAT" = tests ]; then
  echo "== pytest -m gpu"
  timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_pytest.txt 2>&1
  echo "rc=$?"; tail -5 $OUT/${TAG}_pytest.txt
fi
if [ "$WHAT" = all ] || [ "$WHAT" = bench ]; then
  echo "== bench (default: config 2)"
  timeout 300 python bench.py --cpu-seconds 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  echo "rc=$?"; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
  timeout 200 python bench.py --impl reference --steps 8 --warmup 2 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
  echo "== bench --method pdip"
  timeout 300 python bench.py --method pdip --steps 128 --warmup 4 --cpu-seconds 1 > $OUT/${TAG}_bench_pdip.json 2>> $OUT/${TAG}_bench.err
  echo "rc=$?"; cat $OUT/${TAG}_bench_pdip.json
  for c in 3 4 5; do
    echo "== bench --config $c"
    timeout 400 python bench.py --config $c --cpu-seconds 3 > $OUT/${TAG}_bench_c$c.json 2>> $OUT/${TAG}_bench.err
    echo "rc=$?"; cat $OUT/${TAG}_bench_c$c.json
  done
  for n in 8 32; do
    timeout 300 python bench.py --config 5 --horizon $n --batch 262144 --cpu-seconds 2 > $OUT/${TAG}_bench_c5_N$n.json 2>> $OUT/${TAG}_bench.err
    echo "rc=$?"; cat $OUT/${TAG}_bench_c5_N$n.json
  done
  tail -5 $OUT/${TAG}_bench.err
fi
