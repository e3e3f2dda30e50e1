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
  for c in 3 4 5 6; do
    echo "== bench --config $c"
    timeout 400 python bench.py --config $c --cpu-seconds 3 > $OUT/${TAG}_bench_c$c.json 2>> $OUT/${TAG}_bench.err
    echo "rc=$?"; cat $OUT/${TAG}_bench_c$c.json
  done
  for n in 8 32; do
    timeout 300 python bench.py --config 5 --horizon $n --batch 262144 --cpu-seconds 2 > $OUT/${TAG}_bench_c5_N$n.json 2>> $OUT/${TAG}_bench.err
    echo "rc=$?"; cat $OUT/${TAG}_bench_c5_N$n.json
  done
  echo "== config 2's shape with one shared model: full path against the factored one"
  timeout 200 python tools/shared_model_bench.py > $OUT/${TAG}_shared_model.txt 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_shared_model.txt
  tail -5 $OUT/${TAG}_bench.err
fi
if [ "$WHAT" = all ] || [ "$WHAT" = prof ]; then
  echo "== ncu launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
  echo "rc=$?"; grep -c mpc_ $OUT/${TAG}_launches.csv
  echo "== ncu --set full: headline kernel (config 2), shared-model kernel (config 3), fp32 (config 4)"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:mpc_solve_kernel -c 1 -o $OUT/${TAG}_prof_ti16 -f \
      python tools/tune.py --wpc 8 --N 16 > $OUT/${TAG}_ncu_ti16.log 2>&1; echo "rc=$?"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:mpc_solve_kernel -c 1 -o $OUT/${TAG}_prof_c3 -f \
      python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_c3.log 2>&1; echo "rc=$?"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:mpc_solve_kernel -c 1 -o $OUT/${TAG}_prof_c4 -f \
      python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_c4.log 2>&1; echo "rc=$?"
fi
if [ "$WHAT" = all ] || [ "$WHAT" = sweep ]; then
  echo "== horizon x batch sweep (BASELINE config 5), one GPU"
  timeout 600 python tools/sweep.py --cpu-seconds 1 --out $OUT/${TAG}_sweep.json > $OUT/${TAG}_sweep.jsonl 2>> $OUT/${TAG}_bench.err
  echo "rc=$?"; tail -3 $OUT/${TAG}_sweep.jsonl | cut -c1-600
fi
