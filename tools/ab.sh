#!/bin/bash
# A/B of library builds on one GPU: tools/ab.sh <tag> "<env assignments>" ... ; each runs tools/tune.py
TAG=$1; shift
OUT=gpurun_out
for cfg in "$@"; do
  echo "== $cfg" | tee -a $OUT/${TAG}_ab.txt
  env $cfg python tools/tune.py --wpc ${WPC:-8} --N ${NS:-16} 2>&1 | tee -a $OUT/${TAG}_ab.txt
done
