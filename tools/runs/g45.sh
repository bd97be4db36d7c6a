#!/bin/bash
set -u
export KNN_SUSTAIN=0
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
for rot in 0 1; do
  echo "== ORBX_KNN_ROTATE=$rot"
  ORBX_KNN_ROTATE=$rot timeout 200 python tools/knn_time.py 1000000 10000000 2>&1 | grep -E "fp4|i8"
done
