#!/bin/bash
set -u
mkdir -p gpurun_out
export KNN_SUSTAIN=0
for sp in 0 1; do
  echo "== ORBX_KNN_SEED_SPLIT=$sp"
  ORBX_KNN_SEED_SPLIT=$sp timeout 200 python tools/knn_time.py 300000 1000000 3000000 10000000 2>&1 | grep fp4
done
