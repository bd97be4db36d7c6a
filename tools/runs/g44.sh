#!/bin/bash
set -u
export KNN_SUSTAIN=0
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "knn2" 2>&1 | tail -2
for rot in 0 1 3; do
  echo "== ORBX_KNN_ROTATE=$rot"
  ORBX_KNN_ROTATE=$rot timeout 200 python tools/knn_time.py 300000 1000000 1250000 10000000 2>&1 | grep fp4
done
