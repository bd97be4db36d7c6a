#!/bin/bash
set -u
export KNN_SUSTAIN=0
cp send_slam_b200/liborbx.so /tmp/liborbx_new.so
for v in base new0 new1 new3; do
  if [ $v = base ]; then cp send_slam_b200/liborbx_base.so.keep send_slam_b200/liborbx.so; else cp /tmp/liborbx_new.so send_slam_b200/liborbx.so; fi
  echo "== $v"
  ORBX_KNN_ROTATE=${v#new} timeout 200 python tools/knn_time.py 1000000 1250000 10000000 2>&1 | grep -E "fp4"
done
cp /tmp/liborbx_new.so send_slam_b200/liborbx.so
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "knn2" 2>&1 | tail -2
