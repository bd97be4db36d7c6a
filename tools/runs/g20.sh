#!/bin/bash
set -u
mkdir -p gpurun_out
export KNN_SUSTAIN=0
timeout 120 python tools/knn_time.py 1000000 > gpurun_out/g20_plain.jsonl 2>&1 || exit 1
cat gpurun_out/g20_plain.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g20_knn_launches.csv python tools/knn_time.py 1000000 > gpurun_out/g20_ncu.log 2>&1
grep -E "k_knn2_fp4|k_expand_fp4|k_knn2_merge|k_knn2_tc|k_expand_pm1" gpurun_out/g20_knn_launches.csv | awk -F'","' '{n=$5; sub(/\(.*/,"",n); print n, $NF}' | tr -d '"' | tail -24
