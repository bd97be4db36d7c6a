#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/knn_time.py 300000 > gpurun_out/g17_knn_small.jsonl 2>&1; cat gpurun_out/g17_knn_small.jsonl | tail -5
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "knn2" 2>&1 | tail -4
timeout 300 python tools/knn_time.py > gpurun_out/g17_knn.jsonl 2>&1; cat gpurun_out/g17_knn.jsonl
