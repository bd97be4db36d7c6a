#!/bin/bash
set -u
export KNN_SUSTAIN=0
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_nif_mock_gpu.py -x -q -k "knn2" 2>&1 | tail -3
timeout 200 python tools/knn_time.py 300000 1000000 10000000 2>&1 | grep -E "fp4|i8"
