#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/bench_n.sh 1
bash tools/capture_profiles.sh r02 2>&1 | tail -12
