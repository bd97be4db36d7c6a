#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/capture_profiles.sh r02 2>&1 | tail -10
bash tools/bench_n.sh 1
