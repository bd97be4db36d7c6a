#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python tests/checks/fuzz_parity.py 150 21 2>&1 | tail -3 | tee gpurun_out/fuzz_r02_a.txt
ORBX_BLUR_TC=0 timeout 300 python tests/checks/fuzz_parity.py 90 22 2>&1 | tail -3 | tee gpurun_out/fuzz_r02_b.txt
