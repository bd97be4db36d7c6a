#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_shim_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/stage_sweep.py "" "ORBX_BLUR_WORDS=1" > gpurun_out/g16_sweep.jsonl 2>&1; cut -c1-300 gpurun_out/g16_sweep.jsonl
timeout 100 python tools/whatif.py 2>&1 | tail -1
