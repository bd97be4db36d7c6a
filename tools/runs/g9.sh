#!/bin/bash
set -u
mkdir -p gpurun_out; : > gpurun_out/g9_whatif.jsonl
for m in 0 1 2 4 8 16 3 19 27; do ORBX_SKIP_STAGES=$m timeout 100 python tools/whatif.py >> gpurun_out/g9_whatif.jsonl 2>&1; done
cat gpurun_out/g9_whatif.jsonl
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "submit or pinned or batch" 2>&1 | tail -3
