#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g10_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g10_pytest.log
timeout 300 python tools/stage_sweep.py "" "ORBX_NO_CONE=1" > gpurun_out/g10_sweep.jsonl 2>&1; cut -c1-300 gpurun_out/g10_sweep.jsonl
: > gpurun_out/g10_whatif.jsonl
for m in 0 1; do ORBX_SKIP_STAGES=$m timeout 100 python tools/whatif.py >> gpurun_out/g10_whatif.jsonl 2>&1; done
ORBX_NO_CONE=1 timeout 100 python tools/whatif.py >> gpurun_out/g10_whatif.jsonl 2>&1
cat gpurun_out/g10_whatif.jsonl
