#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g14_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g14_pytest.log
timeout 300 python tools/stage_sweep.py "" "ORBX_NO_CONE=1" > gpurun_out/g14_sweep.jsonl 2>&1; cut -c1-300 gpurun_out/g14_sweep.jsonl
: > gpurun_out/g14_whatif.jsonl
for cfg in "A=0" "ORBX_NO_CONE=1" "ORBX_DEV_SPLIT=1" "ORBX_DEV_SPLIT=2" "ORBX_DEV_SPLIT=1 ORBX_NO_CONE=1"; do echo "$cfg" >> gpurun_out/g14_whatif.jsonl; env $cfg timeout 100 python tools/whatif.py >> gpurun_out/g14_whatif.jsonl 2>&1; done
cat gpurun_out/g14_whatif.jsonl
