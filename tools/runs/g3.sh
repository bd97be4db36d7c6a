#!/bin/bash
set -u
mkdir -p gpurun_out
ORBX_FAST_V=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/g3_pytest_v1.log 2>&1; echo "pytest v1 rc=$?"; tail -3 gpurun_out/g3_pytest_v1.log
ORBX_FAST_V=2 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/g3_pytest_v2.log 2>&1; echo "pytest v2 rc=$?"; tail -3 gpurun_out/g3_pytest_v2.log
timeout 500 python tools/stage_sweep.py "ORBX_FAST_V=1" "ORBX_FAST_V=2 ORBX_FAST_MIX=2" "ORBX_FAST_V=2 ORBX_FAST_MIX=0" "ORBX_FAST_V=2 ORBX_FAST_MIX=4" \
  "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_CH=20" "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_CH=40" "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_WARPS=4" \
  "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_WARPS=1" "ORBX_FAST_V=2 ORBX_FAST_MIX=0 ORBX_FAST_WARPS=4" "ORBX_FAST_V=2 ORBX_FAST_MIX=5" \
  > gpurun_out/g3_sweep.jsonl 2>&1
cut -c1-330 gpurun_out/g3_sweep.jsonl
