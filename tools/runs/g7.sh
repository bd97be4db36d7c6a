#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python tools/stage_sweep.py "ORBX_FAST_PM=3" "ORBX_FAST_PM=5" "ORBX_FAST_PM=6" "ORBX_FAST_PM=0" > gpurun_out/g7_sweep.jsonl 2>&1
cut -c1-300 gpurun_out/g7_sweep.jsonl
