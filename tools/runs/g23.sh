#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "640 480 1000 64" "1920 1080 2000 16" "1280 720 1250 32" "752 480 1200 64"; do
  ORBX_DEV_SPLIT=1 timeout 200 python tools/stage_times.py $cfg 6 2>&1 | tail -1 | cut -c1-400
done
