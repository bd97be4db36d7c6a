#!/bin/bash
set -u
mkdir -p gpurun_out; : > gpurun_out/g11_whatif.jsonl
for cfg in "ORBX_FAST_CTAS=6" "ORBX_FAST_CTAS=5" "ORBX_FAST_CTAS=4" "ORBX_FAST_CTAS=3" "ORBX_FAST_CTAS=5 ORBX_NO_CONE=1" "ORBX_FAST_CTAS=4 ORBX_NO_CONE=1" "ORBX_FAST_CTAS=5 ORBX_DEV_SPLIT=2" "ORBX_FAST_CTAS=5 ORBX_DEV_SPLIT=1" "ORBX_DEV_SPLIT=1" "ORBX_DEV_SPLIT=2"; do
  echo "$cfg" >> gpurun_out/g11_whatif.jsonl
  env $cfg timeout 100 python tools/whatif.py >> gpurun_out/g11_whatif.jsonl 2>&1
done
cat gpurun_out/g11_whatif.jsonl
