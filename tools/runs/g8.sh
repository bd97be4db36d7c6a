#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/g8_e2e.jsonl
for cfg in "ORBX_E2E_LANES=4" "ORBX_E2E_LANES=2" "ORBX_E2E_LANES=8" "ORBX_E2E_LANES=4 ORBX_CHUNK=16" "ORBX_E2E_LANES=4 ORBX_CHUNK=64" "ORBX_E2E_LANES=8 ORBX_CHUNK=64" \
           "ORBX_E2E_LANES=4 ORBX_SUB=1" "ORBX_E2E_LANES=4 ORBX_SUB=4" "ORBX_E2E_LANES=8 ORBX_SUB=1" "ORBX_E2E_LANES=8 ORBX_CHUNK=64 ORBX_SUB=4" "ORBX_E2E_LANES=4 ORBX_STREAMS=8"; do
  env $cfg timeout 120 python tools/e2e_stream.py 120 >> gpurun_out/g8_e2e.jsonl 2>> gpurun_out/g8_e2e.err
done
cat gpurun_out/g8_e2e.jsonl
timeout 120 python tools/h2d_probe.py >> gpurun_out/g8_h2d.jsonl 2>&1; cat gpurun_out/g8_h2d.jsonl
