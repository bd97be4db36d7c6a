#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python tools/dev_timeline4.py 4 > gpurun_out/g12_timeline4.txt 2>&1; cat gpurun_out/g12_timeline4.txt | tail -20
timeout 200 python tools/dev_timeline4.py 1 > gpurun_out/g12_timeline1.txt 2>&1; cat gpurun_out/g12_timeline1.txt | tail -20
ORBX_NO_CONE=1 timeout 200 python tools/dev_timeline4.py 4 > gpurun_out/g12_timeline4_nocone.txt 2>&1; cat gpurun_out/g12_timeline4_nocone.txt | tail -20
