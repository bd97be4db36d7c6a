#!/bin/bash
# ncu capture of the pair-plane FAST kernel as the pipeline runs it (64 x 640x480)
set -u
mkdir -p gpurun_out
export ORBX_FAST_V=2 ORBX_FAST_MIX=2
CMD="python tools/stage_times.py 640 480 1000 64 2"
$CMD > gpurun_out/g2_plain.json 2> gpurun_out/g2_plain.err || { echo "plain run failed"; tail -5 gpurun_out/g2_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_fast_pairs -s 1 -c 1 -o gpurun_out/prof_k_fast_pairs_a -f $CMD > gpurun_out/g2_ncu.log 2>&1
tail -3 gpurun_out/g2_ncu.log
