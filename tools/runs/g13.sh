#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python tools/stage_times.py 640 480 1000 64 2"
$CMD > gpurun_out/g13_plain.json 2> gpurun_out/g13_plain.err || { echo "plain run failed"; tail -5 gpurun_out/g13_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_pyramid_cone -s 4 -c 2 -o gpurun_out/prof_k_pyramid_cone_a -f $CMD > gpurun_out/g13_ncu.log 2>&1
tail -3 gpurun_out/g13_ncu.log
