#!/bin/bash
# Run on the GPU box (under gpurun): launch list of the bench command + one `ncu --set full` capture per hot kernel.
# Every ncu run follows a plain run of the same command that exited 0.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r02}
export ORBX_DEV_SPLIT=1 ORBX_GRAPHS=0 ORBX_BENCH_LANES=1   # one lane, single-range launches, no graph replay: every launch is one whole-batch kernel
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-euroc --no-latency"
$CMD > gpurun_out/bench_plain_$TAG.json 2> gpurun_out/bench_plain_$TAG.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
for k in k_fast_tma k_blur_tc k_describe_tma k_octree k_finalize k_knn2_fp4 k_knn2_tc k_knn2 k_pyramid_cone; do
  skip=3; [ $k = k_pyramid_cone ] && skip=6     # even instances = the first (four-level) cone
  [ $k = k_knn2_tc -o $k = k_knn2_fp4 ] && skip=5      # odd instances = the second (main) pass over the shard
  ncu --set full --clock-control none --import-source on -k regex:^$k -s $skip -c 1 -o gpurun_out/prof_${k}_$TAG -f $CMD > gpurun_out/ncu_${k}_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_${k}_$TAG.log
done
