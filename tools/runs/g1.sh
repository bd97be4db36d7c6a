#!/bin/bash
# first GPU call of round 2: parity suite on the pair-plane FAST kernel, pipe microbenchmarks, FAST variant sweep, short bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/g1_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/g1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/g1_pytest.log
tail -5 gpurun_out/g1_pytest.log
timeout 60 tools/_build/ubench_mix > gpurun_out/g1_ubench_mix.jsonl 2>&1
timeout 120 tools/_build/blur_tma_probe > gpurun_out/g1_blur_probe.txt 2>&1
timeout 500 python tools/stage_sweep.py "ORBX_FAST_V=1" "ORBX_FAST_V=2 ORBX_FAST_MIX=0" "ORBX_FAST_V=2 ORBX_FAST_MIX=1" "ORBX_FAST_V=2 ORBX_FAST_MIX=2" \
  "ORBX_FAST_V=2 ORBX_FAST_MIX=3" "ORBX_FAST_V=2 ORBX_FAST_MIX=4" "ORBX_FAST_V=2 ORBX_FAST_MIX=5" "ORBX_FAST_V=2 ORBX_FAST_MIX=6" "ORBX_FAST_V=2 ORBX_FAST_MIX=7" \
  "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_CH=12" "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_CH=16" "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_CH=20" \
  "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_CH=32" "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_CH=64" \
  "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_WARPS=1" "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_WARPS=4" "ORBX_FAST_V=2 ORBX_FAST_MIX=2 ORBX_FAST_WARPS=3" \
  "ORBX_FAST_V=2 ORBX_FAST_MIX=0 ORBX_FAST_CH=16" "ORBX_FAST_V=2 ORBX_FAST_MIX=4 ORBX_FAST_CH=16" \
  > gpurun_out/g1_sweep.jsonl 2>&1
cat gpurun_out/g1_sweep.jsonl | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/g1_bench.json
