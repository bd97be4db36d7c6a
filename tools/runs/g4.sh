#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/g4_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g4_pytest.log
timeout 300 python tools/stage_sweep.py "" "ORBX_BLUR_WORDS=1" "ORBX_FAST_V=2 ORBX_FAST_CH=20" > gpurun_out/g4_sweep.jsonl 2>&1
cut -c1-330 gpurun_out/g4_sweep.jsonl
timeout 300 python tools/stage_sweep.py "" -- 1920 1080 2000 16 4 > gpurun_out/g4_sweep_1080.jsonl 2>&1; cut -c1-330 gpurun_out/g4_sweep_1080.jsonl
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/g4_bench.json 2> gpurun_out/g4_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g4_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['single_lane']['value'], d['e2e']['value'], d['roofline']['stage_ms'])
PY
