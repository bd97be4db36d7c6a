#!/bin/bash
set -u
mkdir -p gpurun_out
export ORBX_BLUR_TC=1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for st in 2 3 4; do
for ctas in 1 2; do
  ORBX_BLUR_TC_STAGES=$st ORBX_BLUR_TC_CTAS=$ctas ORBX_DEV_SPLIT=1 timeout 200 python tools/stage_times.py 640 480 1000 64 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stages $st ctas $ctas blur', d['stage_us']['blur'], d['result_sha1'])"
done
done
for st in 3 4; do
ORBX_BLUR_TC_STAGES=$st timeout 600 python bench.py --no-cpu --no-knn --no-latency --no-two-callers --no-euroc --steps 40 > gpurun_out/g36.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/g36.json').read().strip().splitlines()[-1])
print('stages=$st ctas=1 value', round(d['value']), 'single', round(d['single_lane']['value']), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']))
PY
done
