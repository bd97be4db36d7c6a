#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for ctas in 1 2; do
  ORBX_BLUR_TC_CTAS=$ctas ORBX_DEV_SPLIT=1 timeout 200 python tools/stage_times.py 640 480 1000 64 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ctas $ctas blur', d['stage_us']['blur'], d['result_sha1'])"
done
for rep in 1 2; do
timeout 600 python bench.py --no-cpu --no-knn --no-latency --no-two-callers --steps 40 > gpurun_out/g40.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/g40.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'single', round(d['single_lane']['value']), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'cfg1', round(d['config1_752x480_nf1200']['value']), 'cfg3', round(d['config3_1920x1080_nf2000']['value']), 'cfg5', round(d['config5_1280x720_nf1250']['value']))
PY
done
