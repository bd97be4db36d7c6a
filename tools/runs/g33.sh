#!/bin/bash
set -u
mkdir -p gpurun_out
export ORBX_BLUR_TC=2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for cfg in "640 480 1000 64" "1920 1080 2000 16"; do
  ORBX_DEV_SPLIT=1 timeout 200 python tools/stage_times.py $cfg 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['shape'], d['stage_us'], d['result_sha1'])"
done
for tc in 2 1 0; do
for ctas in 2 1; do
[ $tc != 1 -a $ctas = 1 ] && continue
ORBX_BLUR_TC=$tc ORBX_BLUR_TC_CTAS=$ctas timeout 600 python bench.py --no-cpu --no-knn --no-latency --no-two-callers --no-euroc > gpurun_out/g33_$tc.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/g33_$tc.json').read().strip().splitlines()[-1])
print('ORBX_BLUR_TC=$tc ctas=$ctas value', round(d['value']), 'single', round(d['single_lane']['value']), 'e2e', round(d['e2e']['value']), 'blur', d['roofline']['stage_ms']['blur'])
PY
done
done
