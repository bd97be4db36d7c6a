#!/bin/bash
set -u
mkdir -p gpurun_out
for rep in 1 2 3; do
for tc in 0 1; do
ORBX_BLUR_TC=$tc ORBX_BLUR_TC_CTAS=1 timeout 600 python bench.py --no-cpu --no-knn --no-latency --no-two-callers --no-euroc --steps 40 > gpurun_out/g34.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/g34.json').read().strip().splitlines()[-1])
print('rep $rep ORBX_BLUR_TC=$tc value', round(d['value']), 'single', round(d['single_lane']['value']), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']))
PY
done
done
