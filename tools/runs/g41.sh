#!/bin/bash
set -u
mkdir -p gpurun_out
for fc in 6 5 4; do
ORBX_FAST_CTAS=$fc timeout 600 python bench.py --no-cpu --no-knn --no-latency --no-two-callers --no-euroc --steps 40 > gpurun_out/g41.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/g41.json').read().strip().splitlines()[-1])
print('FAST_CTAS=$fc value', round(d['value']), 'single', round(d['single_lane']['value']), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']))
PY
done
