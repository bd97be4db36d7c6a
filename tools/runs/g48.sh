#!/bin/bash
set -u
mkdir -p gpurun_out
for sp in 1 2 4 8; do
ORBX_DEV_SPLIT=$sp timeout 600 python bench.py --no-cpu --no-knn --no-latency --no-two-callers --no-euroc > gpurun_out/g48.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/g48.json').read().strip().splitlines()[-1])
print('DEV_SPLIT=$sp value', round(d['value']), 'single', round(d['single_lane']['value']), 'sustained', round(d['sustained']['value']))
PY
done
