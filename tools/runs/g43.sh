#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "stages or batch or blur" 2>&1 | tail -2
for lanes in 3 4 5 6; do
ORBX_BENCH_LANES=$lanes timeout 600 python bench.py --no-cpu --no-knn --no-latency --no-two-callers --no-euroc --steps 40 > gpurun_out/g43.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/g43.json').read().strip().splitlines()[-1])
print('lanes=$lanes value', round(d['value']), 'single', round(d['single_lane']['value']), 'sustained', round(d['sustained']['value']))
PY
done
