#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu --no-knn --no-two-callers > gpurun_out/g28.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/g28.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'single', round(d['single_lane']['value']), 'e2e', round(d['e2e']['value']), 'cfg1', round(d['config1_752x480_nf1200']['value']), 'cfg3', round(d['config3_1920x1080_nf2000']['value']), 'cfg5', round(d['config5_1280x720_nf1250']['value']))
print([(x['shape'], x['nfeatures'], round(x['median_ms'],4)) for x in d['latency']])
print(d['config5_1280x720_nf1250']['per_frame_stream'])
PY
