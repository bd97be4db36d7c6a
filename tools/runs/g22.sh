#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "frame_to_frame or windowed or grid" 2>&1 | tail -4
timeout 600 python bench.py --no-cpu --no-knn --no-latency > gpurun_out/g22_bench.json 2> gpurun_out/g22_bench.err; echo rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/g22_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('per_rank'))
print(json.dumps(d['config3_1920x1080_nf2000'])[:900])
PY
