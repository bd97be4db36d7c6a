#!/bin/bash
set -u
mkdir -p gpurun_out
for wide in 0 2; do
  echo "== ORBX_OCTREE_WIDE=$wide"
  ORBX_OCTREE_WIDE=$wide ORBX_DEV_SPLIT=1 timeout 200 python tools/stage_times.py 640 480 1000 64 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['shape'], d['stage_us'], d['result_sha1'])"
  ORBX_OCTREE_WIDE=$wide timeout 600 python bench.py --no-cpu --no-knn --no-latency --no-two-callers > gpurun_out/g27_$wide.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/g27_$wide.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'single', round(d['single_lane']['value']), 'e2e', round(d['e2e']['value']), 'cfg1', round(d['config1_752x480_nf1200']['value']), 'cfg3', round(d['config3_1920x1080_nf2000']['value']), 'cfg5', round(d['config5_1280x720_nf1250']['value']))
PY
done
