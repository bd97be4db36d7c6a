#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/stage_sweep.py "" > gpurun_out/g15_sweep.jsonl 2>&1; cut -c1-300 gpurun_out/g15_sweep.jsonl
timeout 100 python tools/whatif.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --no-cpu --no-knn --no-euroc > gpurun_out/g15_bench.json 2>gpurun_out/g15_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/g15_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['single_lane']['value'], d['sustained']['value'], d['e2e']['value'], d['e2e']['h2d_ceiling']['value'])
print(json.dumps(d['latency'])[:1200])
print(d['roofline']['stage_ms'])
PY
