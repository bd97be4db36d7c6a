#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g5_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g5_pytest.log
timeout 900 python bench.py > gpurun_out/g5_bench.json 2> gpurun_out/g5_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/g5_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g5_bench.json').read().strip().splitlines()[-1])
def short(x, depth=0):
    if isinstance(x, dict):
        return {k: short(v, depth+1) for k, v in x.items() if k not in ("note","how","api","workload","sample","l2","lanes","route")}
    if isinstance(x, float): return round(x, 4)
    return x
for k in ("value","ms_per_step","single_lane","sustained","e2e","latency","config1_752x480_nf1200","config3_1920x1080_nf2000","config5_1280x720_nf1250"):
    print(k, json.dumps(short(d.get(k)))[:900])
print("hamming", json.dumps(short(d["hamming"]))[:1800])
print("roofline", json.dumps(short(d["roofline"]))[:900])
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-600
