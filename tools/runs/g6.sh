#!/bin/bash
# two GPUs: the multi-GPU tests, the C harness under torch.distributed.run, the bench at N = 2
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q -k "two or sharded or knn2" > gpurun_out/g6_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g6_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/g6_bench_2gpu.json 2> gpurun_out/g6_bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/g6_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g6_bench_2gpu.json').read().strip().splitlines()[-1])
def short(x):
    if isinstance(x, dict):
        return {k: short(v) for k, v in x.items() if k not in ("note","how","api","workload","sample","l2","lanes")}
    if isinstance(x, float): return round(x, 4)
    return x
for k in ("value","ms_per_step","single_lane","sustained","e2e","config3_1920x1080_nf2000","config5_1280x720_nf1250"):
    print(k, json.dumps(short(d.get(k)))[:900])
print("cfg4", json.dumps(short(d["hamming"]["cfg4"]))[:1500])
PY
