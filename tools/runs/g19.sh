#!/bin/bash
# N-GPU bench + the upload probe at the same N
set -u
N=${1:-8}
bash tools/bench_n.sh $N
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/h2d_probe.py > gpurun_out/h2d_probe_${N}gpu.json 2>/dev/null
cat gpurun_out/h2d_probe_${N}gpu.json
