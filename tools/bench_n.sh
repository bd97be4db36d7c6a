#!/bin/bash
# bench.py at N GPUs (N = $1) -> gpurun_out/bench_r02_${N}gpu.json (+ the reference arm at N = 1)
set -u
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err; echo "bench rc=$?"
  timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r02_reference.json 2>/dev/null
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err; echo "bench rc=$?"
fi
tail -2 gpurun_out/bench_r02_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02_${N}gpu.json').read().strip().splitlines()[-1])
h=d['hamming']
print('N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'single', round(d['single_lane']['value']), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'ceiling', round(d['e2e']['h2d_ceiling']['value']), 'whole', round(d['roofline']['whole_step']['frac'],4))
print('cfg4', h['cfg4'].get('pairs_per_s'), h['cfg4'].get('ms_per_query_batch'), h['cfg4'].get('allgather_us'), h['cfg4'].get('known_answers_ok'), h['cfg4'].get('route','')[:40], '| shard', h['value'], h['popc_backend']['value'])
for k in ('config1_752x480_nf1200','config3_1920x1080_nf2000','config5_1280x720_nf1250'):
    print(k, round(d[k]['value']), round(d[k]['whole_step_hbm_frac'],4))
print('clocks', d['clocks'])
PY
