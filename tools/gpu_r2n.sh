#!/bin/bash
tag=${1:-r02v}
mkdir -p gpurun_out
for v in "F184_GATHER_DBG=4"; do
env $v F184_PARITY_DEBUG=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload c4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "$v"; grep "parity debug\|gather dbg" gpurun_out/bench_${tag}.err | cut -c1-250 | head -8
python -c "
import json; d=json.loads(open('gpurun_out/bench_${tag}.json').read().strip().splitlines()[-1]); print({k:v[:60] for k,v in d['parity_vs_1gpu'].items() if k!='checked'}, d['gather']['bytes_per_rank_min_max'])"
done
