#!/bin/bash
# Round-2 session H (gpurun --gpus 2): why is the exchange slow inside the frame pipeline?  C4 at 2 GPUs, pipeline on/off x gather TMA/LDG.
tag=${1:-r02m}
mkdir -p gpurun_out
nvidia-smi topo -m | head -6
for ldg in 0 1; do for extra in "" "--no-overlap"; do
  F184_GATHER_LDG=$ldg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload c4 --steps 20 --warmup 3 --no-cpu-baseline --no-extras $extra > gpurun_out/bench_${tag}_ldg${ldg}${extra}.json 2> gpurun_out/bench_${tag}_ldg${ldg}${extra}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_ldg${ldg}${extra}.json").read().strip().splitlines()[-1])
    print("C4 N=2 gather_ldg=$ldg $extra:", round(d["value"],4), "ms/frame", d["stages_ms"]); print("    gather", d["gather"]["gbs_per_rank_min_max"], d["gather"]["bytes_per_rank_min_max"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_${tag}_ldg${ldg}${extra}.err").read()[-2000:])
PY
done; done
