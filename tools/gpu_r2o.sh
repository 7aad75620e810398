#!/bin/bash
# Round-2 session O (gpurun --gpus 2): what bounds the exchange?  C4 at 2 GPUs, every pass on one stream, with the gather's stores
# switched off (F184_GATHER_DRY: 1 = no level 2/3 stores, 3 = no stores at all) and with half the CTAs.
tag=${1:-r02o2}
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload c4 --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-overlap > gpurun_out/bench_${tag}_$name.json 2> gpurun_out/bench_${tag}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$name.json").read().strip().splitlines()[-1])
    print("C4 N=2 $name:", round(d["value"],4), "ms/frame", d["stages_ms"]); print("    gather", d["gather"]["gbs_per_rank_min_max"], d["gather"]["bytes_per_rank_min_max"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_${tag}_$name.err").read()[-2000:])
PY
}
run base F184_GATHER_DRY=0
run dry1 F184_GATHER_DRY=1
run dry3 F184_GATHER_DRY=3
run ctas148 F184_GATHER_CTAS=148
run all F184_GATHER_ALL=1
run all_dry3 F184_GATHER_ALL=1 F184_GATHER_DRY=3
