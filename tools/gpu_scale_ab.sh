#!/bin/bash
# 8-GPU A/B of the frame-overlap placement: default, accumulation behind the gather, no overlap.
tag=${1:-ab8}; g=${2:-8}
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $g --steps 100 --warmup 10 --no-cpu-baseline $EXTRA > gpurun_out/bench_${tag}_$name.json 2> gpurun_out/bench_${tag}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],4), "ms/frame", d["stages_ms"], "e2e", round(d["e2e"]["value"],3)); print("   min/max", d.get("stages_ms_min_max_over_ranks"))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/bench_${tag}_$name.err").read()[-1500:])
PY
}
EXTRA="" run default F184_NONE=0
EXTRA="" run vox_after_gather F184_VOX_AFTER_GATHER=1
EXTRA="--no-overlap" run no_overlap F184_NONE=0
