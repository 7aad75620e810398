#!/bin/bash
# Scaling session (gpurun --gpus 8): the bench line at the given rank counts, overlap on (default) and, at the largest, off.
tag=${1:-scale}; shift
mkdir -p gpurun_out
for g in "$@"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $g --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_${tag}_g$g.json 2> gpurun_out/bench_${tag}_g$g.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_g$g.json").read().strip().splitlines()[-1])
    print("N=$g", round(d["value"],4), "ms/frame", d["stages_ms"], "e2e", round(d["e2e"]["value"],3)); print("   min/max over ranks", d.get("stages_ms_min_max_over_ranks"))
except Exception as e:
    print("N=$g failed", e); print(open("gpurun_out/bench_${tag}_g$g.err").read()[-2500:])
PY
done
