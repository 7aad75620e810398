#!/bin/bash
# Round-2 session J (gpurun --gpus 8): the scaling run — the bench line at 8 ranks (C3 headline + c4_scaling), pipelined and with every pass on one stream (solo stage times).
tag=${1:-r02q}
mkdir -p gpurun_out
run() { # n extra suffix
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $1 --steps 50 --warmup 5 --no-cpu-baseline $2 > gpurun_out/bench_${tag}_g$1$3.json 2> gpurun_out/bench_${tag}_g$1$3.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_g$1$3.json").read().strip().splitlines()[-1])
    print("N=$1 $2 c3", round(d["value"],4), "ms/frame", d["stages_ms"], "e2e", round(d["e2e"]["value"],3))
    print("   min/max", d.get("stages_ms_min_max_over_ranks")); g=d.get("gather") or {}; print("   gather", g.get("gbs_per_rank_min_max"), g.get("bytes_per_rank_min_max")); print("   parity", {k:v for k,v in (d.get("parity_vs_1gpu") or {}).items() if k!="checked"})
    c=d.get("c4_scaling") or {}
    print("   c4", c.get("ms_per_frame"), c.get("stages_ms")); print("   c4 min/max", c.get("stages_ms_min_max_over_ranks")); g=c.get("gather") or {}; print("   c4 gather", g.get("gbs_per_rank_min_max"), g.get("bytes_per_rank_min_max")); print("   c4 parity", {k:v for k,v in (c.get("parity_vs_1gpu") or {}).items() if k!="checked"})
except Exception as e:
    print("N=$1 failed", e); print(open("gpurun_out/bench_${tag}_g$1$3.err").read()[-3000:])
PY
}
run ${2:-8} "--no-overlap" "_nooverlap"
run ${2:-8} "" ""
