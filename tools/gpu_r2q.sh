#!/bin/bash
# Round-2 final scaling point (gpurun --gpus N): the bench line at N ranks, pipelined, as the driver launches it.
tag=${1:-r02final}; n=${2:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_g$n.json 2> gpurun_out/bench_${tag}_g$n.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_g$n.json").read().strip().splitlines()[-1])
    print("N=$n c3", round(d["value"],4), "ms/frame", d["stages_ms"], "e2e", round(d["e2e"]["value"],3))
    g=d.get("gather") or {}; print("   gather", g.get("gbs_per_rank_min_max"), g.get("bytes_per_rank_min_max")); print("   parity", {k:v for k,v in (d.get("parity_vs_1gpu") or {}).items() if k!="checked"})
    c=d.get("c4_scaling") or {}
    print("   c4", c.get("ms_per_frame"), c.get("stages_ms")); print("   c4 parity", {k:v for k,v in (c.get("parity_vs_1gpu") or {}).items() if k!="checked"})
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/bench_${tag}_g$n.err").read()[-3000:])
PY
