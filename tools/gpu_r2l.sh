#!/bin/bash
# Session L (gpurun --gpus 2): shared-memory carve-out of the tracer A/B (1 GPU, pipelined C3 + C4), then the 2-GPU line.
tag=${1:-r02s}
mkdir -p gpurun_out
for co in -1 44 70; do
  F184_TRACE_CARVEOUT=$co timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_${tag}_g1_co$co.json 2> gpurun_out/bench_${tag}_g1_co$co.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_g1_co$co.json").read().strip().splitlines()[-1])
    c=d.get("c4_scaling") or {}
    print("N=1 carveout=$co: c3", round(d["value"],4), d["stages_ms"], "| c4", c.get("ms_per_frame"), c.get("stages_ms"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_${tag}_g1_co$co.err").read()[-2000:])
PY
done
for co in -1 44; do
  F184_TRACE_CARVEOUT=$co timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_g2_co$co.json 2> gpurun_out/bench_${tag}_g2_co$co.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_g2_co$co.json").read().strip().splitlines()[-1])
    print("N=2 carveout=$co c3", round(d["value"],4), "ms/frame", d["stages_ms"], "e2e", round(d["e2e"]["value"],3))
    g=d.get("gather") or {}; print("   gather", g.get("gbs_per_rank_min_max"), g.get("bytes_per_rank_min_max")); print("   parity", {k:v for k,v in (d.get("parity_vs_1gpu") or {}).items() if k!="checked"})
    c=d.get("c4_scaling") or {}
    print("   c4", c.get("ms_per_frame"), c.get("stages_ms")); g=c.get("gather") or {}; print("   c4 gather", g.get("gbs_per_rank_min_max"), g.get("bytes_per_rank_min_max")); print("   c4 parity", {k:v for k,v in (c.get("parity_vs_1gpu") or {}).items() if k!="checked"})
except Exception as e:
    print("N=2 failed", e); print(open("gpurun_out/bench_${tag}_g2_co$co.err").read()[-3000:])
PY
done
