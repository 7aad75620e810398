#!/bin/bash
# Round-2 session I (gpurun --gpus N): fragment queues.  Loopback parity on GPU 0, multi-GPU parity, then C3 + C4 at N.
tag=${1:-r02n}; n=${2:-2}
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_loopback.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_loopback_$tag.log
tail -6 gpurun_out/pytest_loopback_$tag.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | tail -30 > gpurun_out/pytest_multi_$tag.log
tail -5 gpurun_out/pytest_multi_$tag.log
for extra in "" "--no-overlap"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline $extra > gpurun_out/bench_${tag}_g$n$extra.json 2> gpurun_out/bench_${tag}_g$n$extra.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_g$n$extra.json").read().strip().splitlines()[-1])
    print("N=$n $extra c3", round(d["value"],4), "ms/frame", d["stages_ms"], "e2e", round(d["e2e"]["value"],3))
    print("   min/max", d.get("stages_ms_min_max_over_ranks")); print("   gather", (d.get("gather") or {}).get("gbs_per_rank_min_max")); print("   parity", {k:v for k,v in (d.get("parity_vs_1gpu") or {}).items() if k!="checked"})
    c=d.get("c4_scaling") or {}
    print("   c4", c.get("ms_per_frame"), c.get("stages_ms")); print("   c4 min/max", c.get("stages_ms_min_max_over_ranks")); print("   c4 gather", (c.get("gather") or {}).get("gbs_per_rank_min_max")); print("   c4 parity", {k:v for k,v in (c.get("parity_vs_1gpu") or {}).items() if k!="checked"})
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/bench_${tag}_g$n$extra.err").read()[-3000:])
PY
done
