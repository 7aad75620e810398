#!/bin/bash
# Round-2 session F (gpurun --gpus N): real multi-GPU — parity over NVLink (tests/test_gpu_multi.py), the new mode-N tests, bench at every power of two up to N.
tag=${1:-r02j}; n=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_mode_n.py -m gpu -q -s 2>&1 | tail -30 > gpurun_out/pytest_multi_$tag.log
tail -12 gpurun_out/pytest_multi_$tag.log
g=1
while [ $g -le $n ]; do
  if [ $g -eq 1 ]; then
    timeout 500 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_g$g.json 2> gpurun_out/bench_${tag}_g$g.err
  else
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $g --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_g$g.json 2> gpurun_out/bench_${tag}_g$g.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_g$g.json").read().strip().splitlines()[-1])
    print("N=$g c3", round(d["value"],4), "ms/frame", d["stages_ms"], "e2e", round(d["e2e"]["value"],3))
    print("   min/max", d.get("stages_ms_min_max_over_ranks")); print("   gather", d.get("gather")); print("   parity", d.get("parity_vs_1gpu"))
    c=d.get("c4_scaling") or {}
    print("   c4", c.get("ms_per_frame"), c.get("stages_ms")); print("   c4 min/max", c.get("stages_ms_min_max_over_ranks")); print("   c4 gather", c.get("gather")); print("   c4 parity", c.get("parity_vs_1gpu"))
    print("   secondary", d.get("secondary_ms"))
except Exception as e:
    print("N=$g failed", e); print(open("gpurun_out/bench_${tag}_g$g.err").read()[-3000:])
PY
  g=$((g*2))
done
