#!/bin/bash
# Multi-GPU session (gpurun --gpus N): parity tests (incl. tests/test_gpu_multi.py), then the bench line at every power of two up to N.
tag=${1:-multi}; n=${2:-2}; shift; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/pytest_gpu_$tag.log
tail -6 gpurun_out/pytest_gpu_$tag.log
g=1
while [ $g -le $n ]; do
  if [ $g -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_${tag}_g$g.json 2> gpurun_out/bench_${tag}_g$g.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $g --steps 50 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_${tag}_g$g.json 2> gpurun_out/bench_${tag}_g$g.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_g$g.json").read().strip().splitlines()[-1])
    print("N=$g", round(d["value"],4), "ms/frame", d["stages_ms"], "comm", d.get("comm_ms"), "e2e", round(d["e2e"]["value"],3))
except Exception as e:
    print("N=$g failed", e); print(open("gpurun_out/bench_${tag}_g$g.err").read()[-1500:])
PY
  g=$((g*2))
done
