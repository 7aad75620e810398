#!/bin/bash
# Multi-GPU A/B (gpurun --gpus N): the multi-GPU parity test, then the bench line at N with and without the frame overlap.
tag=${1:-ab}; n=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_multi_$tag.log
tail -12 gpurun_out/pytest_multi_$tag.log
for extra in "" "--no-overlap"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline $extra > gpurun_out/bench_${tag}_g$n$extra.json 2> gpurun_out/bench_${tag}_g$n$extra.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_g$n$extra.json").read().strip().splitlines()[-1])
    print("N=$n $extra", round(d["value"],4), "ms/frame", d["stages_ms"], "e2e", round(d["e2e"]["value"],3))
except Exception as e:
    print("N=$n $extra failed", e); print(open("gpurun_out/bench_${tag}_g$n$extra.err").read()[-1500:])
PY
done
