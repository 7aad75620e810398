#!/bin/bash
# Round-2 session B (1 GPU): the frame pipeline — full GPU test-suite (incl. the loopback ranks), then the bench line with and without it.
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_$tag.log
tail -12 gpurun_out/pytest_gpu_$tag.log
for extra in "" "--no-overlap"; do
  timeout 400 python bench.py --no-cpu-baseline --steps 50 --warmup 5 $extra > gpurun_out/bench_${tag}$extra.json 2> gpurun_out/bench_${tag}$extra.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}$extra.json").read().strip().splitlines()[-1])
    print("c3 $extra", round(d["value"],4), d["stages_ms"], "e2e", round(d["e2e"]["value"],4))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${tag}$extra.err").read()[-2000:])
PY
done
timeout 400 python bench.py --workload c4 --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_${tag}_c4.json 2> gpurun_out/bench_${tag}_c4.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_c4.json").read().strip().splitlines()[-1])
    print("c4", round(d["value"],4), d["stages_ms"], "e2e", round(d["e2e"]["value"],4))
except Exception as e:
    print("bench c4 failed", e); print(open("gpurun_out/bench_${tag}_c4.err").read()[-2000:])
PY
