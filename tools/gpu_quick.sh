#!/bin/bash
# quick GPU check of a kernel change: parity tests, then the bench line (no CPU leg) under each setting of one A/B env knob
# usage (under gpurun): bash tools/gpu_quick.sh <tag> <ENVVAR> <value> [<value> ...]
tag=${1:-q}; var=${2:-F184_NONE}; shift; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$tag.log
tail -4 gpurun_out/pytest_gpu_$tag.log
for v in "$@"; do
  for extra in "" "--no-overlap"; do
    env $var=$v timeout 300 python bench.py --no-cpu-baseline --steps 50 --warmup 5 $extra > gpurun_out/bench_${tag}_$v$extra.json 2> gpurun_out/bench_${tag}_$v$extra.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}_$v$extra.json").read().strip().splitlines()[-1])
print("$var=$v $extra", round(d["value"],4), d["stages_ms"], "e2e", round(d["e2e"]["value"],4), "tex", round(d["other_bounds"]["trace_tex"]["frac"],3))
PY
  done
done
