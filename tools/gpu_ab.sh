#!/bin/bash
# quick A/B on the GPU: bench with and without the frame overlap (no tests, no ncu)
tag=${1:-ab}; shift
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print("overlap   ", round(d["value"],4), d["stages_ms"], "e2e", round(d["e2e"]["value"],4), d["clocks"])
PY
tail -3 gpurun_out/bench_$tag.err
timeout 300 python bench.py --no-overlap --no-cpu-baseline "$@" > gpurun_out/bench_${tag}_nooverlap.json 2> gpurun_out/bench_${tag}_nooverlap.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}_nooverlap.json").read().strip().splitlines()[-1])
print("no-overlap", round(d["value"],4), d["stages_ms"], "e2e", round(d["e2e"]["value"],4), d["clocks"])
PY
