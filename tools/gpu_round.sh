#!/bin/bash
# One GPU session: parity tests, the bench line, the ncu launch list of the same command.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [bench args]
tag=${1:-r01}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$tag.log
python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 3000 gpurun_out/bench_$tag.json
tail -5 gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/bench_under_ncu_$tag.log 2>&1
tail -3 gpurun_out/pytest_gpu_$tag.log
