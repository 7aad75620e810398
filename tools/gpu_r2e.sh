#!/bin/bash
# Round-2 session E (1 GPU): loopback ranks (eager module loading), bench line.
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loopback.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_loopback_$tag.log
tail -12 gpurun_out/pytest_loopback_$tag.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -c 1500 gpurun_out/bench_${tag}.json; tail -5 gpurun_out/bench_${tag}.err
