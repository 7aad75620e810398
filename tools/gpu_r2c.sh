#!/bin/bash
# Round-2 session C (1 GPU): GPU test-suite (loopback ranks first), then the new bench line.
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loopback.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_loopback_$tag.log
tail -25 gpurun_out/pytest_loopback_$tag.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_loopback.py 2>&1 | tail -30 > gpurun_out/pytest_gpu_$tag.log
tail -8 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -c 6000 gpurun_out/bench_${tag}.json; tail -5 gpurun_out/bench_${tag}.err
