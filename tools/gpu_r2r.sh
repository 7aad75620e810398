#!/bin/bash
# Round-2 last session (1 GPU, ~3 minutes): the static / dynamic split on the device, what it buys, and the default path re-checked.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32 CUDA_MODULE_LOADING=EAGER
timeout 60 python tests/loopback_worker.py static > gpurun_out/static_worker.log 2>&1; echo "static worker rc=$?"; tail -3 gpurun_out/static_worker.log
unset CUDA_DEVICE_MAX_CONNECTIONS CUDA_MODULE_LOADING
timeout 50 python tools/static_cache_bench.py > gpurun_out/static_cache_bench.json 2> gpurun_out/static_cache_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/static_cache_bench.json; tail -3 gpurun_out/static_cache_bench.err
timeout 40 python -m pytest tests/test_gpu_mode_n.py -m gpu -q -x 2>&1 | tail -4
timeout 60 python -m pytest tests/test_gpu_loopback.py -m gpu -q -x -k "two_loopback or sponza_256 or pipelined" 2>&1 | tail -4
