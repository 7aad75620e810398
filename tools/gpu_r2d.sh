#!/bin/bash
# Round-2 session D (1 GPU): loopback ranks + the new full-size oracle parity tests.
tag=${1:-r02h}
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 900 python -m pytest tests/test_gpu_loopback.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_loopback_$tag.log
tail -12 gpurun_out/pytest_loopback_$tag.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/pytest_fullsize_$tag.log
tail -25 gpurun_out/pytest_fullsize_$tag.log
