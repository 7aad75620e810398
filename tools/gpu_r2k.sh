#!/bin/bash
# 1 GPU: loopback suite only (fast check of the multi-rank code on one device)
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loopback.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_loopback_$tag.log
tail -15 gpurun_out/pytest_loopback_$tag.log
