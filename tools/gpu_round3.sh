#!/bin/bash
# GPU session: parity tests, bench line (+ no-overlap A/B), launch list, one full ncu capture of the frame's kernels.
tag=${1:-r01}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$tag.log
tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 3500 gpurun_out/bench_$tag.json
tail -5 gpurun_out/bench_$tag.err
timeout 300 python bench.py --no-overlap --no-cpu-baseline "$@" > gpurun_out/bench_${tag}_nooverlap.json 2> gpurun_out/bench_${tag}_nooverlap.err
tail -c 1500 gpurun_out/bench_${tag}_nooverlap.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_under_ncu_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace_n|k_voxelize_raster|k_voxelize_setup|k_normalise_n|k_mips_bricks|k_inject_n' \
    --launch-skip 30 -c 8 -o gpurun_out/full_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-overlap "$@" > gpurun_out/ncu_full_$tag.log 2>&1
ncu -i gpurun_out/full_$tag.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/full_${tag}_raw.csv
