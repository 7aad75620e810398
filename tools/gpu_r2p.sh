#!/bin/bash
# Round-2 final session (1 GPU): the whole GPU suite, smoke(), the bench line, the ncu launch list of the same command and one
# --set full capture of every frame kernel at C3.
tag=${1:-r02final}
mkdir -p gpurun_out
timeout 480 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_$tag.log
tail -6 gpurun_out/pytest_gpu_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -c 700 gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-c4 > gpurun_out/launches_${tag}.log 2>&1
python tools/launch_summary.py gpurun_out/launches_${tag}.csv | tail -30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_trace_n|k_voxelize_raster|k_voxelize_setup|k_normalise_n|k_mips_bricks|k_inject_n|k_mips_tail|k_brick_compact' \
    --launch-skip 24 -c 8 -o gpurun_out/full_${tag}_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-c4 --no-overlap > gpurun_out/ncu_full_${tag}_c3.log 2>&1
tail -2 gpurun_out/ncu_full_${tag}_c3.log
ncu -i gpurun_out/full_${tag}_c3.ncu-rep --page raw --csv > gpurun_out/full_${tag}_c3_raw.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/full_${tag}_c3_raw.csv
