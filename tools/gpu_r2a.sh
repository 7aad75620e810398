#!/bin/bash
# Round-2 session A (1 GPU): where does the time go at BASELINE configs[3] (8 x tiled Sponza, 1024^3, 4K)?
# bench line (overlap on / off) + launch list + ncu --set full of every frame kernel at C4.
tag=${1:-r02e}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 420 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_c4.json 2> gpurun_out/bench_${tag}_c4.err
tail -c 600 gpurun_out/bench_${tag}_c4.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline --no-overlap > gpurun_out/bench_${tag}_c4_nooverlap.json 2> gpurun_out/bench_${tag}_c4_nooverlap.err
python - <<PY
import json
for f in ("gpurun_out/bench_${tag}_c4.json", "gpurun_out/bench_${tag}_c4_nooverlap.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 3), d["stages_ms"], d["counters"])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_trace_n|k_voxelize_raster|k_voxelize_setup|k_normalise_n|k_mips_bricks|k_inject_n|k_mips_tail|k_brick_compact' \
    --launch-skip 24 -c 8 -o gpurun_out/full_${tag}_c4 -f python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-overlap > gpurun_out/ncu_full_${tag}_c4.log 2>&1
tail -3 gpurun_out/ncu_full_${tag}_c4.log
ncu -i gpurun_out/full_${tag}_c4.ncu-rep --page raw --csv > gpurun_out/full_${tag}_c4_raw.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/full_${tag}_c4_raw.csv
