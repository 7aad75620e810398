#!/bin/bash
# GPU session for the reference-faithful mode at BASELINE configs[0]: parity tests, C1 stage table, full ncu capture of the mode R kernels.
tag=${1:-c1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$tag.log
tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 300 python tools/c1_frames.py 20 | tee gpurun_out/c1_$tag.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_voxelize_r|k_resolve_r|k_trace_r|k_gtao|k_blur|k_lighting|k_composite' \
    --launch-skip 27 -c 9 -o gpurun_out/full_$tag -f python tools/c1_frames.py 1 > gpurun_out/ncu_full_$tag.log 2>&1
ncu -i gpurun_out/full_$tag.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/full_${tag}_raw.csv
timeout 300 python bench.py --no-cpu-baseline --steps 50 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print(round(d["value"],4), d["stages_ms"], "e2e", round(d["e2e"]["value"],4), d["secondary_ms"]); print(d["c1_reference_mode"]["gpu"])
PY
