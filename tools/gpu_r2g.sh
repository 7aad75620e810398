#!/bin/bash
# Round-2 session G (1 GPU): TMA-staged brick mips + TMA-fed gather under the loopback ranks; mode-N tests; fast-GTAO diagnostics; bench.
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loopback.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_loopback_$tag.log
tail -12 gpurun_out/pytest_loopback_$tag.log
timeout 600 python -m pytest tests/test_gpu_mode_n.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_moden_$tag.log
tail -8 gpurun_out/pytest_moden_$tag.log
timeout 120 python tools/debug_gtao_fast.py 2>&1 | tail -25
for v in 0 1; do
F184_MIPS_BRICKS_LDG=$v timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras --no-overlap > gpurun_out/bench_${tag}_ldg$v.json 2> gpurun_out/bench_${tag}_ldg$v.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_ldg$v.json").read().strip().splitlines()[-1])
    print("bricks LDG=$v (no overlap: solo stage times) c3", round(d["value"],4), d["stages_ms"]); c=d.get("c4_scaling") or {}; print("   c4", c.get("ms_per_frame"), c.get("stages_ms"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${tag}_ldg$v.err").read()[-2000:])
PY
done
