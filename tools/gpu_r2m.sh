#!/bin/bash
# Session M (gpurun --gpus 2): which part loses a fragment at C4?  parity of the C4 frame under three fragment paths.
tag=${1:-r02u}
mkdir -p gpurun_out
run() {
  env $1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload c4 --steps 10 --warmup 2 --no-cpu-baseline $3 > gpurun_out/bench_${tag}_$2.json 2> gpurun_out/bench_${tag}_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$2.json").read().strip().splitlines()[-1])
    print("$1 $3:", round(d["value"],4), d["stages_ms"]); print("    parity", {k:v for k,v in (d.get("parity_vs_1gpu") or {}).items() if k!="checked"}, "frags", d["counters"]["fragments"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/bench_${tag}_$2.err").read()[-2000:])
PY
}
run F184_X=0 default ""
run F184_FRAG_QUEUE_RECORDS=32 remote_atomics ""
run F184_FRAG_NO_AGG=1 no_agg ""
run F184_GATHER_ALL=1 gather_all ""
run F184_X=0 default_nooverlap "--no-overlap"
