"""Peer-read shapes over NVLink (f184_microbench_peer), every rank reading its right-hand neighbour at the same time.
torchrun --nproc-per-node N tools/peer_read_bench.py  ->  one table on rank 0 (GB/s per rank, min over ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from final184_b200 import api as A          # noqa: E402
from final184_b200 import dist as D         # noqa: E402
from final184_b200 import scene as S        # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    sc = S.procedural_scene(seed=1)
    cam = S.fixture_constants("voxel")
    g = D.ShardedVoxelGI(512, 64, 64, shadow_res=64, device=torch.cuda.current_device(), rank=rank, nranks=world, scene=sc, voxel_cam=cam)
    g.connect()
    peer = (rank + 1) % world
    total = 256 << 20
    cases = []
    for copy, depth in ((128, 256), (256, 16), (256, 64), (256, 256), (512, 32), (512, 128), (1024, 16), (1024, 64), (1024, 128), (2048, 16), (2048, 64), (4096, 32), (16384, 2), (16384, 8)):
        for ctas in (148, 296, 592):
            cases.append((0, copy, depth, ctas))
    for copy, depth in ((256, 64), (256, 256), (512, 128), (1024, 64), (2048, 64)):
        for ctas in (148, 296):
            cases.append((1, copy, depth, ctas))
    for ctas in (148, 148 * 4, 148 * 8):
        cases.append((2, 16, 8, ctas))
    rows = []
    for mode, copy, depth, ctas in cases:
        dist.barrier()
        torch.cuda.synchronize()
        try:
            gbs = g.ctx.microbench_peer(peer, mode, copy, depth, ctas, total)
        except A.F184Error as e:
            gbs = -1.0
        t = torch.tensor([gbs], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        rows.append({"mode": mode, "copy_bytes": copy, "depth": depth, "ctas": ctas, "in_flight_per_sm_kb": round(copy * depth * max(1, ctas // 148) / 1024, 1), "gbs_min_over_ranks": round(float(t.item()), 1)})
        if rank == 0:
            print(rows[-1], flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump({"n_gpus": world, "total_bytes": total, "rows": rows}, open(f"gpurun_out/peer_read_g{world}.json", "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
