"""Worker for tests/test_gpu_multi.py (run under torchrun, one rank per GPU): the slab schedule over peer memory must
give every rank the same volume, texture storage and image rows as one GPU computing the whole frame."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from final184_b200 import api as A, dist as D, scene as S   # noqa: E402
from final184_b200.fixture import frame_inputs               # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, W, H, SH = 64, 160, 96, 256
    sc = S.procedural_scene(seed=1)
    cams = {n: S.fixture_constants(n) for n in ("main", "shadow", "voxel")}
    fi = frame_inputs(sc, cams["main"], cams["shadow"], W, H, SH, 0, cache=False)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    g = D.ShardedVoxelGI(N, W, H, shadow_res=SH, device=local, rank=rank, nranks=world, scene=sc, voxel_cam=cams["voxel"], flags=A.FLAG_GATHER_LINEAR)
    one = A.VoxelGI(N, W, H, A.MODE_NORTHSTAR, shadow_res=SH, device=local)
    one.upload_scene(sc)
    for c in (g.ctx, one):
        for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material"), (A.SLOT_SHADOW, "shadow")):
            c.upload(slot, fi[key])
    g.connect()
    bad = []
    for frame in range(3):
        if frame == 2:      # fewer triangles: bricks that became empty must be cleared on every rank
            half = sc.n_tris // 2
            f0, cnt = g.tri_range
            g.ctx.set_triangle_range(f0, max(0, min(f0 + cnt, half) - f0))
            one.set_triangle_range(0, half)
        g.frame(cams["voxel"], k)
        one.voxelize(cams["voxel"]); one.inject(k); one.build_mips(); one.trace_indirect(k)
        g.ctx.sync(); one.sync()
        dist.barrier()       # nobody starts the next frame's atomics while a peer still reads this frame back
        if not np.array_equal(g.ctx.readback(A.SLOT_RADIANCE), one.readback(A.SLOT_RADIANCE)): bad.append(f"frame {frame}: radiance")
        if not np.array_equal(g.ctx.readback(A.SLOT_MIPS), one.readback(A.SLOT_MIPS)): bad.append(f"frame {frame}: mips")
        if not np.array_equal(g.ctx.read_array(-1, 0, N), one.read_array(-1, 0, N)): bad.append(f"frame {frame}: level-0 array")
        m, lvl = N // 2, 0
        while m >= 1:
            for d in range(6):
                if not np.array_equal(g.ctx.read_array(d, lvl, m), one.read_array(d, lvl, m)): bad.append(f"frame {frame}: array dir {d} level {lvl + 1}")
            m //= 2; lvl += 1
        m = g.own_rows_mask()
        a, b = g.ctx.readback(A.SLOT_INDIRECT_OUT)[m], one.readback(A.SLOT_INDIRECT_OUT)[m]
        if not np.array_equal(a.view(np.uint16), b.view(np.uint16)): bad.append(f"frame {frame}: own image tile rows")
        dist.barrier()
    # back to back, no host sync between frames: the accumulation of frame f+1 runs on an internal stream beside the gather and
    # the cone trace of frame f (f184_voxelize_accumulate, one NVLink box).  The triangle range changes every frame, so every
    # frame's volume differs and emptied bricks must be cleared; each rank's image rows must equal the one-GPU frames.
    T = sc.n_tris
    f0, cnt = g.tri_range
    own = lambda lo, hi: (max(f0, lo), max(0, min(f0 + cnt, hi) - max(f0, lo)))       # this rank's share of triangles [lo, hi)
    spans = [(0, T), (0, T // 2), (T // 3, T), (0, T), (T // 2, T), (0, T // 4)]
    want = []
    for lo, hi in spans:
        one.set_triangle_range(lo, hi - lo)
        one.voxelize(cams["voxel"]); one.inject(k); one.build_mips(); one.trace_indirect(k)
        want.append(one.readback(A.SLOT_INDIRECT_OUT).copy())
    nbytes = g.ctx.image_info(A.SLOT_INDIRECT_OUT).size_bytes
    hosts = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in spans]
    dist.barrier()
    for i, (lo, hi) in enumerate(spans):
        g.ctx.set_triangle_range(*own(lo, hi))
        g.frame(cams["voxel"], k)
        g.ctx.readback_async_ptr(A.SLOT_INDIRECT_OUT, hosts[i].data_ptr(), nbytes)
        if i:
            g.ctx.readback_wait(1)
    g.ctx.sync()
    dist.barrier()
    m = g.own_rows_mask()
    for i in range(len(spans)):
        got = hosts[i].numpy().view(np.uint16).reshape(H, W, 4)
        if not np.array_equal(got[m], want[i].view(np.uint16).reshape(H, W, 4)[m]): bad.append(f"back-to-back frame {i}: own image tile rows")
    if np.array_equal(want[0], want[1]): bad.append("back-to-back frames do not differ")
    g.ctx.set_triangle_range(f0, cnt)
    frags = torch.tensor([float(g.ctx.counter(A.COUNTER_FRAGMENTS))], device=f"cuda:{local}")
    dist.all_reduce(frags)
    # BASELINE configs[4]: a probe batch partitioned by whole views; the volume comes from the slab schedule, each rank traces
    # its own views (f184_trace_views) — they must equal the same views traced alone on one GPU
    import test_probe_views as P
    one.set_triangle_range(0, 0xffffffff)
    nv, vs = 2 * world + 1, 64
    cams_, views, per_view, shadow, ks = P.batch_inputs(sc, nv, vs, SH, stride=7)
    pb = D.ProbeBatch(N, vs, nv, shadow_res=SH, device=local, rank=rank, nranks=world, scene=sc, voxel_cam=cams_["voxel"], volume_mode="slab")
    pb.gi.connect()
    pb.upload_views(per_view, shadow)
    pb.frame(cams_["voxel"], ks)
    pb.ctx.sync()
    own = pb.own_views()
    one.upload(A.SLOT_SHADOW, shadow)
    one.voxelize(cams_["voxel"]); one.inject(ks[0]); one.build_mips()
    small = A.VoxelGI(N, vs, vs, A.MODE_NORTHSTAR, shadow_res=SH, device=local)
    small.upload_scene(sc); small.upload(A.SLOT_SHADOW, shadow)
    small.voxelize(cams_["voxel"]); small.inject(ks[0]); small.build_mips()
    for v, img in own:
        for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material")):
            small.upload(slot, per_view[v][key])
        small.trace_indirect(ks[v])
        if not np.array_equal(small.readback(A.SLOT_INDIRECT_OUT).view(np.uint16), img.view(np.uint16)): bad.append(f"probe view {v}")
    if [v for v, _ in own] != list(range(*D.view_ranges(nv, world)[rank])): bad.append("probe view partition")
    dist.barrier()
    pb.close(); small.close()
    # the real scene at a real size: Sponza 256^3 (BASELINE configs[1]'s volume), two frames back to back — 0.7 M fragments through
    # peer atomics, ~8 k bricks through the gather — against one GPU
    if S.sponza_available():
        g.close(); one.close()
        sp = S.load_sponza()
        N2, W2, H2 = 256, 320, 184
        fi2 = frame_inputs(sp, cams["main"], cams["shadow"], W2, H2, 1024, 0, cache=False)
        k2 = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W2, H2, 0, True)
        g = D.ShardedVoxelGI(N2, W2, H2, shadow_res=1024, device=local, rank=rank, nranks=world, scene=sp, voxel_cam=cams["voxel"], flags=A.FLAG_GATHER_LINEAR)
        one = A.VoxelGI(N2, W2, H2, A.MODE_NORTHSTAR, shadow_res=1024, device=local)
        one.upload_scene(sp)
        for c in (g.ctx, one):
            for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material"), (A.SLOT_SHADOW, "shadow")):
                c.upload(slot, fi2[key])
        g.connect()
        for frame in range(2):
            g.frame(cams["voxel"], k2)
        for frame in range(2):
            one.voxelize(cams["voxel"]); one.inject(k2); one.build_mips(); one.trace_indirect(k2)
        g.ctx.sync(); one.sync()
        dist.barrier()
        if not np.array_equal(g.ctx.readback(A.SLOT_RADIANCE), one.readback(A.SLOT_RADIANCE)): bad.append("sponza 256: radiance")
        if not np.array_equal(g.ctx.readback(A.SLOT_MIPS), one.readback(A.SLOT_MIPS)): bad.append("sponza 256: mips")
        m2 = g.own_rows_mask()
        if not np.array_equal(g.ctx.readback(A.SLOT_INDIRECT_OUT)[m2].view(np.uint16), one.readback(A.SLOT_INDIRECT_OUT)[m2].view(np.uint16)): bad.append("sponza 256: own image tile rows")
        fr = torch.tensor([float(g.ctx.counter(A.COUNTER_FRAGMENTS))], device=f"cuda:{local}")
        dist.all_reduce(fr)
        if int(fr.item()) != one.counter(A.COUNTER_FRAGMENTS): bad.append(f"sponza 256: fragments {int(fr.item())} vs {one.counter(A.COUNTER_FRAGMENTS)}")
        dist.barrier()
    print(f"rank {rank}: {'OK' if not bad else 'MISMATCH ' + '; '.join(bad[:6])} (fragments over ranks {int(frags.item())})", flush=True)
    g.close(); one.close()
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
