"""Host-side logic of the multi-GPU schedule (final184_b200/dist.py), on CPU: the partitioning helpers, and the
sharded frame at world_size 2 over the gloo backend with the CPU oracle standing in for the device library —
triangle ranges + summed partial accumulators + row bands must reproduce the single-process frame bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from final184_b200 import api as A
from final184_b200 import dist as D
from final184_b200 import scene as S
from final184_b200.fixture import frame_inputs

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(REPO, "oracle", "_build", "libf184_oracle.so")
W, H, N, SH = 64, 40, 32, 128


def test_triangle_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    w = rng.pareto(1.5, 10000) + 1.0                   # heavy tail, like Sponza's triangle areas
    for n in (1, 2, 3, 8):
        r = D.triangle_ranges(w, n)
        assert r[0][0] == 0 and sum(c for _, c in r) == len(w)
        assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(n - 1))
        loads = [w[f:f + c].sum() for f, c in r]
        assert max(loads) <= w.sum() / n + w.max() + 1e-9
    assert D.triangle_ranges(np.ones(0), 4) == [(0, 0)] * 4
    assert D.triangle_ranges(np.ones(3), 8)[-1][0] == 3


def test_slab_row_view_ranges():
    assert D.slab_ranges(512, 8) == [(64 * r, 64 * (r + 1)) for r in range(8)]
    assert D.slab_ranges(4, 8)[3] == (1, 2) and D.slab_ranges(4, 8)[0] == (0, 0)       # empty slabs once n < nranks
    for h, n in ((2160, 8), (1080, 8), (36, 4), (7, 2)):
        rr = D.row_ranges(h, n)
        assert rr[0][0] == 0 and rr[-1][1] == h and all(rr[i][1] == rr[i + 1][0] for i in range(n - 1))
        assert all(y0 % 8 == 0 for y0, _ in rr)
    assert D.view_ranges(64, 8) == [(8 * r, 8 * r + 8) for r in range(8)]
    # chunk lists: every chunk exactly once, sorted, loads within a chunk of each other even for a heavy-tailed cost
    rng = np.random.default_rng(5)
    wts = rng.pareto(1.2, 262267) + 4.0
    parts = D.triangle_chunks(wts, 8)
    allc = np.concatenate(parts)
    assert sorted(allc.tolist()) == list(range((262267 + 127) // 128)) and all((np.diff(p) > 0).all() for p in parts)
    cost = np.add.reduceat(wts, np.arange(0, len(wts), 128))
    loads = np.array([cost[p].sum() for p in parts])
    assert loads.max() - loads.min() <= cost.max() + 1e-9 and loads.max() / loads.mean() < 1.05
    assert [len(p) for p in D.triangle_chunks(wts, 1)] == [(262267 + 127) // 128]
    assert [D.brick_owner(0, 0, z, 4) for z in (0, 7, 8, 31, 32, 511)] == [0, 0, 1, 3, 0, 3]
    assert D.brick_owner(8, 16, 24, 4) == 2 and D.brick_owner(7, 7, 7, 8) == 0
    # every axis-aligned sheet of bricks is dealt evenly
    b = np.arange(64)
    for fixed in range(3):
        own = (b[:, None] + b[None, :] + 5) % 8
        assert (np.bincount(own.ravel(), minlength=8) == 512).all()


def test_triangle_weights_track_projected_area(proc_scene, cams):
    w = D.triangle_weights(proc_scene, cams["voxel"], 128)
    assert w.shape == (proc_scene.n_tris,) and (w >= 4.0).all()
    big = D.triangle_weights(proc_scene, cams["voxel"], 256)
    assert big.sum() > 2.5 * w.sum() - 4.0 * len(w) * 3     # columns grow ~4x with the grid edge doubling


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        olib = A.Library(ORACLE, "f184o_", product=False)
        sc = S.procedural_scene(seed=1)
        cams = {n: S.fixture_constants(n) for n in ("main", "shadow", "voxel")}
        fi = frame_inputs(sc, cams["main"], cams["shadow"], W, H, SH, 0, cache=False)
        g = D.ShardedVoxelGI(N, W, H, shadow_res=SH, rank=rank, nranks=world, scene=sc, mode="host", lib=olib, voxel_cam=cams["voxel"])
        for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material"), (A.SLOT_SHADOW, "shadow")):
            g.ctx.upload(slot, fi[key])
        k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
        g.frame(cams["voxel"], k)
        img = g.gather_image()
        rad, frags = g.ctx.readback(A.SLOT_RADIANCE).copy(), g.ctx.counter(A.COUNTER_FRAGMENTS)
        # static / dynamic split across the ranks: the first two thirds of the scene captured once, the rest voxelized per frame
        s_ = (2 * sc.n_tris) // 3
        g.ctx.set_triangle_range(0, s_)
        g.capture_static(cams["voxel"])
        g.ctx.set_triangle_range(s_, sc.n_tris - s_)
        g.frame(cams["voxel"], k)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), img=img, rad=rad, chunks=g.chunks, rows=g.own_rows_mask(), frags=frags,
                 img_split=g.gather_image(), rad_split=g.ctx.readback(A.SLOT_RADIANCE), frags_split=g.ctx.counter(A.COUNTER_FRAGMENTS))
    finally:
        dist.destroy_process_group()


def test_sharded_frame_world2_gloo_equals_single_process(tmp_path, oracle_lib, proc_scene, cams):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(world)]
    # single-process reference run of the same frame
    o = A.VoxelGI(N, W, H, A.MODE_NORTHSTAR, shadow_res=SH, lib=oracle_lib)
    o.upload_scene(proc_scene)
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], W, H, SH, 0, cache=False)
    for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material"), (A.SLOT_SHADOW, "shadow")):
        o.upload(slot, fi[key])
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    o.voxelize(cams["voxel"]); o.inject(k); o.build_mips(); o.trace_indirect(k)
    want = o.readback(A.SLOT_INDIRECT_OUT)
    # the chunk lists partition the scene, the row bands partition the screen
    assert sorted(np.concatenate([r[0]["chunks"], r[1]["chunks"]]).tolist()) == list(range((proc_scene.n_tris + D.CHUNK - 1) // D.CHUNK))
    assert min(int(r[0]["frags"]), int(r[1]["frags"])) > 0
    assert (r[0]["rows"] ^ r[1]["rows"]).all()                 # interleaved 8-row tiles: every row traced by exactly one rank
    assert int(r[0]["frags"]) + int(r[1]["frags"]) == o.counter(A.COUNTER_FRAGMENTS)
    for i in range(world):
        assert np.array_equal(r[i]["rad"], o.readback(A.SLOT_RADIANCE)), f"rank {i}: summed partial volumes differ from the whole"
        assert np.array_equal(r[i]["img"].view(np.uint16), want.view(np.uint16)), f"rank {i}: assembled image differs"
        assert np.array_equal(r[i]["rad_split"], r[i]["rad"]) and np.array_equal(r[i]["img_split"].view(np.uint16), want.view(np.uint16)), f"rank {i}: static/dynamic split"
    assert int(r[0]["frags_split"]) + int(r[1]["frags_split"]) == o.counter(A.COUNTER_FRAGMENTS)


def test_row_selective_transfers(oracle_lib):
    """f184_upload_image_rows / f184_readback_async_rows move only the rows a rank traces (interleaved 8-row tiles, ragged last
    tile, a row window); everything else stays as it was.  Non-image slots travel whole."""
    w, h = 24, 44                                  # 5 full tiles + one of 4 rows
    c = A.VoxelGI(32, w, h, A.MODE_NORTHSTAR, shadow_res=16, lib=oracle_lib)
    rng = np.random.default_rng(0)
    old, new = (rng.random((h, w), dtype=np.float32) for _ in range(2))
    for first, stride, y0, y1 in ((1, 4, 0, 0xffffffff), (0, 2, 0, 0xffffffff), (1, 2, 8, 40), (0, 1, 0, 0xffffffff)):
        c.set_trace_tiles(first, stride); c.set_trace_rows(y0, y1)
        mine = ((np.arange(h) // 8) % stride == first) & (np.arange(h) >= y0) & (np.arange(h) < min(y1, h))
        c.upload(A.SLOT_DEPTH, old)
        c.upload_ptr(A.SLOT_DEPTH, new.ctypes.data, new.nbytes, rows=True)
        got = c.readback(A.SLOT_DEPTH)
        assert np.array_equal(got[mine], new[mine]) and np.array_equal(got[~mine], old[~mine]), (first, stride)
        host = np.full((h, w), -1.0, np.float32)
        c.readback_async_ptr(A.SLOT_DEPTH, host.ctypes.data, host.nbytes, rows=True)
        c.sync()
        assert np.array_equal(host[mine], got[mine]) and (host[~mine] == -1.0).all()
    sh = rng.random((16, 16), dtype=np.float32)
    c.set_trace_tiles(1, 4)
    c.upload_ptr(A.SLOT_SHADOW, sh.ctypes.data, sh.nbytes, rows=True)
    assert np.array_equal(c.readback(A.SLOT_SHADOW), sh)
    c.close()
