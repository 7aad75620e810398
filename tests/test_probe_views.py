"""BASELINE configs[4] — a batch of probe views traced against one volume, whole views per rank (f184_trace_views,
final184_b200.dist.ProbeBatch) — on the CPU: the stacked batch must equal the same views traced one by one as ordinary
images, and at world_size 2 (gloo) each rank's share must equal the corresponding views of the single-process batch."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from final184_b200 import api as A
from final184_b200 import dist as D
from final184_b200 import scene as S
from final184_b200.fixture import Fixture

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(REPO, "oracle", "_build", "libf184_oracle.so")
N, VS, SH, NV = 32, 40, 128, 5          # 5 views over 2 ranks: ragged split (2 + 3)


def batch_inputs(sc, n_views=NV, size=VS, shadow_res=SH, stride=13):
    cams = {n: S.fixture_constants(n) for n in ("shadow", "voxel")}
    names = [f"probe{(i * stride) % 64:02d}" for i in range(n_views)]
    views = [S.fixture_constants(n) for n in names]
    fx = Fixture(sc)
    per_view = [fx.gbuffer(v, size, size, 0) for v in views]
    shadow = fx.shadow(cams["shadow"], shadow_res)
    ks = [A.trace_constants_c(v, cams["shadow"], cams["voxel"], size, size, 0, True) for v in views]
    return cams, views, per_view, shadow, ks


def test_stacked_views_equal_views_traced_one_by_one(oracle_lib, proc_scene):
    cams, views, per_view, shadow, ks = batch_inputs(proc_scene)
    b = D.ProbeBatch(N, VS, NV, shadow_res=SH, scene=proc_scene, voxel_cam=cams["voxel"], lib=oracle_lib)
    b.upload_views(per_view, shadow)
    b.frame(cams["voxel"], ks)
    got = dict(b.own_views())
    total = b.ctx.counter(A.COUNTER_MARCH_STEPS)
    assert sorted(got) == list(range(NV))
    one = A.VoxelGI(N, VS, VS, A.MODE_NORTHSTAR, shadow_res=SH, lib=oracle_lib)
    one.upload_scene(proc_scene)
    one.upload(A.SLOT_SHADOW, shadow)
    one.voxelize(cams["voxel"]); one.inject(ks[0]); one.build_mips()
    samples = 0
    for v in range(NV):
        for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material")):
            one.upload(slot, per_view[v][key])
        one.trace_indirect(ks[v])
        want = one.readback(A.SLOT_INDIRECT_OUT)
        samples += one.counter(A.COUNTER_MARCH_STEPS)
        assert np.array_equal(got[v].view(np.uint16), want.view(np.uint16)), f"view {v}"
        assert np.isfinite(want.astype(np.float32)).all()
    assert total == samples > 0
    assert len({got[v].tobytes() for v in got}) == NV            # the views really differ
    # second frame with history: each view reprojects into ITS OWN rows of the history image
    b.ctx.copy_indirect_to_history()
    ks2 = [A.trace_constants_c(v, cams["shadow"], cams["voxel"], VS, VS, 1, False) for v in views]
    b.ctx.trace_views(ks2, VS)
    got2 = dict(b.own_views())
    for v in (0, NV - 1):
        for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material")):
            one.upload(slot, per_view[v][key])
        one.upload(A.SLOT_INDIRECT_HISTORY, got[v].view(np.uint16))
        one.trace_indirect(ks2[v])
        assert np.array_equal(got2[v].view(np.uint16), one.readback(A.SLOT_INDIRECT_OUT).view(np.uint16)), f"history, view {v}"
    # a sub-range of views leaves the other rows alone
    before = b.ctx.readback(A.SLOT_INDIRECT_OUT).copy()
    b.ctx.trace_views(ks, VS, first=1, count=2)
    after = b.ctx.readback(A.SLOT_INDIRECT_OUT)
    assert np.array_equal(after[:VS], before[:VS]) and np.array_equal(after[3 * VS:], before[3 * VS:])
    assert not np.array_equal(after[VS:3 * VS], before[VS:3 * VS])
    b.close(); one.close()


def test_trace_views_rejects_bad_windows(oracle_lib, proc_scene):
    c = A.VoxelGI(N, VS, VS * 2, A.MODE_NORTHSTAR, shadow_res=SH, lib=oracle_lib)
    k = A.trace_constants_c(*(S.fixture_constants(n) for n in ("main", "shadow", "voxel")), VS, VS, 0, True)
    for vh, first, count in ((0, 0, 1), (VS - 1, 0, 1), (VS, 0, 3), (VS, 2, 1)):
        try:
            c.trace_views([k, k, k], vh, first, count)
        except A.F184Error:
            continue
        raise AssertionError((vh, first, count))
    r = A.VoxelGI(N, VS, VS, A.MODE_REFERENCE, shadow_res=SH, lib=oracle_lib)
    try:
        r.trace_views([k], VS)
        raise AssertionError("reference mode accepted trace_views")
    except A.F184Error:
        pass


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        olib = A.Library(ORACLE, "f184o_", product=False)
        sc = S.procedural_scene(seed=1)
        cams, views, per_view, shadow, ks = batch_inputs(sc)
        b = D.ProbeBatch(N, VS, NV, shadow_res=SH, rank=rank, nranks=world, scene=sc, voxel_cam=cams["voxel"], volume_mode="host", lib=olib)
        b.upload_views(per_view, shadow)
        b.frame(cams["voxel"], ks)
        own = b.own_views()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=np.array([v for v, _ in own]), img=np.stack([i for _, i in own]))
    finally:
        dist.destroy_process_group()


def test_views_partitioned_over_two_ranks_gloo(tmp_path, oracle_lib, proc_scene):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    cams, views, per_view, shadow, ks = batch_inputs(proc_scene)
    b = D.ProbeBatch(N, VS, NV, shadow_res=SH, scene=proc_scene, voxel_cam=cams["voxel"], lib=oracle_lib)
    b.upload_views(per_view, shadow)
    b.frame(cams["voxel"], ks)
    want = dict(b.own_views())
    seen = []
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        for v, img in zip(z["idx"].tolist(), z["img"]):
            assert np.array_equal(img.view(np.uint16), want[v].view(np.uint16)), f"rank {r}, view {v}"
            seen.append(v)
    assert sorted(seen) == list(range(NV)) and D.view_ranges(NV, world) == [(0, 2), (2, 5)]
