"""Worker for tests/test_gpu_loopback.py: the multi-rank slab schedule (include/f184.h "one NVLink box") driven on ONE GPU.

G contexts of this process act as the ranks of a box ("loopback ranks": the peers' buffers are installed as plain device
pointers, final184_b200.dist.ShardedVoxelGI.connect_loopback).  Everything the real schedule does happens — triangle-range
voxelization with the reduction into the owner's accumulators (peer reductions at system scope, here on the same device), the
device-side flag barriers (the ranks' barrier kernels wait for each other while they share the GPU), owner-side normalise /
inject / mips with the packed export records, the gather, the level-0 skip, the three-stream frame pipeline — and every rank's
volume, texture storage and image rows must equal the frame one context computes alone, bit for bit.

Run with CUDA_DEVICE_MAX_CONNECTIONS=32 so that every stream has its own hardware queue (a kernel that waits in a barrier must
never sit in front of another rank's work) and CUDA_MODULE_LOADING=EAGER (the first launch of a lazily loaded kernel waits for the
device to drain, which a barrier waiting for a rank not yet enqueued never lets happen).  Every rank's frame is enqueued before
any rank is synchronised.  None of this concerns the real schedule: one process per GPU, every rank with its own host thread.
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from final184_b200 import api as A, dist as D, scene as S   # noqa: E402
from final184_b200.fixture import frame_inputs               # noqa: E402

SLOTS = ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material"), (A.SLOT_SHADOW, "shadow"))


def make_box(G, N, W, H, SH, sc, cams, fi, flags):
    ranks = [D.ShardedVoxelGI(N, W, H, shadow_res=SH, device=0, rank=r, nranks=G, scene=sc, voxel_cam=cams["voxel"], mode="slab", flags=flags)
             for r in range(G)]
    for g in ranks:
        for slot, key in SLOTS:
            g.ctx.upload(slot, fi[key])
    D.ShardedVoxelGI.connect_loopback(ranks)
    return ranks


def box_frame(ranks, cam, k):
    if os.environ.get("LOOPBACK_PHASED"):
        # diagnostics: the same frame phase by phase, every rank synchronised after each phase — a failing kernel names its phase
        phases = [("accumulate", lambda c: c.voxelize_accumulate(cam)), ("barrier 1", lambda c: c.peer_barrier()), ("normalise", lambda c: c.normalise()),
                  ("inject", lambda c: c.inject(k)), ("mips", lambda c: c.build_mips()), ("barrier 2", lambda c: c.peer_barrier()),
                  ("gather", lambda c: c.gather_volume(k)), ("trace", lambda c: c.trace_indirect(k))]
        for name, fn in phases:
            if name.startswith("barrier") and os.environ.get("LOOPBACK_PHASED") == "2":
                continue                # (under compute-sanitizer kernels run one at a time: the host synchronisation between the phases stands in)
            for g in ranks:
                fn(g.ctx)
            for r, g in enumerate(ranks):
                try:
                    g.ctx.sync()
                except A.F184Error as e:
                    raise SystemExit(f"phase '{name}' failed on rank {r}: {e}")
        return
    for g in ranks:                 # enqueue every rank's frame before synchronising any of them
        g.frame(cam, k)


def box_sync(ranks):
    for g in ranks:
        g.ctx.sync()


def single(N, W, H, SH, sc, fi, flags=0):
    one = A.VoxelGI(N, W, H, A.MODE_NORTHSTAR, shadow_res=SH, device=0, flags=flags)
    one.upload_scene(sc)
    for slot, key in SLOTS:
        one.upload(slot, fi[key])
    return one


def one_frame(one, cam, k):
    one.voxelize(cam); one.inject(k); one.build_mips(); one.trace_indirect(k)


def compare_arrays(bad, tag, g, one, N, level0=True, level1=True):
    if level0 and not np.array_equal(g.ctx.read_array(-1, 0, N), one.read_array(-1, 0, N)):
        bad.append(f"{tag}: level-0 array")
    m, lvl = N // 2, 0
    while m >= 1:
        for d in range(6):
            if (lvl or level1) and not np.array_equal(g.ctx.read_array(d, lvl, m), one.read_array(d, lvl, m)):
                bad.append(f"{tag}: array dir {d} level {lvl + 1}")
        m //= 2; lvl += 1


def compare_rows(bad, tag, g, one):
    m = g.own_rows_mask()
    a, b = g.ctx.readback(A.SLOT_INDIRECT_OUT)[m], one.readback(A.SLOT_INDIRECT_OUT)[m]
    if not np.array_equal(a.view(np.uint16), b.view(np.uint16)):
        bad.append(f"{tag}: own image tile rows")


def set_ranges(ranks, one, lo, hi):
    """triangles [lo, hi) of the scene: each rank takes its share, the single context all of them"""
    for g in ranks:
        f0, cnt = g.tri_range
        a, b = max(f0, lo), min(f0 + cnt, hi)
        g.ctx.set_triangle_range(a, max(0, b - a))
    one.set_triangle_range(lo, hi - lo)


def run_box(G, sc, cams, N, W, H, SH, bad, sponza=False):
    tag = f"G={G} {'sponza' if sponza else 'procedural'} {N}^3"
    fi = frame_inputs(sc, cams["main"], cams["shadow"], W, H, SH, 0, cache=False)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    T = sc.n_tris
    # ---- (1) full comparison: F184_FLAG_GATHER_LINEAR also fills the linear slots and always moves level 0
    ranks = make_box(G, N, W, H, SH, sc, cams, fi, A.FLAG_GATHER_LINEAR)
    one = single(N, W, H, SH, sc, fi)
    for frame, (lo, hi) in enumerate(((0, T), (0, T), (0, T // 2))):      # the last frame empties bricks: they must be cleared on every rank
        set_ranges(ranks, one, lo, hi)
        box_frame(ranks, cams["voxel"], k)
        one_frame(one, cams["voxel"], k)
        box_sync(ranks); one.sync()
        for r, g in enumerate(ranks):
            t = f"{tag} frame {frame} rank {r}"
            if not np.array_equal(g.ctx.readback(A.SLOT_RADIANCE), one.readback(A.SLOT_RADIANCE)): bad.append(t + ": radiance")
            if not np.array_equal(g.ctx.readback(A.SLOT_MIPS), one.readback(A.SLOT_MIPS)): bad.append(t + ": mips")
            compare_arrays(bad, t, g, one, N)
            compare_rows(bad, t, g, one)
    frags = sum(g.ctx.counter(A.COUNTER_FRAGMENTS) for g in ranks)
    if frags != one.counter(A.COUNTER_FRAGMENTS): bad.append(f"{tag}: fragments {frags} vs {one.counter(A.COUNTER_FRAGMENTS)}")
    for g in ranks: g.close()
    if sponza:
        # the product path at a real size: view-driven gather (records of three sizes in one ring), two frames so both texture sets are used
        ranks = make_box(G, N, W, H, SH, sc, cams, fi, 0)
        set_ranges(ranks, one, 0, T)
        for frame in range(2):
            box_frame(ranks, cams["voxel"], k)
            one_frame(one, cams["voxel"], k)
            box_sync(ranks); one.sync()
            for r, g in enumerate(ranks):
                t = f"{tag} view-driven frame {frame} rank {r}"
                compare_arrays(bad, t, g, one, N, level0=False, level1=False)
                compare_rows(bad, t, g, one)
        for g in ranks: g.close()
        one.close()
        return
    # ---- (2) the product path: only what this rank's cones sample travels.  Level 0: the fixture's materials are rough (roughness 1:
    # no cone reads level 0); a glossy G-buffer (roughness 0.2) makes the specular cones read it, and then every listed brick travels
    # whole.  Level 1: only the bricks around the surfaces the rank's pixels show — so the camera moves between frames (two fixture
    # cameras, each with its own G-buffer): bricks needed for one view and not for the next must read zero, not the stale copy, when a
    # later view needs them again after they have changed.  The image rows are the check (bit for bit against one context).
    ranks = make_box(G, N, W, H, SH, sc, cams, fi, 0)
    glossy = fi["material"].copy(); glossy[..., 1] = 51
    cam_b = S.fixture_constants("probe05")
    views = {"a": (fi, k), "b": (frame_inputs(sc, cam_b, cams["shadow"], W, H, SH, 0, cache=False), A.trace_constants_c(cam_b, cams["shadow"], cams["voxel"], W, H, 0, True))}
    schedule = [("rough", "a", 0, T), ("rough", "a", 0, T), ("glossy", "a", 0, T), ("rough", "b", T // 3, T), ("rough", "a", 0, T // 2), ("rough", "b", 0, T),
                ("glossy", "b", T // 4, T), ("glossy", "a", 0, T), ("rough", "b", 0, T // 2), ("rough", "a", 0, T)]
    for frame, (mat, view, lo, hi) in enumerate(schedule):
        set_ranges(ranks, one, lo, hi)
        vfi, vk = views[view]
        m_ = vfi["material"].copy()
        if mat == "glossy":
            m_[..., 1] = 51
        for c in [g.ctx for g in ranks] + [one]:
            for slot, key in SLOTS[:2]:
                c.upload(slot, vfi[key])
            c.upload(A.SLOT_MATERIAL, m_)
        box_frame(ranks, cams["voxel"], vk)
        one_frame(one, cams["voxel"], vk)
        box_sync(ranks); one.sync()
        for r, g in enumerate(ranks):
            t = f"{tag} skip-path frame {frame} ({mat}, view {view}) rank {r}"
            compare_arrays(bad, t, g, one, N, level0=(mat == "glossy"), level1=(mat == "glossy"))
            compare_rows(bad, t, g, one)
    for c in [g.ctx for g in ranks] + [one]:
        for slot, key in SLOTS[:2]:
            c.upload(slot, fi[key])
    # ---- (3) the frame pipeline: frames back to back, no host synchronisation in between, a different triangle range each
    # (every frame's volume differs, emptied bricks must be cleared in BOTH texture sets); read-backs staged asynchronously
    for c in [g.ctx for g in ranks] + [one]:
        c.upload(A.SLOT_MATERIAL, fi["material"])
    spans = [(0, T), (0, T // 2), (T // 3, T), (0, T), (T // 2, T), (0, T // 4), (0, T)]
    want = []
    for lo, hi in spans:
        one.set_triangle_range(lo, hi - lo)
        one_frame(one, cams["voxel"], k)
        want.append(one.readback(A.SLOT_INDIRECT_OUT).copy())
    if np.array_equal(want[0], want[1]): bad.append(f"{tag}: back-to-back frames do not differ")
    nbytes = ranks[0].ctx.image_info(A.SLOT_INDIRECT_OUT).size_bytes
    hosts = [[torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in spans] for _ in ranks]
    for i, (lo, hi) in enumerate(spans):
        set_ranges(ranks, one, lo, hi)
        box_frame(ranks, cams["voxel"], k)
        for r, g in enumerate(ranks):
            g.ctx.readback_async_ptr(A.SLOT_INDIRECT_OUT, hosts[r][i].data_ptr(), nbytes)
    box_sync(ranks)
    for r, g in enumerate(ranks):
        m = g.own_rows_mask()
        for i in range(len(spans)):
            got = hosts[r][i].numpy().view(np.uint16).reshape(H, W, 4)
            if not np.array_equal(got[m], want[i].view(np.uint16).reshape(H, W, 4)[m]): bad.append(f"{tag}: back-to-back frame {i} rank {r}: own image tile rows")
    for g in ranks: g.close()
    one.close()


def run_single_pipeline(sc, cams, bad):
    """one GPU: pipelined frames back to back == the same frames with every pass on one stream (F184_FLAG_NO_OVERLAP)"""
    N, W, H, SH = 64, 160, 96, 256
    fi = frame_inputs(sc, cams["main"], cams["shadow"], W, H, SH, 0, cache=False)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    T = sc.n_tris
    spans = [(0, T), (0, T // 2), (T // 3, T), (0, T), (T // 2, T), (0, T // 4), (0, T), (0, T)]
    outs = []
    for flags in (A.FLAG_NO_OVERLAP, 0):
        c = single(N, W, H, SH, sc, fi, flags)
        nbytes = c.image_info(A.SLOT_INDIRECT_OUT).size_bytes
        hosts = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in spans]
        for i, (lo, hi) in enumerate(spans):
            c.set_triangle_range(lo, hi - lo)
            one_frame(c, cams["voxel"], k)
            c.readback_async_ptr(A.SLOT_INDIRECT_OUT, hosts[i].data_ptr(), nbytes)
        c.sync()
        outs.append([h.numpy().copy() for h in hosts])
        # the volume of the last frame, both storages
        outs[-1].append(c.readback(A.SLOT_MIPS).copy().view(np.uint8).reshape(-1))
        outs[-1].append(c.read_array(0, 0, N // 2).copy().reshape(-1))
        c.close()
    for i, (a, b) in enumerate(zip(*outs)):
        if not np.array_equal(a, b): bad.append(f"single-GPU pipeline: output {i} differs from the one-stream frames")


def run_barrier_timeout(bad):
    """a barrier whose peer never arrives must surface as F184_ERR_PEER_TIMEOUT from the next synchronous call, not as a silently
    wrong volume (F184_BARRIER_TIMEOUT_MS shortens the 5 s default)"""
    N, W, H = 32, 32, 16
    sc = S.procedural_scene(seed=2)
    cam = S.fixture_constants("voxel")
    c = A.VoxelGI(N, W, H, A.MODE_NORTHSTAR, shadow_res=64, device=0, rank=0, nranks=2)
    c.upload_scene(sc)
    sizes = {A.IPC_ACCUM_COLOR: N ** 3 * 16, A.IPC_ACCUM_NORMAL: N ** 3 * 16, A.IPC_BRICK_FLAGS: (N // 8) ** 3 * 4, A.IPC_EXPORT: (N // 8) ** 3 * 4096,
             A.IPC_COUNTERS: 256, A.IPC_BRICK_LIST: (N // 8) ** 3 * 4, A.IPC_SYNC: 64, A.IPC_FRAG_QUEUE: 2 * (1 << 20) * 16, A.IPC_FRAG_COUNTS: 8 * 128 * 128}
    fake = {b: torch.zeros(n, dtype=torch.uint8, device="cuda:0") for b, n in sizes.items()}     # a "peer" nobody runs
    for b, t in fake.items():
        c.ipc_ptr(b)
        c.set_peer(1, b, t.data_ptr())
    c.voxelize_accumulate(cam)
    c.peer_barrier()
    try:
        c.sync()
        bad.append("barrier timeout: f184_sync returned OK although the peer never arrived")
    except A.F184Error as e:
        if "(-7)" not in str(e) or "peer" not in str(e): bad.append(f"barrier timeout: unexpected error {e}")
    c.sync()      # reported once: the context stays usable
    c.close()


def run_static_cache(sc, cams, bad):
    """Static / dynamic split (include/f184.h): triangles [0, S) are accumulated once and captured; every frame
    voxelizes only a dynamic range [S, end).  Volumes, texture sets, image and counters must equal a context that voxelizes [0, end)
    every frame — on one context (synchronised frames, then frames back to back through the pipeline) and on two loopback ranks."""
    N, W, H, SH = 64, 160, 96, 256
    fi = frame_inputs(sc, cams["main"], cams["shadow"], W, H, SH, 0, cache=False)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    cam = cams["voxel"]
    T = sc.n_tris
    S_ = (2 * T) // 3
    ends = [T, S_ + (T - S_) // 2, S_, T, S_ + (T - S_) // 4, S_, S_, T]          # S_: a frame without any dynamic triangle

    class Shim:                       # compare_arrays / compare_rows take a sharded rank
        def __init__(self, ctx): self.ctx = ctx
        def own_rows_mask(self): return np.ones(H, bool)

    def check(tag, c, ref, counters=True):
        for slot, name in ((A.SLOT_VOX_ALBEDO, "albedo"), (A.SLOT_VOX_NORMAL, "normal"), (A.SLOT_RADIANCE, "radiance"), (A.SLOT_MIPS, "mips")):
            if not np.array_equal(c.readback(slot), ref.readback(slot)): bad.append(f"{tag}: {name}")
        compare_arrays(bad, tag, Shim(c), ref, N)
        compare_rows(bad, tag, Shim(c), ref)
        if counters:
            for w, name in ((A.COUNTER_FRAGMENTS, "fragments"), (A.COUNTER_OCCUPIED, "occupied"), (A.COUNTER_BRICKS, "bricks")):
                if c.counter(w) != ref.counter(w): bad.append(f"{tag}: counter {name} {c.counter(w)} vs {ref.counter(w)}")

    # ---- (a) one context, a synchronised comparison after every frame
    ref, c = single(N, W, H, SH, sc, fi), single(N, W, H, SH, sc, fi)
    c.set_triangle_range(0, S_)
    c.voxelize_accumulate(cam)
    c.static_cache_capture()
    for i, end in enumerate(ends):
        ref.set_triangle_range(0, end); one_frame(ref, cam, k)
        c.set_triangle_range(S_, end - S_); one_frame(c, cam, k)
        ref.sync(); c.sync()
        check(f"static cache frame {i} (dynamic triangles {end - S_})", c, ref)
    # another voxel camera: refused, and the context stays usable
    try:
        c.voxelize_accumulate(cams["main"])
        bad.append("static cache: accumulation with another voxel camera was accepted")
    except A.F184Error as e:
        if "another voxel camera" not in str(e): bad.append(f"static cache: unexpected error {e}")
    # capture again with a different split, then drop the cache: both back to the reference
    S2 = T // 3
    c.set_triangle_range(0, S2); c.voxelize_accumulate(cam); c.static_cache_capture()
    for i, end in enumerate((T, S2, S2 + 5)):
        ref.set_triangle_range(0, end); one_frame(ref, cam, k)
        c.set_triangle_range(S2, end - S2); one_frame(c, cam, k)
        ref.sync(); c.sync()
        check(f"static cache, second capture, frame {i}", c, ref)
    c.static_cache_clear()
    for i, end in enumerate((T // 2, T)):
        ref.set_triangle_range(0, end); one_frame(ref, cam, k)
        c.set_triangle_range(0, end); one_frame(c, cam, k)
        ref.sync(); c.sync()
        check(f"static cache cleared, frame {i}", c, ref)
    # ---- (b) frames back to back (three-stream pipeline, both texture sets), images staged asynchronously
    c.set_triangle_range(0, S_); c.voxelize_accumulate(cam); c.static_cache_capture()
    nbytes = c.image_info(A.SLOT_INDIRECT_OUT).size_bytes
    want = []
    for end in ends:
        ref.set_triangle_range(0, end); one_frame(ref, cam, k)
        want.append(ref.readback(A.SLOT_INDIRECT_OUT).copy())
    hosts = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in ends]
    for i, end in enumerate(ends):
        c.set_triangle_range(S_, end - S_); one_frame(c, cam, k)
        c.readback_async_ptr(A.SLOT_INDIRECT_OUT, hosts[i].data_ptr(), nbytes)
    c.sync()
    for i in range(len(ends)):
        if not np.array_equal(hosts[i].numpy().view(np.uint16), want[i].view(np.uint16).reshape(-1)): bad.append(f"static cache: back-to-back frame {i}")
    c.close()
    # ---- (c) two loopback ranks: every rank captures its own bricks (after the barrier that completes the static accumulation)
    ranks = make_box(2, N, W, H, SH, sc, cams, fi, A.FLAG_GATHER_LINEAR)
    set_ranges(ranks, ref, 0, S_)
    for g in ranks: g.ctx.voxelize_accumulate(cam)
    for g in ranks: g.ctx.peer_barrier()
    box_sync(ranks)
    for g in ranks: g.ctx.static_cache_capture()
    for g in ranks: g.ctx.peer_barrier()
    box_sync(ranks)
    for i, end in enumerate(ends[:5]):
        set_ranges(ranks, ref, S_, end)
        ref.set_triangle_range(0, end)
        box_frame(ranks, cam, k); one_frame(ref, cam, k)
        box_sync(ranks); ref.sync()
        for r, g in enumerate(ranks):
            t = f"static cache, 2 ranks, frame {i} rank {r}"
            if not np.array_equal(g.ctx.readback(A.SLOT_RADIANCE), ref.readback(A.SLOT_RADIANCE)): bad.append(t + ": radiance")
            if not np.array_equal(g.ctx.readback(A.SLOT_MIPS), ref.readback(A.SLOT_MIPS)): bad.append(t + ": mips")
            compare_arrays(bad, t, g, ref, N)
            compare_rows(bad, t, g, ref)
        if sum(g.ctx.counter(A.COUNTER_FRAGMENTS) for g in ranks) != ref.counter(A.COUNTER_FRAGMENTS): bad.append(f"static cache, 2 ranks, frame {i}: fragments")
    for g in ranks: g.close()
    ref.close()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.cuda.set_device(0)
    if os.environ.get("CUDA_MODULE_LOADING") != "EAGER":
        print("warning: CUDA_MODULE_LOADING is not EAGER: loopback ranks may time out in their first barrier", file=sys.stderr)
    sc = S.procedural_scene(seed=1)
    cams = {n: S.fixture_constants(n) for n in ("main", "shadow", "voxel")}
    bad = []
    if which in ("all", "timeout"):
        run_barrier_timeout(bad)
    if which in ("all", "single"):
        run_single_pipeline(sc, cams, bad)
    if which in ("all", "static"):
        run_static_cache(sc, cams, bad)
    if which in ("all", "box2"):
        run_box(2, sc, cams, 64, 160, 96, 256, bad)
    if which in ("all", "box4"):
        run_box(4, sc, cams, 64, 160, 96, 256, bad)
    if which in ("all", "sponza") and S.sponza_available():
        run_box(4, S.load_sponza(), cams, 256, 320, 184, 1024, bad, sponza=True)
    print(("OK" if not bad else "MISMATCH " + "; ".join(bad[:8])) + f" [{which}]", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
