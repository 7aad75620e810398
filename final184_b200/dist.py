"""One-process-per-GPU sharding of the voxel-GI frame (DESIGN.md "Multi-GPU").

The reference is single-GPU, single-queue (RHI/Private/Vulkan/DeviceVk.cpp:301-304); everything here is new.
The host-side partitioning logic is pure Python (tested with world_size-2 gloo on CPU); the data path is
libf184 on each rank's GPU, and `torch.distributed` is only the rendezvous / collective plumbing.

Partitioning helpers
  triangle_chunks(weights, n)   the scene dealt over the ranks in 128-triangle chunks, largest first, each to the least loaded rank
                                (Sponza's primitives range from 5 to 27,796 triangles; a contiguous cut by the same cost model left
                                one of two ranks with 0.50 ms of voxelization and the other with 0.32)
  triangle_ranges(weights, n)   contiguous triangle ranges balanced by the cost model (kept for callers that want one range)
  slab_ranges(N, n)             contiguous Z ranges of a level (host-mode tests; the device schedule owns bricks by diagonals: brick_owner)
  row_ranges(H, n, tile)        screen bands, multiples of the tracer's 8-row tile
  view_ranges(n_views, n)       whole views per rank (probe batches)

  brick_owner(x, y, z, n)       owner of the 8^3 brick holding voxel (x, y, z): bricks are dealt over the ranks along diagonals

Frame schedules: see ShardedVoxelGI.
"""
from __future__ import annotations

import numpy as np

from . import api as A
from . import scene as S


# ---- partitioning (pure host logic) ----------------------------------------------------------------
def triangle_ranges(weights: np.ndarray, nranks: int):
    """Contiguous [first, first+count) per rank with (almost) equal total weight.  weights: per-triangle cost
    estimate (projected area in voxels + a constant set-up cost)."""
    w = np.asarray(weights, np.float64)
    n = len(w)
    if nranks <= 1 or n == 0:
        return [(0, n)] + [(n, 0)] * (max(nranks, 1) - 1)
    c = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, nranks):
        cuts.append(int(np.searchsorted(c, c[-1] * r / nranks, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n))
    return [(int(cuts[r]), int(cuts[r + 1] - cuts[r])) for r in range(nranks)]


CHUNK = 128       # F184_TRIANGLE_CHUNK


def triangle_chunks(weights: np.ndarray, nranks: int, chunk: int = CHUNK):
    """Chunk c = triangles [chunk * c, chunk * (c + 1)).  Longest-processing-time-first: chunks sorted by cost, each given to the rank
    with the least cost so far.  Because the big chunks are dealt round the ranks first, every rank ends up with the same MIX of
    expensive and cheap triangles — so an error of the cost model that depends on triangle size cancels between the ranks instead of
    loading one of them.  Returns one sorted uint32 array of chunk ids per rank (sorted: neighbouring chunks share vertices)."""
    import heapq
    w = np.asarray(weights, np.float64)
    n_chunks = (len(w) + chunk - 1) // chunk
    if nranks <= 1:
        return [np.arange(n_chunks, dtype=np.uint32)]
    cost = np.add.reduceat(w, np.arange(0, len(w), chunk)) if len(w) else np.zeros(0)
    order = np.argsort(-cost, kind="stable")
    heap = [(0.0, r) for r in range(nranks)]
    mine = [[] for _ in range(nranks)]
    for c in order:
        load, r = heapq.heappop(heap)
        mine[r].append(int(c))
        heapq.heappush(heap, (load + float(cost[c]), r))
    return [np.array(sorted(m), dtype=np.uint32) for m in mine]


def triangle_weights(sc: S.Scene, voxel_cam: S.ViewConstants, grid_n: int, setup_cost: float = 4.0):
    """Cost model of the voxelizer: columns of the dominant-axis projection (+ set-up)."""
    mm = np.stack([voxel_cam.proj @ voxel_cam.view @ m for m in sc.model_mats]).astype(np.float64)
    p = np.concatenate([sc.pos.astype(np.float64), np.ones((len(sc.pos), 1))], 1)
    out = np.empty(sc.n_tris)
    for m in range(len(mm)):
        sel = np.nonzero(sc.tri_model == m)[0]
        if not len(sel):
            continue
        q = p @ mm[m].T
        q = q[:, :3] / q[:, 3:4]
        v = np.stack([(q[:, 0] * 0.5 + 0.5), (q[:, 1] * 0.5 + 0.5), q[:, 2]], 1) * grid_n
        t = v[sc.idx[sel]]
        n = np.abs(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]))
        ext = t.max(1) - t.min(1) + 1.0
        d = n.argmax(1)
        ua, va = (d + 1) % 3, (d + 2) % 3
        r = np.arange(len(sel))
        out[sel] = ext[r, ua] * ext[r, va] + setup_cost
    return out


def slab_ranges(n: int, nranks: int):
    """Contiguous Z-slab of each rank at a level of edge n (empty slabs once n < nranks)."""
    return [(r * n // nranks, (r + 1) * n // nranks) for r in range(nranks)]


def brick_owner(x: int, y: int, z: int, nranks: int) -> int:
    """Rank that owns the 8^3 brick of voxel (x, y, z): owner = (x//8 + y//8 + z//8) % nranks, a diagonal interleave.
    (Contiguous Z slabs left the ranks holding Sponza's floor and arcades with 40 % of the bricks and the top slab with none;
    whole brick LAYERS dealt round-robin still gave the rank holding the floor layer 3x the bricks of the lightest rank at
    8 GPUs — every axis-aligned sheet of bricks now lands evenly on all ranks.)"""
    return (x // 8 + y // 8 + z // 8) % nranks


def row_ranges(height: int, nranks: int, tile: int = 8):
    tiles = (height + tile - 1) // tile
    cuts = [min(height, (r * tiles // nranks) * tile) for r in range(nranks)] + [height]
    return [(cuts[r], cuts[r + 1]) for r in range(nranks)]


def view_ranges(n_views: int, nranks: int):
    return [(r * n_views // nranks, (r + 1) * n_views // nranks) for r in range(nranks)]


# ---- the sharded frame ------------------------------------------------------------------------------
class ShardedVoxelGI:
    """The voxel/indirect section of one frame on rank `rank` of `nranks` GPUs of one NVLink box.

    mode "single"     one GPU.
         "slab"       the north-star schedule (include/f184.h "one NVLink box"): every rank voxelizes its triangle chunks; fragments
                      of bricks another rank owns (diagonal ownership) go into local queues the owner pulls over NVLink
                      (an all-to-all of fragments = the reduce-scatter); the owners normalise / inject / build levels 1-3 of
                      their bricks; a peer gather of the finished bricks (bulk copies, level 1 only where this rank's cones
                      need it); tracing of interleaved 8-row screen tiles.  Needs connect().
         "replicate"  no exchange at all: every rank builds the whole volume, only the trace is split.
         "host"       the slab schedule's HOST logic with the exchange done by torch.distributed on host arrays
                      (all_reduce of the partial accumulators): what the CPU tests drive with the gloo backend and the
                      oracle library; never used on the product path.
    """

    def __init__(self, grid_n, width, height, shadow_res=2048, device=0, rank=0, nranks=1, scene: S.Scene | None = None,
                 mode=None, lib=None, voxel_cam: S.ViewConstants | None = None, flags=0):
        self.rank, self.nranks = rank, nranks
        self.mode = mode or ("slab" if nranks > 1 else "single")
        if nranks == 1:
            self.mode = "single"
        self.grid_n, self.width, self.height = grid_n, width, height
        ctx_ranks = (rank, nranks) if self.mode == "slab" else (0, 1)       # only the slab schedule shards the volume
        self.ctx = A.VoxelGI(grid_n, width, height, A.MODE_NORTHSTAR, shadow_res=shadow_res, device=device, rank=ctx_ranks[0], nranks=ctx_ranks[1],
                             lib=lib, flags=flags)
        self.rows = (0, height)
        self.tiles = (rank, nranks)              # this rank traces the 8-row tile rows t with t % nranks == rank
        self.ctx.set_trace_tiles(*self.tiles)
        self.tri_range = None
        self.connected = False
        if scene is not None:
            self.upload_scene(scene, voxel_cam)

    def upload_scene(self, scene: S.Scene, voxel_cam: S.ViewConstants | None = None):
        self.ctx.upload_scene(scene)
        if self.mode in ("slab", "host"):
            w = triangle_weights(scene, voxel_cam, self.grid_n) if voxel_cam is not None else np.ones(scene.n_tris)
            self.chunks = triangle_chunks(w, self.nranks)[self.rank]
            self.ctx.set_triangle_chunks(self.chunks)
            self.tri_range = (0, scene.n_tris)        # a range still filters on top of the chunk list (f184_set_triangle_range)
            self.ctx.set_triangle_range(*self.tri_range)

    def connect(self):
        """Exchange CUDA IPC handles of the shared buffers with the other ranks (torch.distributed is the rendezvous)."""
        if self.mode != "slab" or self.connected:
            return
        import torch.distributed as dist
        mine = {b: self.ctx.ipc_export(b) for b in range(A.IPC_COUNT)}
        everyone = [None] * self.nranks
        dist.all_gather_object(everyone, mine)
        for p, handles in enumerate(everyone):
            if p == self.rank:
                continue
            for b, h in handles.items():
                self.ctx.ipc_import(p, b, h)
        dist.barrier()
        self.connected = True

    @staticmethod
    def connect_loopback(shards):
        """Several ranks inside ONE process (tests on a one-GPU box): cudaIpcOpenMemHandle refuses handles of the own process, so
        the peers' buffers are installed as plain device pointers (f184_debug_get_ipc_ptr / f184_debug_set_peer).  The contexts
        may share a device; their barrier kernels then wait for each other on the same GPU, so the caller must enqueue every
        rank's frame before synchronising any of them."""
        ptrs = [{b: s.ctx.ipc_ptr(b) for b in range(A.IPC_COUNT)} for s in shards]
        for s in shards:
            for p, other in enumerate(shards):
                if other is s:
                    continue
                for b, ptr in ptrs[p].items():
                    s.ctx.set_peer(p, b, ptr)
            s.connected = True

    def describe(self):
        if self.mode == "single":
            return "1 GPU"
        if self.mode == "replicate":
            return f"{self.nranks} GPUs: volume replicated per rank (no exchange), trace split by interleaved 8-row tiles"
        return (f"{self.nranks} GPUs: 128-triangle chunks dealt over the ranks by projected area (largest first); 8^3 bricks owned along diagonals; fragments of foreign bricks queued "
                f"locally and pulled by the owner over NVLink; owner-local normalise/inject/mips; view-driven peer gather of the listed bricks (TMA bulk copies); "
                f"trace split by interleaved 8-row tiles")

    def frame(self, voxel_cam, k, trace=True):
        c = self.ctx
        if self.mode in ("single", "replicate"):
            c.voxelize(voxel_cam)
            c.inject(k)
            c.build_mips()
        elif self.mode == "slab":
            if not self.connected:
                raise A.F184Error("ShardedVoxelGI.frame: call connect() first (slab mode shares buffers between ranks)")
            c.voxelize_accumulate(voxel_cam)     # own bricks: local reductions; foreign bricks: records in local queues
            c.peer_barrier()                     # every rank's queues are complete; normalise starts by pulling and applying them
            c.normalise()
            c.inject(k)
            c.build_mips()                       # levels 1-3 of the own bricks + the export arrays
            c.peer_barrier()
            c.gather_volume(k if trace else None)    # fetch the other ranks' bricks (level 1 only where this rank's cones sample it), finish the small levels
        else:                                    # "host": same order, exchange through torch.distributed on host arrays
            import torch
            import torch.distributed as dist
            c.voxelize_accumulate(voxel_cam)
            for slot in (A.SLOT_ACCUM_COLOR, A.SLOT_ACCUM_NORMAL):
                t = torch.from_numpy(c.readback(slot))
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                c.upload(slot, t.numpy())
            c.normalise()
            c.inject(k)
            c.build_mips()
        if trace:
            c.trace_indirect(k)

    def capture_static(self, voxel_cam):
        """Static / dynamic split (include/f184.h): the CURRENT triangle selection (set_triangle_range on top of this rank's chunks) is
        the geometry that never moves — accumulate it once, complete it across the ranks, and keep every rank's own bricks of it
        (f184_static_cache_capture).  Afterwards select the dynamic triangles and call frame() as usual."""
        c = self.ctx
        c.voxelize_accumulate(voxel_cam)
        if self.mode == "slab":
            if not self.connected:
                raise A.F184Error("ShardedVoxelGI.capture_static: call connect() first")
            c.peer_barrier()
        elif self.mode == "host":
            import torch
            import torch.distributed as dist
            for slot in (A.SLOT_ACCUM_COLOR, A.SLOT_ACCUM_NORMAL):
                t = torch.from_numpy(c.readback(slot))
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                c.upload(slot, t.numpy())
        c.static_cache_capture()
        if self.mode == "slab":
            c.peer_barrier()              # every rank has pulled its fragments out of the others' queues: they may be started over

    def gather_image(self):
        """The full traced image on every rank (a consumer that wants one image; not part of the frame)."""
        img = self.ctx.readback(A.SLOT_INDIRECT_OUT)
        if self.nranks == 1:
            return img
        import torch
        import torch.distributed as dist
        parts = [None] * self.nranks
        dist.all_gather_object(parts, (self.tiles, img[self.own_rows_mask()].copy()))
        out = np.zeros_like(img)
        for (first, stride), rows in parts:
            out[(np.arange(self.height) // 8) % stride == first] = rows
        return out

    def own_rows_mask(self):
        first, stride = self.tiles
        return (np.arange(self.height) // 8) % stride == first

    def comm_ms_per_frame(self):
        return 0.0

    def close(self):
        self.ctx.close()


class ProbeBatch:
    """BASELINE configs[4]: `n_views` square probe views cone-traced against one volume, whole views per rank
    (view_ranges) — no exchange for the views.  Each rank's context holds ITS views stacked top to bottom
    (f184_trace_views); the volume is built by any ShardedVoxelGI schedule ("replicate": every rank builds it;
    "slab": built once across the box and gathered; "host": the CPU-test stand-in)."""

    def __init__(self, grid_n, view_size, n_views, shadow_res=2048, device=0, rank=0, nranks=1, scene=None, voxel_cam=None,
                 volume_mode=None, lib=None, flags=0):
        self.rank, self.nranks, self.n_views, self.view_size = rank, nranks, n_views, view_size
        self.v0, self.v1 = view_ranges(n_views, nranks)[rank]
        self.local = self.v1 - self.v0
        mode = volume_mode or ("replicate" if nranks > 1 else None)
        self.gi = ShardedVoxelGI(grid_n, view_size, view_size * max(self.local, 1), shadow_res=shadow_res, device=device, rank=rank,
                                 nranks=nranks, scene=scene, mode=mode, lib=lib, voxel_cam=voxel_cam, flags=flags)
        self.ctx = self.gi.ctx
        # a rank traces ALL rows of its views (whole views per rank), not interleaved tile rows: the gather's "what do my cones sample"
        # pass (level 0 for glossy pixels) must look at every row of this context's image
        self.gi.tiles = (0, 1)
        self.ctx.set_trace_tiles(0, 1)

    def upload_views(self, per_view_inputs, shadow):
        """per_view_inputs: one dict(depth, normals, material) per view of the WHOLE batch (indexed by global view);
        only this rank's views are stacked and uploaded."""
        mine = per_view_inputs[self.v0:self.v1]
        if mine:
            for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material")):
                self.ctx.upload(slot, np.ascontiguousarray(np.concatenate([fi[key] for fi in mine], axis=0)))
        self.ctx.upload(A.SLOT_SHADOW, shadow)

    def frame(self, voxel_cam, ks):
        """ks: trace constants of every view of the batch (global index).  Volume once, then this rank's views."""
        self.gi.frame(voxel_cam, ks[0], trace=False)
        if self.local:
            self.ctx.trace_views(list(ks[self.v0:self.v1]), self.view_size)

    def own_views(self):
        """(global view index, image) for each of this rank's views"""
        if not self.local:
            return []
        img = self.ctx.readback(A.SLOT_INDIRECT_OUT)
        s = self.view_size
        return [(self.v0 + i, img[i * s:(i + 1) * s].copy()) for i in range(self.local)]

    def close(self):
        self.gi.close()
