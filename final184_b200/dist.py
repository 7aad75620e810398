"""One-process-per-GPU sharding of the voxel-GI frame (DESIGN.md "Multi-GPU").

The reference is single-GPU, single-queue (RHI/Private/Vulkan/DeviceVk.cpp:301-304); everything here is new.
The host-side partitioning logic is pure Python (tested with world_size-2 gloo on CPU); the data path is
libf184 on each rank's GPU, and `torch.distributed` is only the rendezvous / collective plumbing.

Partitioning helpers
  triangle_ranges(weights, n)   contiguous triangle ranges balanced by projected area, not by count
                                (Sponza's primitives range from 5 to 27,796 triangles)
  slab_ranges(N, n)             Z-slab [z0, z1) of every volume level owned by each rank
  row_ranges(H, n, tile)        screen bands, multiples of the tracer's 8-row tile
  view_ranges(n_views, n)       whole views per rank (probe batches)

Frame schedules (ShardedVoxelGI.mode)
  "replicate"  every rank voxelizes / injects / builds the whole volume (no data-path collective at all) and
               traces its own band of rows; the image bands are disjoint, a consumer gathers them only if it wants
               one image on one rank.  This is the schedule that wins while voxelize+mips are a small part of the
               frame, because it has no exchange step to pay for.
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
    """Z-slab of each rank at a level of edge n (empty slabs once n < nranks)."""
    return [(r * n // nranks, (r + 1) * n // nranks) for r in range(nranks)]


def row_ranges(height: int, nranks: int, tile: int = 8):
    tiles = (height + tile - 1) // tile
    cuts = [min(height, (r * tiles // nranks) * tile) for r in range(nranks)] + [height]
    return [(cuts[r], cuts[r + 1]) for r in range(nranks)]


def view_ranges(n_views: int, nranks: int):
    return [(r * n_views // nranks, (r + 1) * n_views // nranks) for r in range(nranks)]


# ---- the sharded frame ------------------------------------------------------------------------------
class ShardedVoxelGI:
    """The voxel/indirect section of one frame on rank `rank` of `nranks` GPUs of one NVLink box."""

    def __init__(self, grid_n, width, height, shadow_res=2048, device=0, rank=0, nranks=1, scene: S.Scene | None = None,
                 mode="replicate", lib=None):
        self.rank, self.nranks, self.mode = rank, nranks, mode
        self.grid_n, self.width, self.height = grid_n, width, height
        self.ctx = A.VoxelGI(grid_n, width, height, A.MODE_NORTHSTAR, shadow_res=shadow_res, device=device, rank=rank, nranks=nranks, lib=lib)
        self.rows = row_ranges(height, nranks)[rank]
        self.ctx.set_trace_rows(*self.rows)
        if scene is not None:
            self.ctx.upload_scene(scene)

    def describe(self):
        if self.nranks == 1:
            return "1 GPU"
        return f"{self.nranks} GPUs: volume replicated per rank (no exchange), trace split into {self.nranks} row bands"

    def frame(self, voxel_cam, k):
        c = self.ctx
        c.voxelize(voxel_cam)
        c.inject(k)
        c.build_mips()
        c.trace_indirect(k)

    def comm_ms_per_frame(self):
        return 0.0

    def close(self):
        self.ctx.close()
