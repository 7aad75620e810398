"""GPU tests at BASELINE.json's sizes.  Where the oracle still finishes in seconds (C2: Sponza 256^3) the comparison is
direct; at 512^3 / 1024^3 the checks are size-independent properties of the domain: determinism, additivity of partial
volumes (the multi-GPU exchange), sparse == dense mip chain, band/tile composition of the trace."""
import numpy as np
import pytest

import helpers as Hh
from final184_b200 import api as A
from final184_b200 import scene as S
from final184_b200.fixture import frame_inputs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not S.sponza_available(), reason="Sponza pack not staged")]


@pytest.fixture(scope="module")
def sponza():
    return S.load_sponza()


def test_c2_sponza_256_volume_stages_bit_exact(cuda_lib, oracle_lib, sponza, cams):
    """configs[1]: Sponza 256^3 — voxelize+normalise, inject, mips against the oracle, byte for byte."""
    n = 256
    g, o = Hh.make_pair(cuda_lib, oracle_lib, sponza, grid_n=n, width=64, height=36, mode=A.MODE_NORTHSTAR, shadow_res=2048)
    fi = frame_inputs(sponza, cams["main"], cams["shadow"], 64, 36, 2048, 0)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], 64, 36, 0, True)
    for c in (g, o):
        c.upload(A.SLOT_SHADOW, fi["shadow"])
        c.voxelize(cams["voxel"]); c.inject(k); c.build_mips()
    for cnt in (A.COUNTER_FRAGMENTS, A.COUNTER_OCCUPIED, A.COUNTER_BRICKS):
        assert g.counter(cnt) == o.counter(cnt) > 0
    for slot in (A.SLOT_VOX_ALBEDO, A.SLOT_VOX_NORMAL, A.SLOT_RADIANCE, A.SLOT_MIPS):
        assert np.array_equal(g.readback(slot), o.readback(slot)), slot


def test_c3_sponza_512_properties(cuda_lib, sponza, cams):
    """The bench workload's volume: deterministic, additive over triangle ranges, sparse mips == dense TMA mips."""
    n, W, H = 512, 256, 144
    g = A.VoxelGI(grid_n=n, width=W, height=H, mode=A.MODE_NORTHSTAR, lib=cuda_lib)
    d = A.VoxelGI(grid_n=n, width=W, height=H, mode=A.MODE_NORTHSTAR, lib=cuda_lib, flags=A.FLAG_DENSE_MIPS)
    fi = frame_inputs(sponza, cams["main"], cams["shadow"], W, H, 2048, 0)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    for c in (g, d):
        c.upload_scene(sponza)
        Hh.upload_frame(c, fi)
        c.voxelize(cams["voxel"]); c.inject(k); c.build_mips()
    alb, mips = g.readback(A.SLOT_VOX_ALBEDO), g.readback(A.SLOT_MIPS)
    assert g.counter(A.COUNTER_FRAGMENTS) == 4662509 and g.counter(A.COUNTER_OCCUPIED) == 2940738      # pinned: order-independent integers
    assert (alb[..., 3] == 255).sum() == 2940738
    assert np.array_equal(mips, d.readback(A.SLOT_MIPS)), "sparse brick mips differ from the dense TMA chain"
    d.close()
    # frame 2 == frame 1 (atomics in any order, accumulators re-zeroed by normalise)
    g.voxelize(cams["voxel"]); g.inject(k); g.build_mips()
    assert np.array_equal(g.readback(A.SLOT_VOX_ALBEDO), alb) and np.array_equal(g.readback(A.SLOT_MIPS), mips)
    # partial volumes add up: three triangle ranges accumulated one after the other, one normalise
    T = sponza.n_tris
    for first, count in ((0, T // 3), (T // 3, T // 2), (T // 3 + T // 2, T)):
        g.set_triangle_range(first, count)
        g.voxelize_accumulate(cams["voxel"])
    g.normalise()
    assert np.array_equal(g.readback(A.SLOT_VOX_ALBEDO), alb)
    g.set_triangle_range(0, 0xffffffff)
    # the trace composes from bands and from interleaved tiles
    g.inject(k); g.build_mips(); g.trace_indirect(k)
    img = g.readback(A.SLOT_INDIRECT_OUT).copy()
    assert np.isfinite(img.astype(np.float32)).all() and img[..., :3].astype(np.float32).mean() > 1e-2
    out = np.zeros_like(img)
    for first in range(4):
        g.set_trace_tiles(first, 4); g.trace_indirect(k)
        m = (np.arange(H) // 8) % 4 == first
        out[m] = g.readback(A.SLOT_INDIRECT_OUT)[m]
    assert np.array_equal(out.view(np.uint16), img.view(np.uint16))


def test_c4_tiled_sponza_1024_properties(cuda_lib, sponza):
    """configs[3]: 8 x tiled Sponza (2.1 M triangles) at 1024^3 under the C4 voxel camera: deterministic and additive."""
    sc = S.tile_scene(sponza, S.C4_OFFSETS)
    assert sc.n_tris == 8 * 262267
    cam = S.fixture_constants("voxel_c4")
    g = A.VoxelGI(grid_n=1024, width=64, height=36, mode=A.MODE_NORTHSTAR, lib=cuda_lib)
    g.upload_scene(sc)
    g.voxelize(cam)
    frags, occ = g.counter(A.COUNTER_FRAGMENTS), g.counter(A.COUNTER_OCCUPIED)
    assert frags > 8 * 1_000_000 and occ > 1_000_000
    nrm = g.readback(A.SLOT_VOX_NORMAL)
    assert (np.abs(nrm[..., :3].astype(np.int32)).max(-1) > 0).sum() <= occ
    T = sc.n_tris
    for first, count in ((0, T // 2), (T // 2, T)):
        g.set_triangle_range(first, count)
        g.voxelize_accumulate(cam)
    g.normalise()
    assert g.counter(A.COUNTER_OCCUPIED) == occ
    assert np.array_equal(g.readback(A.SLOT_VOX_NORMAL), nrm)
