"""GPU tests at BASELINE.json's sizes.  C2 (256^3) and C3 (512^3, the headline) are compared with the oracle directly — every
volume byte for byte, and a band of the full-resolution cone-traced image within north_star's tolerance; C4 (1024^3) against
golden vectors the oracle produced once (tests/golden/c4_oracle.npz, tools/gen_c4_golden.py: the oracle needs ~48 GB and minutes
there).  Beside them the size-independent properties of the domain: determinism, additivity of partial volumes (the multi-GPU
exchange), sparse == dense mip chain, band/tile composition of the trace."""
import os
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


def _trace_band_vs_oracle(g, o, k, y0, y1, tol=1e-2):
    """rows [y0, y1) of the cone-traced image on both sides: north_star's 1e-2 relative L2 per image, applied to the band"""
    for c in (g, o):
        c.set_trace_rows(y0, y1)
        c.trace_indirect(k)
    ig = g.readback(A.SLOT_INDIRECT_OUT)[y0:y1].astype(np.float32)
    io = o.readback(A.SLOT_INDIRECT_OUT)[y0:y1].astype(np.float32)
    assert np.isfinite(ig).all() and io[..., :3].mean() > 1e-3
    err = Hh.rel_l2(ig[..., :3], io[..., :3])
    sg, so = g.counter(A.COUNTER_MARCH_STEPS), o.counter(A.COUNTER_MARCH_STEPS)
    print(f"trace band rows [{y0}, {y1}): rel. L2 {err:.2e}, cone-samples {sg} vs {so}")
    assert err <= tol, err
    assert np.allclose(ig[..., 3], io[..., 3], rtol=2e-3, atol=1e-3)          # -viewZ
    assert abs(sg - so) <= 0.005 * so, (sg, so)
    for c in (g, o):
        c.set_trace_rows(0, 0xffffffff)
    return err


def test_c2_sponza_256_trace_1080p_vs_oracle(cuda_lib, oracle_lib, sponza, cams):
    """configs[1] at its own resolution: volumes byte for byte, then 256 rows of the 1920 x 1080 cone trace against the oracle."""
    n, W, H = 256, 1920, 1080
    g, o = Hh.make_pair(cuda_lib, oracle_lib, sponza, grid_n=n, width=W, height=H, mode=A.MODE_NORTHSTAR, shadow_res=2048)
    fi = frame_inputs(sponza, cams["main"], cams["shadow"], W, H, 2048, 0)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    for c in (g, o):
        Hh.upload_frame(c, fi)
        c.voxelize(cams["voxel"]); c.inject(k); c.build_mips()
    for slot in (A.SLOT_RADIANCE, A.SLOT_MIPS):
        assert np.array_equal(g.readback(slot), o.readback(slot)), slot
    _trace_band_vs_oracle(g, o, k, 400, 656)
    g.close(); o.close()


def test_c3_sponza_512_every_volume_and_a_4k_band_vs_oracle(cuda_lib, oracle_lib, sponza, cams):
    """The headline configuration itself (Sponza 512^3, 3840 x 2160) against the oracle: counters, mean albedo, mean normal, injected
    radiance and the whole six-direction mip chain byte for byte; a 64-row band of the 3840-wide cone-traced image within 1e-2."""
    n, W, H = 512, 3840, 2160
    g, o = Hh.make_pair(cuda_lib, oracle_lib, sponza, grid_n=n, width=W, height=H, mode=A.MODE_NORTHSTAR, shadow_res=2048)
    fi = frame_inputs(sponza, cams["main"], cams["shadow"], W, H, 2048, 0)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    for c in (g, o):
        Hh.upload_frame(c, fi)
        c.voxelize(cams["voxel"]); c.inject(k); c.build_mips()
    for cnt in (A.COUNTER_FRAGMENTS, A.COUNTER_OCCUPIED, A.COUNTER_BRICKS):
        assert g.counter(cnt) == o.counter(cnt) > 0
    assert g.counter(A.COUNTER_FRAGMENTS) == 4662509 and g.counter(A.COUNTER_OCCUPIED) == 2940738 and g.counter(A.COUNTER_BRICKS) == 30940
    for slot in (A.SLOT_VOX_ALBEDO, A.SLOT_VOX_NORMAL, A.SLOT_RADIANCE, A.SLOT_MIPS):
        assert np.array_equal(g.readback(slot), o.readback(slot)), slot
    _trace_band_vs_oracle(g, o, k, 1048, 1112)
    g.close(); o.close()


def test_c4_tiled_sponza_1024_vs_oracle_golden(cuda_lib, sponza):
    """configs[3] against the oracle's pinned answers (tests/golden/c4_oracle.npz): the counters of the whole 1024^3 volume and the
    densest 128^3 sub-cube of mean albedo, mean normal, injected radiance and of every level of the six-direction chain."""
    from final184_b200.fixture import Fixture
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c4_oracle.npz"))
    N, SUB = 1024, 128
    x0, y0, z0 = (int(v) for v in z["origin"])
    sc = S.tile_scene(sponza, S.C4_OFFSETS)
    cam = S.fixture_constants("voxel_c4")
    main_, shadow_ = S.fixture_constants("main"), S.fixture_constants("shadow")
    k = A.trace_constants_c(main_, shadow_, cam, 64, 36, 0, True)
    g = A.VoxelGI(grid_n=N, width=64, height=36, mode=A.MODE_NORTHSTAR, lib=cuda_lib)
    g.upload_scene(sc)
    g.upload(A.SLOT_SHADOW, Fixture(sc).shadow(shadow_, 2048))
    g.voxelize(cam); g.inject(k); g.build_mips()
    got = [g.counter(A.COUNTER_FRAGMENTS), g.counter(A.COUNTER_OCCUPIED), g.counter(A.COUNTER_BRICKS)]
    assert got == [int(v) for v in z["counters"]], (got, z["counters"])
    cut = lambda v, n, s: v[z0 * n // N:z0 * n // N + s, y0 * n // N:y0 * n // N + s, x0 * n // N:x0 * n // N + s]
    for slot, key in ((A.SLOT_VOX_ALBEDO, "albedo"), (A.SLOT_VOX_NORMAL, "normal"), (A.SLOT_RADIANCE, "radiance")):
        assert np.array_equal(cut(g.readback(slot), N, SUB), z[key]), key
    assert (z["albedo"][..., 3] != 0).sum() > 100000 and z["radiance"][..., :3].max() > 0
    mips = g.readback(A.SLOT_MIPS).reshape(-1, 4)
    off, n, lvl = 0, N // 2, 1
    while n >= 1:
        s_ = max(1, SUB * n // N)
        for d in range(6):
            v = mips[off + d * n ** 3: off + (d + 1) * n ** 3].reshape(n, n, n, 4)
            assert np.array_equal(cut(v, n, s_), z[f"mip{lvl}_d{d}"]), (lvl, d)
        off += 6 * n ** 3
        n //= 2; lvl += 1
    g.close()


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


def test_c5_probe_batch_64_views_512(cuda_lib, sponza, cams):
    """configs[4]: 64 probe views (512 x 512, FovY 90, Halton positions — SURVEY.md §8(d)) cone-traced against the
    Sponza 512^3 volume in one context.  Size-independent properties: the batch equals the same views traced alone
    (bit for bit), the cone-sample counter is the sum over the views, tracing a sub-range leaves the other rows alone."""
    from final184_b200 import dist as D
    from final184_b200.fixture import Fixture
    n, vs, nv = 512, 512, 64
    views = [S.fixture_constants(f"probe{i:02d}") for i in range(nv)]
    fx = Fixture(sponza)
    per_view = [fx.gbuffer(v, vs, vs, 0) for v in views]
    shadow = fx.shadow(cams["shadow"], 2048)
    ks = [A.trace_constants_c(v, cams["shadow"], cams["voxel"], vs, vs, 0, True) for v in views]
    b = D.ProbeBatch(n, vs, nv, scene=sponza, voxel_cam=cams["voxel"], lib=cuda_lib)
    b.upload_views(per_view, shadow)
    b.frame(cams["voxel"], ks)
    img = b.ctx.readback(A.SLOT_INDIRECT_OUT).copy()
    total = b.ctx.counter(A.COUNTER_MARCH_STEPS)
    f = img.astype(np.float32)
    assert np.isfinite(f).all() and f[..., :3].mean() > 1e-2
    ms = b.ctx.stage_ms(A.STAGE_TRACE)
    print(f"C5: 64 x 512^2 views, {total / 1e6:.0f} M cone-samples, trace {ms:.2f} ms = {total / ms / 1e6:.1f} Gcone-samples/s")
    one = A.VoxelGI(n, vs, vs, A.MODE_NORTHSTAR, lib=cuda_lib)
    one.upload_scene(sponza)
    one.upload(A.SLOT_SHADOW, shadow)
    one.voxelize(cams["voxel"]); one.inject(ks[0]); one.build_mips()
    part = 0
    for v in (0, 17, 63):
        for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material")):
            one.upload(slot, per_view[v][key])
        one.trace_indirect(ks[v])
        assert np.array_equal(one.readback(A.SLOT_INDIRECT_OUT).view(np.uint16), img[v * vs:(v + 1) * vs].view(np.uint16)), f"view {v}"
        part += one.counter(A.COUNTER_MARCH_STEPS)
    b.ctx.trace_views(ks, vs, first=0, count=1); s0 = b.ctx.counter(A.COUNTER_MARCH_STEPS)
    b.ctx.trace_views(ks, vs, first=17, count=1); s17 = b.ctx.counter(A.COUNTER_MARCH_STEPS)
    b.ctx.trace_views(ks, vs, first=63, count=1); s63 = b.ctx.counter(A.COUNTER_MARCH_STEPS)
    assert s0 + s17 + s63 == part and 0 < part < total
    assert np.array_equal(b.ctx.readback(A.SLOT_INDIRECT_OUT), img)          # re-tracing single views reproduced them in place
    b.close(); one.close()
