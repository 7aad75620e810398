"""GPU parity, north-star mode: conservative voxelization + normalise, injection and the six-direction mips
are byte/integer work and must be BIT-EXACT against the CPU oracle; the cone-traced image is floating point
through the texture unit and must be within north_star's 1e-2 relative L2 per image."""
import numpy as np
import pytest

import helpers as Hh
from final184_b200 import api as A
from final184_b200 import scene as S
from final184_b200.fixture import frame_inputs

pytestmark = pytest.mark.gpu


def _pipeline(cuda_lib, oracle_lib, scene, cams, n, w, h, shadow_res=512, flags=0, stop_after="trace"):
    g, o = Hh.make_pair(cuda_lib, oracle_lib, scene, grid_n=n, width=w, height=h, mode=A.MODE_NORTHSTAR, shadow_res=shadow_res, flags=flags)
    fi = frame_inputs(scene, cams["main"], cams["shadow"], w, h, shadow_res, 0)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], w, h, 0, True)
    for c in (g, o):
        Hh.upload_frame(c, fi)
        c.voxelize(cams["voxel"])
        if stop_after == "voxelize":
            continue
        c.inject(k)
        c.build_mips()
        if stop_after == "mips":
            continue
        c.trace_indirect(k)
    return g, o, k


@pytest.mark.parametrize("n", [32, 128])
def test_voxelize_normalise_bit_exact(cuda_lib, oracle_lib, proc_scene, cams, n):
    g, o, _ = _pipeline(cuda_lib, oracle_lib, proc_scene, cams, n, 64, 36, stop_after="voxelize")
    for cnt in (A.COUNTER_FRAGMENTS, A.COUNTER_OCCUPIED, A.COUNTER_BRICKS):
        assert g.counter(cnt) == o.counter(cnt) > 0, cnt
    ag, ao = g.readback(A.SLOT_VOX_ALBEDO), o.readback(A.SLOT_VOX_ALBEDO)
    assert np.array_equal(ag[..., 3], ao[..., 3]), "occupancy mask differs"
    assert np.array_equal(ag, ao), f"albedo: {np.count_nonzero((ag != ao).any(-1))} voxels differ"
    assert np.array_equal(g.readback(A.SLOT_VOX_NORMAL), o.readback(A.SLOT_VOX_NORMAL))
    # frame 2: accumulators were re-zeroed by normalise; same answer
    g.voxelize(cams["voxel"])
    assert np.array_equal(g.readback(A.SLOT_VOX_ALBEDO), ao)
    # frame 3 with half the triangles: bricks that became empty must be cleared
    half = proc_scene.n_tris // 2
    for c in (g, o):
        c.set_triangle_range(0, half)
        c.voxelize(cams["voxel"])
    assert np.array_equal(g.readback(A.SLOT_VOX_ALBEDO), o.readback(A.SLOT_VOX_ALBEDO))
    assert g.counter(A.COUNTER_OCCUPIED) == o.counter(A.COUNTER_OCCUPIED)


@pytest.mark.parametrize("flags", [0, A.FLAG_NO_TMA, A.FLAG_DENSE_MIPS])
def test_inject_and_mips_bit_exact(cuda_lib, oracle_lib, proc_scene, cams, flags):
    g, o, _ = _pipeline(cuda_lib, oracle_lib, proc_scene, cams, 128, 64, 36, flags=flags, stop_after="mips")
    rg, ro = g.readback(A.SLOT_RADIANCE), o.readback(A.SLOT_RADIANCE)
    assert (ro[..., :3].sum(-1) > 0).sum() > 100
    assert np.array_equal(rg, ro), f"radiance: {np.count_nonzero((rg != ro).any(-1))} voxels differ"
    mg, mo = g.readback(A.SLOT_MIPS), o.readback(A.SLOT_MIPS)
    assert np.array_equal(mg, mo), f"mips: {np.count_nonzero((mg != mo).any(-1))} texels differ"


def test_texture_side_storage_matches_linear_volumes(cuda_lib, oracle_lib, proc_scene, cams):
    """What the tracer's texture units sample (3D arrays written through surfaces, 16-byte surface stores) must be
    the same bytes as the linear volumes the oracle is compared with — over two frames, the second with half the
    triangles, so bricks that became empty are cleared at every level of the sparse mip builder."""
    n = 64
    g, o, k = _pipeline(cuda_lib, oracle_lib, proc_scene, cams, n, 64, 36, stop_after="mips")
    for frame in range(2):
        if frame == 1:
            for c in (g, o):
                c.set_triangle_range(0, proc_scene.n_tris // 2)
                c.voxelize(cams["voxel"]); c.inject(k); c.build_mips()
        rad, mips = o.readback(A.SLOT_RADIANCE), o.readback(A.SLOT_MIPS).reshape(-1, 4)
        assert np.array_equal(g.readback(A.SLOT_RADIANCE), rad)
        assert np.array_equal(g.readback(A.SLOT_MIPS).reshape(-1, 4), mips), f"frame {frame}"
        assert np.array_equal(g.read_array(-1, 0, n), rad), f"frame {frame}: level-0 array"
        off, m, lvl = 0, n // 2, 0
        while m >= 1:
            for d in range(6):
                want = mips[off + d * m ** 3: off + (d + 1) * m ** 3].reshape(m, m, m, 4)
                assert np.array_equal(g.read_array(d, lvl, m), want), f"frame {frame}: dir {d} level {lvl + 1}"
            off += 6 * m ** 3
            m //= 2
            lvl += 1


def test_cone_trace_within_tolerance(cuda_lib, oracle_lib, proc_scene, cams):
    g, o, _ = _pipeline(cuda_lib, oracle_lib, proc_scene, cams, 128, 160, 90)
    ig, io = g.readback(A.SLOT_INDIRECT_OUT).astype(np.float32), o.readback(A.SLOT_INDIRECT_OUT).astype(np.float32)
    assert np.isfinite(ig).all()
    assert io[..., :3].mean() > 1e-3
    err = Hh.rel_l2(ig[..., :3], io[..., :3])
    assert err <= 1e-2, err                                    # north_star: traced radiance within 1e-2 rel. L2
    assert np.array_equal(ig[..., 3], io[..., 3])              # -viewZ channel is exact arithmetic
    sg, so = g.counter(A.COUNTER_MARCH_STEPS), o.counter(A.COUNTER_MARCH_STEPS)
    assert abs(sg - so) <= 0.01 * so, (sg, so)


@pytest.mark.skipif(not S.sponza_available(), reason="Sponza pack not staged")
def test_sponza_pipeline(cuda_lib, oracle_lib, cams):
    sc = S.load_sponza()
    g, o, _ = _pipeline(cuda_lib, oracle_lib, sc, cams, 128, 160, 90, shadow_res=1024)
    assert g.counter(A.COUNTER_FRAGMENTS) == o.counter(A.COUNTER_FRAGMENTS)
    assert np.array_equal(g.readback(A.SLOT_VOX_ALBEDO), o.readback(A.SLOT_VOX_ALBEDO))
    assert np.array_equal(g.readback(A.SLOT_RADIANCE), o.readback(A.SLOT_RADIANCE))
    assert np.array_equal(g.readback(A.SLOT_MIPS), o.readback(A.SLOT_MIPS))
    ig, io = g.readback(A.SLOT_INDIRECT_OUT).astype(np.float32), o.readback(A.SLOT_INDIRECT_OUT).astype(np.float32)
    assert Hh.rel_l2(ig[..., :3], io[..., :3]) <= 1e-2


def test_frame_overlap_is_invisible(cuda_lib, proc_scene, cams):
    """Default single-GPU scheduling runs voxelize + normalise of frame f+1 on an internal stream beside the cone trace
    of frame f (and read-backs travel on a third stream).  Six frames are submitted back to back without a host
    sync, the triangle range changing every frame so each frame's volume differs and emptied bricks must be
    cleared; every image and the last volume must equal the F184_FLAG_NO_OVERLAP run bit for bit."""
    import torch
    n, w, h = 128, 320, 184
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], w, h, 512, 0)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], w, h, 0, True)
    T = proc_scene.n_tris
    ranges = [(0, T), (0, T // 2), (T // 3, T // 2), (0, T), (T // 2, T), (0, T // 4)]
    results = []
    for flags in (A.FLAG_NO_OVERLAP, 0):
        c = A.VoxelGI(lib=cuda_lib, grid_n=n, width=w, height=h, mode=A.MODE_NORTHSTAR, shadow_res=512, flags=flags)
        c.upload_scene(proc_scene)
        Hh.upload_frame(c, fi)
        nbytes = c.image_info(A.SLOT_INDIRECT_OUT).size_bytes
        hosts = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in ranges]
        for i, (first, count) in enumerate(ranges):
            c.set_triangle_range(first, count)
            c.voxelize(cams["voxel"]); c.inject(k); c.build_mips(); c.trace_indirect(k)
            c.readback_async_ptr(A.SLOT_INDIRECT_OUT, hosts[i].data_ptr(), nbytes)
            if i:
                c.readback_wait(1)
        c.sync()
        results.append(([t.numpy().copy() for t in hosts], c.readback(A.SLOT_VOX_ALBEDO), c.readback(A.SLOT_MIPS),
                        c.counter(A.COUNTER_FRAGMENTS), c.counter(A.COUNTER_OCCUPIED)))
        c.close()
    (ia, va, ma, fa, oa), (ib, vb, mb, fb, ob) = results
    assert fa == fb > 0 and oa == ob > 0
    assert np.array_equal(va, vb) and np.array_equal(ma, mb)
    for i, (x, y) in enumerate(zip(ia, ib)):
        assert np.array_equal(x, y), f"frame {i}"
    assert not np.array_equal(ia[0], ia[1])       # the frames really differ


def test_probe_batch_views(cuda_lib, oracle_lib, proc_scene):
    """BASELINE configs[4] at test size: 6 probe views stacked in one context (f184_trace_views, independent views on
    alternating streams) — each view within the cone-trace tolerance of the oracle, the sample count within 1 %, and
    BIT-identical to the same view traced alone as an ordinary image on the GPU; then a history frame."""
    from final184_b200 import dist as D
    import test_probe_views as P
    n, vs, sh, nv = 64, 96, 256, 6
    cams_, views, per_view, shadow, ks = P.batch_inputs(proc_scene, nv, vs, sh, stride=11)
    res = {}
    for name, lib in (("gpu", cuda_lib), ("cpu", oracle_lib)):
        b = D.ProbeBatch(n, vs, nv, shadow_res=sh, scene=proc_scene, voxel_cam=cams_["voxel"], lib=lib)
        b.upload_views(per_view, shadow)
        b.frame(cams_["voxel"], ks)
        res[name] = (dict(b.own_views()), b.ctx.counter(A.COUNTER_MARCH_STEPS), b)
    (vg, sg, bg), (vo, so, bo) = res["gpu"], res["cpu"]
    assert abs(sg - so) <= 0.01 * so, (sg, so)
    for v in range(nv):
        ig, io = vg[v].astype(np.float32), vo[v].astype(np.float32)
        assert np.isfinite(ig).all() and np.array_equal(ig[..., 3], io[..., 3])
        if io[..., :3].mean() > 1e-3:
            assert Hh.rel_l2(ig[..., :3], io[..., :3]) <= 1e-2, v
    one = A.VoxelGI(n, vs, vs, A.MODE_NORTHSTAR, shadow_res=sh, lib=cuda_lib)
    one.upload_scene(proc_scene)
    one.upload(A.SLOT_SHADOW, shadow)
    one.voxelize(cams_["voxel"]); one.inject(ks[0]); one.build_mips()
    for v in range(nv):
        for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material")):
            one.upload(slot, per_view[v][key])
        one.trace_indirect(ks[v])
        assert np.array_equal(one.readback(A.SLOT_INDIRECT_OUT).view(np.uint16), vg[v].view(np.uint16)), f"view {v} alone vs in the batch"
    # history frame: every view reprojects into its own rows
    ks2 = [A.trace_constants_c(v, cams_["shadow"], cams_["voxel"], vs, vs, 1, False) for v in views]
    for b in (bg, bo):
        b.ctx.copy_indirect_to_history()
        b.ctx.trace_views(ks2, vs)
    g2, o2 = dict(bg.own_views()), dict(bo.own_views())
    for v in range(nv):
        if o2[v][..., :3].astype(np.float32).mean() > 1e-3:
            assert Hh.rel_l2(g2[v][..., :3].astype(np.float32), o2[v][..., :3].astype(np.float32)) <= 1e-2, v
    bg.close(); bo.close(); one.close()


def test_row_selective_transfers_gpu(cuda_lib):
    """Same contract as tests/test_dist_cpu.py::test_row_selective_transfers on the device (one cudaMemcpy2DAsync per transfer,
    double-buffered upload slot, staged read-back)."""
    import torch
    w, h = 24, 44
    c = A.VoxelGI(32, w, h, A.MODE_NORTHSTAR, shadow_res=16, lib=cuda_lib)
    rng = np.random.default_rng(0)
    for first, stride, y0, y1 in ((1, 4, 0, 0xffffffff), (0, 2, 0, 0xffffffff), (1, 2, 8, 40), (0, 1, 0, 0xffffffff)):
        old, new = (rng.random((h, w), dtype=np.float32) for _ in range(2))
        c.set_trace_tiles(first, stride); c.set_trace_rows(y0, y1)
        mine = ((np.arange(h) // 8) % stride == first) & (np.arange(h) >= y0) & (np.arange(h) < min(y1, h))
        # the upload slot is double-buffered: fill BOTH buffers with `old` so the rows that do not travel are defined
        c.upload(A.SLOT_DEPTH, old); c.upload(A.SLOT_DEPTH, old)
        pin = torch.from_numpy(new).pin_memory()
        c.upload_ptr(A.SLOT_DEPTH, pin.data_ptr(), new.nbytes, rows=True)
        got = c.readback(A.SLOT_DEPTH)
        assert np.array_equal(got[mine], new[mine]) and np.array_equal(got[~mine], old[~mine]), (first, stride)
        host = torch.full((h, w), -1.0, dtype=torch.float32).pin_memory()
        c.readback_async_ptr(A.SLOT_DEPTH, host.data_ptr(), new.nbytes, rows=True)
        c.sync()
        hn = host.numpy()
        assert np.array_equal(hn[mine], got[mine]) and (hn[~mine] == -1.0).all()
    c.close()


def test_secondary_passes_fast_within_tolerance_and_exact_on_request(cuda_lib, oracle_lib, proc_scene, cams):
    """Mode N's GTAO (+ 4x4 blur) and bilateral blur run on the hardware's own units (MUFU rcp / rsqrt / sin / cos / ex2, FMA): the
    oracle's reference-faithful restatement bounds them within north_star's 1e-2 relative L2 per image; F184_FLAG_EXACT_SECONDARY
    selects the pinned kernels, which equal the oracle bit for bit (as in mode R)."""
    w, h = 320, 184
    for flags, exact in ((0, False), (A.FLAG_EXACT_SECONDARY, True)):
        g, o, k = _pipeline(cuda_lib, oracle_lib, proc_scene, cams, 64, w, h, flags=flags)
        for c in (g, o):
            c.gtao(cams["main"])
            c.blur_indirect(k)
        ao_g, ao_o = g.readback(A.SLOT_AO_OUT).astype(np.float32)[..., 0], o.readback(A.SLOT_AO_OUT).astype(np.float32)[..., 0]
        raw_g, raw_o = g.readback(A.SLOT_AO_RAW).astype(np.float32)[..., 0], o.readback(A.SLOT_AO_RAW).astype(np.float32)[..., 0]
        # the blur's input is the GPU's own traced image (itself within tolerance of the oracle's): compare blur(GPU image) on both sides
        o.upload(A.SLOT_INDIRECT_OUT, g.readback(A.SLOT_INDIRECT_OUT))
        o.blur_indirect(k)
        bl_g, bl_o = g.readback(A.SLOT_INDIRECT_FINAL).astype(np.float32)[..., :3], o.readback(A.SLOT_INDIRECT_FINAL).astype(np.float32)[..., :3]
        assert ao_o.std() > 1e-3 and bl_o.mean() > 1e-3
        if exact:
            assert np.array_equal(ao_g, ao_o) and np.array_equal(raw_g, raw_o) and np.array_equal(bl_g, bl_o)
        else:
            e_raw, e_ao, e_bl = Hh.rel_l2(raw_g, raw_o), Hh.rel_l2(ao_g, ao_o), Hh.rel_l2(bl_g, bl_o)
            print(f"fast secondary passes: rel. L2 gtao {e_raw:.2e}, gtao+blur {e_ao:.2e}, bilateral blur {e_bl:.2e}")
            assert e_raw <= 1e-2 and e_ao <= 1e-2 and e_bl <= 1e-2, (e_raw, e_ao, e_bl)
        g.close(); o.close()
