"""The CPU oracle, reference-faithful mode, against KNOWN ANSWERS worked out by hand from the reference's shader text
(Pipelang/Internal/main.lua:83-144 VoxelGS, :242-275 VoxelPS) — independent of the oracle's code —
plus the properties the host logic relies on.  The reference ships no golden vectors for this path; the bit-for-bit
pin against the reference's own shader text is tests/test_refshader_pin.py, and these hand cases pin what that cannot:
the fixed-function rasteriser between VoxelGS and VoxelPS, which the reference leaves to the Vulkan driver."""
import numpy as np
import pytest

import cpu_helpers as Hc
from final184_b200 import api as A
from final184_b200 import scene as S
from final184_b200.fixture import frame_inputs

N = 32


def voxelize(oracle_lib, sc, n=N):
    o = A.VoxelGI(grid_n=n, width=16, height=16, mode=A.MODE_REFERENCE, lib=oracle_lib)
    o.upload_scene(sc)
    o.voxelize(S.fixture_constants("voxel"))
    return o, o.readback(A.SLOT_VOXELS)


def occupied(v):
    z, y, x = np.nonzero(v[..., 0])
    return set(zip(x.tolist(), y.tolist(), z.tolist()))


def test_floor_quad_orientation_z(oracle_lib):
    """World-Y normal = view-space Z dominant (orientation 0): voxel = (px, py, zn*(N-1)).
    x,z in [-4,4] -> px,py in [12,20] -> pixels 12..19; zn = 15/19 -> trunc(0.78947*31) = 24."""
    sc = Hc.quad_scene([((-4, 0, -4), (8, 0, 0), (0, 0, 8), (0, 1, 0))])
    o, v = voxelize(oracle_lib, sc)
    assert occupied(v) == {(i, j, 24) for i in range(12, 20) for j in range(12, 20)}
    assert o.counter(A.COUNTER_FRAGMENTS) == 64                       # top-left rule: the shared diagonal is covered once
    vals = v[v[..., 0] != 0]
    assert (vals[:, 0] == 0xFFFF).all()                               # white: (31<<11)|(63<<5)|31
    assert (vals[:, 1] == ((15 << 11) | (63 << 5) | 15)).all()        # normal (0,1,0): floor(n*16+15), floor(n*32+31)


def test_wall_orientation_x(oracle_lib):
    """World-X normal (orientation 1): raster x = zn*N, voxel = ((xn/2+.5)*(N-1), py, zn*N).
    x = 2 -> trunc(0.5625*31) = 17; y in [0,4] -> zn*32 in [18.53,25.26] -> 19..24; z in [-4,4] -> py 12..19."""
    sc = Hc.quad_scene([((2, 0, -4), (0, 4, 0), (0, 0, 8), (1, 0, 0))])
    o, v = voxelize(oracle_lib, sc)
    assert occupied(v) == {(17, j, i) for j in range(12, 20) for i in range(19, 25)}
    assert (v[v[..., 0] != 0][:, 1] == ((31 << 11) | (31 << 5) | 15)).all()


def test_wall_orientation_y_keeps_the_reference_quirk(oracle_lib):
    """World-Z normal (orientation 2): raster y = (1-zn)*N, voxel = (px, (yn/2+.5)*(N-1), (N-1) - py): the -1 quirk of
    VoxelPS (main.lua:262-266) is reproduced, not fixed.  z = -2 -> yn = .125 -> 17; py centres 7.5..12.5 -> z = 30-j."""
    sc = Hc.quad_scene([((-4, 0, -2), (8, 0, 0), (0, 4, 0), (0, 0, 1))])
    o, v = voxelize(oracle_lib, sc)
    assert occupied(v) == {(i, 17, 30 - j) for i in range(12, 20) for j in range(7, 13)}
    assert (v[v[..., 0] != 0][:, 1] == ((15 << 11) | (31 << 5) | 31)).all()


def test_black_albedo_counts_as_written_but_reads_as_empty(oracle_lib):
    """indirect.frag:156 treats .r == 0 as empty; a black texel packs to 0."""
    tex = np.zeros((4, 4, 4), np.uint8); tex[..., 3] = 255
    sc = Hc.quad_scene([((-4, 0, -4), (8, 0, 0), (0, 0, 8), (0, 1, 0))], tex=tex)
    o, v = voxelize(oracle_lib, sc)
    assert o.counter(A.COUNTER_FRAGMENTS) == 64 and not occupied(v)
    assert (v[24, 12:20, 12:20, 1] != 0).all()                        # the normal half-word was stored


def test_alpha_cutout_discards(oracle_lib):
    tex = np.full((4, 4, 4), 255, np.uint8); tex[..., 3] = 10         # 10/255 < 0.05 (main.lua:199)
    sc = Hc.quad_scene([((-4, 0, -4), (8, 0, 0), (0, 0, 8), (0, 1, 0))], tex=tex)
    o, v = voxelize(oracle_lib, sc)
    assert o.counter(A.COUNTER_FRAGMENTS) == 0 and not v.any()


def test_depth_clip_drops_geometry_outside_the_volume(oracle_lib):
    sc = Hc.quad_scene([((-4, 16.5, -4), (8, 0, 0), (0, 0, 8), (0, 1, 0))])     # zn = (15-16.5)/19 < 0
    o, v = voxelize(oracle_lib, sc)
    assert o.counter(A.COUNTER_FRAGMENTS) == 0


def test_last_writer_wins_in_draw_order(oracle_lib):
    """Two coincident floors with different normals: the later triangle's value stays."""
    sc = Hc.quad_scene([((-4, 0, -4), (8, 0, 0), (0, 0, 8), (0, 1, 0)), ((-4, 0, -4), (8, 0, 0), (0, 0, 8), (0, -1, 0))])
    o, v = voxelize(oracle_lib, sc)
    assert (v[v[..., 0] != 0][:, 1] == ((15 << 11) | (0 << 5) | 15)).all()     # floor(-1*32+31) = 0 after saturation


def test_triangle_ranges_compose(oracle_lib, proc_scene):
    cam = S.fixture_constants("voxel")
    o = A.VoxelGI(grid_n=64, width=16, height=16, mode=A.MODE_REFERENCE, lib=oracle_lib)
    o.upload_scene(proc_scene)
    o.voxelize(cam); full = o.readback(A.SLOT_VOXELS).copy(); f_all = o.counter(A.COUNTER_FRAGMENTS)
    h = proc_scene.n_tris // 2
    o.set_triangle_range(0, h); o.voxelize(cam); a = o.readback(A.SLOT_VOXELS).copy(); fa = o.counter(A.COUNTER_FRAGMENTS)
    o.set_triangle_range(h, proc_scene.n_tris - h); o.voxelize(cam); b = o.readback(A.SLOT_VOXELS).copy(); fb = o.counter(A.COUNTER_FRAGMENTS)
    assert fa + fb == f_all
    wrote_b = (b != 0).any(-1, keepdims=True)
    merged = np.where(wrote_b, b, a)
    assert ((merged != full).any(-1)).sum() <= 8           # only a later fragment that packs to (0,0) is invisible in b


@pytest.fixture(scope="module")
def traced(oracle_lib, proc_scene, cams):
    W, H, n = 96, 54, 64
    o = A.VoxelGI(grid_n=n, width=W, height=H, mode=A.MODE_REFERENCE, shadow_res=256, lib=oracle_lib)
    o.upload_scene(proc_scene)
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], W, H, 256, 0, cache=False)
    for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_SHADOW, "shadow")):
        o.upload(slot, fi[key])
    o.voxelize(cams["voxel"])
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    o.trace_indirect(k)
    return o, k, fi, o.readback(A.SLOT_INDIRECT_OUT).copy(), (W, H)


def test_trace_output_contract(traced):
    o, k, fi, img, (W, H) = traced
    rgb = img[..., :3].astype(np.float32)
    assert np.isfinite(rgb).all() and (rgb >= 0).all() and (rgb <= 16).all()       # clamp(.., 0, 16), indirect.frag:240
    assert o.counter(A.COUNTER_MARCH_STEPS) <= 8 * 60 * W * H                       # <= 8 marches x 60 steps per pixel
    # alpha = -viewZ (indirect.frag:242): positive for everything in front of the camera
    assert (img[..., 3].astype(np.float32)[fi["depth"] < 1.0] > 0).all()
    assert rgb[fi["depth"] < 1.0].mean() > 1e-3


def test_trace_is_deterministic_and_bands_compose(traced):
    """Row bands traced separately reassemble the full image bit for bit: the screen split of SURVEY.md §8(e)."""
    o, k, fi, img, (W, H) = traced
    o.trace_indirect(k)
    assert np.array_equal(o.readback(A.SLOT_INDIRECT_OUT).view(np.uint16), img.view(np.uint16))
    out = np.zeros_like(img)
    for y0, y1 in ((0, 16), (16, 40), (40, H)):
        o.set_trace_rows(y0, y1); o.trace_indirect(k)
        out[y0:y1] = o.readback(A.SLOT_INDIRECT_OUT)[y0:y1]
    o.set_trace_rows(0, 0xffffffff)
    assert np.array_equal(out.view(np.uint16), img.view(np.uint16))


def test_temporal_blend_pulls_towards_history(traced):
    """Frame k > 0 with a supplied history: identical camera => w = 0.95 where depth agrees (indirect.frag:225-240)."""
    o, k, fi, img0, (W, H) = traced
    o.copy_indirect_to_history()
    k1 = A.trace_constants_c(S.fixture_constants("main"), S.fixture_constants("shadow"), S.fixture_constants("voxel"), W, H, 1, False)
    o.trace_indirect(k1)
    img1 = o.readback(A.SLOT_INDIRECT_OUT).astype(np.float32)
    k1r = A.trace_constants_c(S.fixture_constants("main"), S.fixture_constants("shadow"), S.fixture_constants("voxel"), W, H, 1, True)
    o.trace_indirect(k1r)
    fresh = o.readback(A.SLOT_INDIRECT_OUT).astype(np.float32)
    m = fi["depth"] < 1.0
    d_hist = np.abs(img1[..., :3] - img0[..., :3].astype(np.float32))[m].mean()
    d_fresh = np.abs(img1[..., :3] - fresh[..., :3])[m].mean()
    assert d_hist < 0.25 * d_fresh


def test_gtao_range_and_blur_is_a_4x4_mean(oracle_lib, proc_scene, cams):
    W, H = 64, 36
    o = A.VoxelGI(grid_n=32, width=W, height=H, mode=A.MODE_REFERENCE, shadow_res=64, lib=oracle_lib)
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], W, H, 64, 0, cache=False)
    o.upload(A.SLOT_DEPTH, fi["depth"]); o.upload(A.SLOT_NORMALS, fi["normals"])
    o.gtao(cams["main"])
    raw, out = o.readback(A.SLOT_AO_RAW).astype(np.float32)[..., 0], o.readback(A.SLOT_AO_OUT).astype(np.float32)[..., 0]
    assert np.isfinite(raw).all() and raw.min() >= 0.0 and raw.max() <= 1.5       # gtao.frag does not clamp; fastAcos overshoots a little
    # GTAO/blur.frag:12-27: mean of the 4x4 neighbourhood [x-1, x+2] x [y-1, y+2] (wrap addressing at the border)
    y, x = 10, 20
    assert abs(out[y, x] - raw[y - 1:y + 3, x - 1:x + 3].mean()) < 2e-3


def test_deferred_lighting_properties(oracle_lib, proc_scene):
    """lighting_deferred (aggregateLights.frag) on the oracle: no lights -> (0, 0, 0, 1); light lists add up; a fully metallic
    material has no diffuse term, so roughness 1 + metallic 1 (every Sponza material) leaves only the weak specular lobe."""
    cams = {n: S.fixture_constants(n) for n in ("main", "shadow", "voxel")}
    w, h, sh = 64, 32, 256
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], w, h, sh, 0, cache=False)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], w, h, 0, True)
    c = A.VoxelGI(grid_n=32, width=w, height=h, mode=A.MODE_REFERENCE, shadow_res=sh, lib=oracle_lib)
    for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_SHADOW, "shadow")):
        c.upload(slot, fi[key])
    f32 = lambda: c.readback(A.SLOT_LIGHTING).astype(np.float32)

    def run(material, point, directional):
        c.upload(A.SLOT_MATERIAL, material)
        c.lighting_deferred(k, A.light_list_c(point), A.light_list_c(directional))
        return f32()

    dielectric = np.zeros((h, w, 4), np.uint8); dielectric[..., 1] = 128
    metal = dielectric.copy(); metal[..., 2] = 255
    none = run(dielectric, [], [])
    assert not none[..., :3].any() and (none[..., 3] == 1).all()
    p1, p2 = ((2.0, 1.0, 0.5), (0.0, 2.0, 0.0)), ((0.5, 3.0, 1.0), (3.0, 4.0, -2.0))
    a, b, ab = run(dielectric, [p1], []), run(dielectric, [p2], []), run(dielectric, [p1, p2], [])
    assert a[..., :3].mean() > 1e-3 and b[..., :3].mean() > 1e-3
    assert np.allclose(ab[..., :3], a[..., :3] + b[..., :3], rtol=4e-3, atol=1e-4)          # one fp16 rounding apart
    m = run(metal, [p1], [])
    assert m[..., :3].mean() < 0.5 * a[..., :3].mean()                                      # the diffuse term is gone
    c.close()
    # the sun: an open floor is lit; the same floor under a roof is in shadow everywhere (12 x 4 shadow taps all fail)
    sun = ((4.4, 3.72, 3.24), tuple(k.sun.position))
    floor = ((-12, 0, -12), (24, 0, 0), (0, 0, 24), (0, 1, 0))
    roof = ((-14, 9, -14), (28, 0, 0), (0, 0, 28), (0, 1, 0))
    means = []
    for quads in ([floor], [floor, roof]):
        sc = Hc.quad_scene(quads)
        fi2 = frame_inputs(sc, cams["main"], cams["shadow"], w, h, sh, 0, cache=False)
        c2 = A.VoxelGI(grid_n=32, width=w, height=h, mode=A.MODE_REFERENCE, shadow_res=sh, lib=oracle_lib)
        for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_SHADOW, "shadow")):
            c2.upload(slot, fi2[key])
        c2.upload(A.SLOT_MATERIAL, dielectric)
        c2.lighting_deferred(k, A.light_list_c([]), A.light_list_c([sun]))
        img = c2.readback(A.SLOT_LIGHTING).astype(np.float32)
        on_floor = fi2["depth"] < 1.0
        means.append(float(img[on_floor][:, :3].mean()))
        c2.close()
    assert means[0] > 0.05 and means[1] == 0.0, means
