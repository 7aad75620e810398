"""GPU parity, reference-faithful mode: the CUDA path against the CPU oracle on identical inputs, through
the C-ABI.  Integer/byte outputs (voxel volume) must be bit-exact; with the shared deterministic math the
fp16 images are expected bit-exact too, and are asserted within the north_star tolerance (1e-2 rel. L2)
plus a mismatch budget that would expose a real divergence."""
import numpy as np
import pytest

import helpers as Hh
from final184_b200 import api as A
from final184_b200 import scene as S
from final184_b200.fixture import frame_inputs

pytestmark = pytest.mark.gpu


def _run_voxelize(cuda_lib, oracle_lib, scene, voxel, n):
    g, o = Hh.make_pair(cuda_lib, oracle_lib, scene, grid_n=n, width=64, height=36, mode=A.MODE_REFERENCE)
    g.voxelize(voxel)
    o.voxelize(voxel)
    return g, o


@pytest.mark.parametrize("n", [32, 128, 256])
def test_voxelize_bit_exact_procedural(cuda_lib, oracle_lib, proc_scene, cams, n):
    g, o = _run_voxelize(cuda_lib, oracle_lib, proc_scene, cams["voxel"], n)
    vg, vo = g.readback(A.SLOT_VOXELS), o.readback(A.SLOT_VOXELS)
    assert g.counter(A.COUNTER_FRAGMENTS) == o.counter(A.COUNTER_FRAGMENTS) > 0
    assert np.array_equal(vg[..., 0] != 0, vo[..., 0] != 0), "occupancy mask differs"
    assert np.array_equal(vg, vo), f"{np.count_nonzero((vg != vo).any(-1))} voxels differ"
    # second frame: the resolve pass must have left the key volume clear (ClearImage semantics)
    g.voxelize(cams["voxel"])
    assert np.array_equal(g.readback(A.SLOT_VOXELS), vo)


def test_voxelize_triangle_range_union(cuda_lib, oracle_lib, proc_scene, cams):
    """Sharding by triangle range: the ordered-store keys make max-merge of partial volumes exact."""
    g, o = _run_voxelize(cuda_lib, oracle_lib, proc_scene, cams["voxel"], 64)
    full = g.readback(A.SLOT_VOXELS)
    half = proc_scene.n_tris // 2
    o.set_triangle_range(0, half); o.voxelize(cams["voxel"]); a = o.readback(A.SLOT_VOXELS).copy()
    o.set_triangle_range(half, proc_scene.n_tris - half); o.voxelize(cams["voxel"]); b = o.readback(A.SLOT_VOXELS).copy()
    g.set_triangle_range(half, proc_scene.n_tris - half); g.voxelize(cams["voxel"])
    assert np.array_equal(g.readback(A.SLOT_VOXELS), b)
    merged = np.where((b != 0).any(-1, keepdims=True), b, a)      # later range wins where it wrote
    written_b_black = 0  # a later fragment that packs to (0,0) cannot be seen in b; tolerated below
    diff = (merged != full).any(-1)
    assert diff.sum() <= written_b_black + 8, diff.sum()


@pytest.mark.skipif(not S.sponza_available(), reason="Sponza pack not staged")
def test_voxelize_bit_exact_sponza(cuda_lib, oracle_lib, cams):
    sc = S.load_sponza()
    g, o = _run_voxelize(cuda_lib, oracle_lib, sc, cams["voxel"], 128)
    vg, vo = g.readback(A.SLOT_VOXELS), o.readback(A.SLOT_VOXELS)
    assert g.counter(A.COUNTER_FRAGMENTS) == o.counter(A.COUNTER_FRAGMENTS)
    assert np.array_equal(vg, vo)


def _trace_pair(cuda_lib, oracle_lib, scene, cams, n, w, h, frame_count=0, shadow_res=512):
    g, o = Hh.make_pair(cuda_lib, oracle_lib, scene, grid_n=n, width=w, height=h, mode=A.MODE_REFERENCE, shadow_res=shadow_res)
    fi = frame_inputs(scene, cams["main"], cams["shadow"], w, h, shadow_res, frame_count)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], w, h, frame_count, True)
    for c in (g, o):
        Hh.upload_frame(c, fi)
        c.voxelize(cams["voxel"])
        c.trace_indirect(k)
    return g, o, k, fi


def test_trace_parity_procedural(cuda_lib, oracle_lib, proc_scene, cams):
    g, o, k, _ = _trace_pair(cuda_lib, oracle_lib, proc_scene, cams, 128, 160, 90)
    ig, io = g.readback(A.SLOT_INDIRECT_OUT), o.readback(A.SLOT_INDIRECT_OUT)
    assert g.counter(A.COUNTER_MARCH_STEPS) == o.counter(A.COUNTER_MARCH_STEPS)
    mism = np.count_nonzero((Hh.bits16(ig) != Hh.bits16(io)).any(-1))
    assert Hh.rel_l2(ig[..., :3], io[..., :3]) <= 1e-2          # north_star tolerance
    assert mism == 0, f"{mism} of {ig.shape[0] * ig.shape[1]} pixels differ bitwise"
    # frame 1 with the previous output as history (temporal blend path)
    k1 = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], 160, 90, 1, False)
    for c in (g, o):
        c.copy_indirect_to_history()
        c.trace_indirect(k1)
    assert np.array_equal(Hh.bits16(g.readback(A.SLOT_INDIRECT_OUT)), Hh.bits16(o.readback(A.SLOT_INDIRECT_OUT)))


def test_gtao_and_blur_parity(cuda_lib, oracle_lib, proc_scene, cams):
    g, o, k, _ = _trace_pair(cuda_lib, oracle_lib, proc_scene, cams, 64, 160, 90)
    for c in (g, o):
        c.gtao(cams["main"])
        c.blur_indirect(k)
    for slot in (A.SLOT_AO_RAW, A.SLOT_AO_OUT, A.SLOT_INDIRECT_BLUR_X, A.SLOT_INDIRECT_FINAL):
        a, b = g.readback(slot), o.readback(slot)
        assert np.array_equal(Hh.bits16(a), Hh.bits16(b)), f"slot {slot}: {np.count_nonzero(Hh.bits16(a) != Hh.bits16(b))} values differ"


def test_detmath_device_equals_host(cuda_lib, oracle_lib):
    import ctypes as C
    rng = np.random.default_rng(0)
    g = A.VoxelGI(grid_n=32, width=8, height=8, lib=cuda_lib)
    cases = {0: rng.uniform(-3e5, 3e5, 200000), 1: rng.uniform(-3e5, 3e5, 200000), 2: rng.uniform(1e-6, 2.0, 200000),
             3: rng.uniform(1e-3, 4096.0, 200000), 4: rng.uniform(-140, 20, 200000), 5: rng.uniform(0, 1, 200000),
             6: rng.uniform(-70000, 70000, 200000)}
    for op, x in cases.items():
        x = x.astype(np.float32)
        y = np.full_like(x, 2.2)
        a, b = np.empty_like(x), np.empty_like(x)
        assert cuda_lib.debug_detmath(g.h, op, x.ctypes.data, y.ctypes.data, a.ctypes.data, x.size) == 0
        oracle_lib.dll.f184o_debug_detmath(C.c_uint32(op), C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_void_p(b.ctypes.data), C.c_size_t(x.size))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"op {op}: {np.count_nonzero(a.view(np.uint32) != b.view(np.uint32))} differ"
    # conversions (the device uses cvt instructions, the host the written-out definitions): every half pattern, and floats that
    # sit on every boundary — NaN, infinities, half subnormals, the 65504/65520 overflow edge, int/uint saturation, negative zero
    edge = np.array([0.0, -0.0, np.nan, -np.nan, np.inf, -np.inf, 65504.0, 65519.99, 65520.0, 65536.0, -65520.0, 5.96e-8, 2.98e-8, 2.9802322e-8,
                     2.9802326e-8, 6.1e-5, 6.0975552e-5, 1e-10, 2147483520.0, 2147483648.0, 4294967040.0, 4294967296.0, -2147483648.0, -2147483904.0,
                     -0.5, 0.99999994, 1e20, -1e20], np.float32)
    bits = rng.integers(0, 2 ** 32, 400000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    halves = np.arange(65536, dtype=np.uint32).view(np.float32)
    near = (rng.uniform(-70000, 70000, 200000).astype(np.float32))
    for op, x in ((6, np.concatenate([edge, bits])), (7, np.concatenate([edge, bits, near])), (8, np.concatenate([edge, bits, near])),
                  (9, np.concatenate([edge, bits, near, (near * 1e-6).astype(np.float32)])), (10, halves)):
        x = np.ascontiguousarray(x, np.float32)
        a, b = np.empty_like(x), np.empty_like(x)
        assert cuda_lib.debug_detmath(g.h, op, x.ctypes.data, x.ctypes.data, a.ctypes.data, x.size) == 0
        oracle_lib.dll.f184o_debug_detmath(C.c_uint32(op), C.c_void_p(x.ctypes.data), C.c_void_p(x.ctypes.data), C.c_void_p(b.ctypes.data), C.c_size_t(x.size))
        bad = np.flatnonzero(a.view(np.uint32) != b.view(np.uint32))
        assert bad.size == 0, f"conversion op {op}: {bad.size} differ, first x bits {x.view(np.uint32)[bad[0]]:#x}: device {a.view(np.uint32)[bad[0]]:#x} host {b.view(np.uint32)[bad[0]]:#x}"


@pytest.mark.parametrize("case", ["atrium", "sponza", "open"])
def test_cuda_matches_reference_shader_text_golden(cuda_lib, case):
    """The CUDA path against what the reference's OWN shader text computes (tests/golden/refshader_<case>.npz, made by
    tools/gen_refshader_golden.py from oracle/_ref/libf184_refshaders.so — see tests/test_refshader_pin.py): voxel volume,
    fragment count, lighting_indirect (first and temporal frame), GTAO + its blur, both bilateral passes — bit for bit.
    The oracle is not in this comparison at all."""
    import refshader as R
    if not R.case_available(case):
        pytest.skip(f"{case}: scene not staged")
    sc, cams_, fis = R.case_inputs(case)
    golden = R.unpack_golden(case)
    assert R.input_digest(sc, fis) == golden["input_sha256"], "the inputs differ from those the golden file was made from"
    got = R.run_library(cuda_lib, sc, cams_, fis)
    assert int(got["fragments"]) == int(golden["fragments"]) > 10000
    assert R.compare(got, golden) == []


def test_deferred_lighting_bit_exact(cuda_lib, oracle_lib, proc_scene, cams):
    """lighting_deferred (aggregateLights.frag) at a non-power-of-two size with three point lights and two directional lights:
    RGBA16F output equal bit for bit to the oracle's restatement (itself pinned to the shader text)."""
    w, h, sh = 320, 184, 1024
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], w, h, sh, 0, cache=False)
    yy, xx = np.mgrid[0:h, 0:w]
    mat = np.zeros((h, w, 4), np.uint8); mat[..., 1] = (xx * 7 + yy * 3) % 256; mat[..., 2] = (xx * 5 + yy * 11) % 256
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], w, h, 0, True)
    point = A.light_list_c([((0.1, 1.0, 0.6), (0.0, 2.0, 0.0)), ((3.0, 0.5, 0.2), (-4.0, 1.0, -2.0)), ((0.0, 0.0, 9.0), (5.0, 6.0, 1.0))])
    sun = (tuple(k.sun.luminance), tuple(k.sun.position))
    directional = A.light_list_c([sun, ((0.5, 0.5, 1.5), (0.3, -0.8, 0.52))])
    outs = []
    for lib in (cuda_lib, oracle_lib):
        c = A.VoxelGI(grid_n=32, width=w, height=h, mode=A.MODE_REFERENCE, shadow_res=sh, lib=lib)
        for slot, arr in ((A.SLOT_DEPTH, fi["depth"]), (A.SLOT_NORMALS, fi["normals"]), (A.SLOT_SHADOW, fi["shadow"]), (A.SLOT_MATERIAL, mat)):
            c.upload(slot, arr)
        c.lighting_deferred(k, point, directional)
        a = c.readback(A.SLOT_LIGHTING).copy()
        c.lighting_deferred(k, A.LightListC(), A.LightListC())
        z = c.readback(A.SLOT_LIGHTING).copy()
        outs.append((a, z))
        c.close()
    (ag, zg), (ao, zo) = outs
    assert np.array_equal(Hh.bits16(ag), Hh.bits16(ao)), f"{np.count_nonzero(Hh.bits16(ag) != Hh.bits16(ao))} values differ"
    assert ao[..., :3].astype(np.float32).mean() > 1e-3 and np.isfinite(ao.astype(np.float32)).all()
    assert np.array_equal(Hh.bits16(zg), Hh.bits16(zo)) and not zo[..., :3].any() and (zo[..., 3] == 1).all()     # no lights: black, alpha 1


def test_cuda_matches_live_reference_shader_text(cuda_lib, oracle_lib):
    """Where oracle/_ref/libf184_refshaders.so travelled with the snapshot: the CUDA path against the reference's shader text run
    live at 512 x 256 — a size no golden file holds — every mode R output bit for bit."""
    import refshader as R
    if not R.available():
        pytest.skip("oracle/_ref/libf184_refshaders.so not present on this box")
    sc, cams_, fis = R.case_inputs("atrium", size=(512, 256))
    ref = R.run_reference_shaders(oracle_lib, sc, cams_, fis)
    got = R.run_library(cuda_lib, sc, cams_, fis)
    assert int(got["fragments"]) == int(ref["fragments"])
    assert R.compare(got, ref) == []
