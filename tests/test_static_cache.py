"""Static / dynamic split (include/f184.h "static / dynamic split") on the oracle: triangles [0, S) accumulated
once and captured, every frame voxelizes only [S, end) — volumes, mips, the traced image and the counters must equal voxelizing [0, end)
every frame, bit for bit (the accumulator sums are integer-valued, so the split cannot change them).  The CUDA path is held to the same
in tests/test_gpu_loopback.py::test_static_cache_equals_full_voxelization."""
import numpy as np
import pytest

from final184_b200 import api as A
from final184_b200.fixture import frame_inputs

N, W, H, SH = 32, 48, 32, 64
SLOTS = ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material"), (A.SLOT_SHADOW, "shadow"))


def _ctx(lib, sc, fi):
    c = A.VoxelGI(N, W, H, A.MODE_NORTHSTAR, shadow_res=SH, lib=lib)
    c.upload_scene(sc)
    for slot, key in SLOTS:
        c.upload(slot, fi[key])
    return c


def _frame(c, cam, k):
    c.voxelize(cam); c.inject(k); c.build_mips(); c.trace_indirect(k)


def test_static_cache_equals_full_voxelization_on_the_oracle(oracle_lib, proc_scene, cams):
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], W, H, SH, 0, cache=False)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    cam = cams["voxel"]
    T = proc_scene.n_tris
    S_ = (2 * T) // 3
    ref, c = _ctx(oracle_lib, proc_scene, fi), _ctx(oracle_lib, proc_scene, fi)
    c.set_triangle_range(0, S_)
    c.voxelize_accumulate(cam)
    static_frags = c.counter(A.COUNTER_FRAGMENTS)
    c.static_cache_capture()
    assert not c.readback(A.SLOT_ACCUM_COLOR).any() and static_frags > 0            # the accumulators are left clear
    for end in (T, S_ + (T - S_) // 2, S_, T):
        ref.set_triangle_range(0, end); _frame(ref, cam, k)
        c.set_triangle_range(S_, end - S_); _frame(c, cam, k)
        for slot in (A.SLOT_VOX_ALBEDO, A.SLOT_VOX_NORMAL, A.SLOT_RADIANCE, A.SLOT_MIPS):
            assert np.array_equal(c.readback(slot), ref.readback(slot)), (end, slot)
        assert np.array_equal(c.readback(A.SLOT_INDIRECT_OUT).view(np.uint16), ref.readback(A.SLOT_INDIRECT_OUT).view(np.uint16))
        for w in (A.COUNTER_FRAGMENTS, A.COUNTER_OCCUPIED, A.COUNTER_BRICKS):
            assert c.counter(w) == ref.counter(w), (end, w)
    with pytest.raises(A.F184Error, match="another voxel camera"):
        c.voxelize_accumulate(cams["main"])
    c.static_cache_clear()
    ref.set_triangle_range(0, T // 2); _frame(ref, cam, k)
    c.set_triangle_range(0, T // 2); _frame(c, cam, k)
    assert np.array_equal(c.readback(A.SLOT_RADIANCE), ref.readback(A.SLOT_RADIANCE)) and c.counter(A.COUNTER_FRAGMENTS) == ref.counter(A.COUNTER_FRAGMENTS)
    c.close(); ref.close()
