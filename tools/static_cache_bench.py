"""What the static / dynamic split buys on one GPU (include/f184.h "static / dynamic split"): Sponza 512^3 — all of it static, as in the
reference's scene — frame time and stage times with the whole scene re-voxelized every frame vs captured once.  One JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from final184_b200 import api as A, scene as S          # noqa: E402
from final184_b200.fixture import frame_inputs          # noqa: E402

N, W, H, SH, K = 512, 1920, 1080, 2048, 20
SLOTS = ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material"), (A.SLOT_SHADOW, "shadow"))


def run(c, cam, k, frames):
    for _ in range(3):
        c.voxelize(cam); c.inject(k); c.build_mips(); c.trace_indirect(k)
    c.sync()
    c.stage_time_reset(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(frames):
        c.voxelize(cam); c.inject(k); c.build_mips(); c.trace_indirect(k)
    c.sync(); t1 = time.perf_counter()
    st = {n: round(c.stage_total_ms(s)[0] / frames, 4) for n, s in (("voxelize", A.STAGE_VOXELIZE), ("normalise", A.STAGE_NORMALISE), ("inject", A.STAGE_INJECT),
                                                                   ("mips", A.STAGE_MIPS), ("trace", A.STAGE_TRACE))}
    c.stage_time_reset(False)
    return round((t1 - t0) * 1e3 / frames, 4), st


def main():
    torch.cuda.set_device(0)
    sc = S.load_sponza() if S.sponza_available() else S.procedural_scene(seed=1)
    cams = {n: S.fixture_constants(n) for n in ("main", "shadow", "voxel")}
    fi = frame_inputs(sc, cams["main"], cams["shadow"], W, H, SH, 0)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    out = {"workload": f"{'Sponza' if S.sponza_available() else 'procedural'} {N}^3, trace at {W}x{H}, {K} frames, wall clock around the pipelined loop", "triangles": int(sc.n_tris)}
    for flags, tag in ((0, "pipelined"), (A.FLAG_NO_OVERLAP, "one_stream")):
        c = A.VoxelGI(N, W, H, A.MODE_NORTHSTAR, shadow_res=SH, device=0, flags=flags)
        c.upload_scene(sc)
        for slot, key in SLOTS:
            c.upload(slot, fi[key])
        c.set_triangle_range(0, sc.n_tris)
        full_ms, full_st = run(c, cams["voxel"], k, K)
        img_full = c.readback(A.SLOT_INDIRECT_OUT).copy()
        c.voxelize_accumulate(cams["voxel"]); c.static_cache_capture()
        c.set_triangle_range(0, 0)                       # nothing moves in the reference's scene
        split_ms, split_st = run(c, cams["voxel"], k, K)
        same = bool(np.array_equal(c.readback(A.SLOT_INDIRECT_OUT).view(np.uint16), img_full.view(np.uint16)))
        out[tag] = {"revoxelize_every_frame": {"ms_per_frame": full_ms, "stages_ms": full_st}, "static_cache": {"ms_per_frame": split_ms, "stages_ms": split_st},
                    "image_bit_identical": same}
        c.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
