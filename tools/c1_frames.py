#!/usr/bin/env python
"""BASELINE configs[0] in reference-faithful mode (Sponza 128^3, 1280x720): a few whole frames of voxelize + march + GTAO + blur +
deferred lighting + composite, with per-stage CUDA-event times.  Run it plain for the stage table, or under
`ncu --set full -k regex:'k_voxelize_r|k_trace_r|k_gtao|k_blur|k_lighting|k_composite'` for the kernel evidence."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from final184_b200 import api as A, scene as S          # noqa: E402
from final184_b200.fixture import frame_inputs           # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    N, W, H, SH = 128, 1280, 720, 2048
    sc = S.get_scene(prefer_sponza=True, seed=1)
    cams = {n: S.fixture_constants(n) for n in ("main", "shadow", "voxel")}
    fi = frame_inputs(sc, cams["main"], cams["shadow"], W, H, SH, 0)
    c = A.VoxelGI(N, W, H, A.MODE_REFERENCE, shadow_res=SH)
    c.upload_scene(sc)
    for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_SHADOW, "shadow"), (A.SLOT_MATERIAL, "material"),
                      (A.SLOT_ALBEDO, "albedo")):
        c.upload(slot, fi[key])

    def frame(f):
        k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, f, f == 0)
        if f:
            c.copy_taa_to_history(); c.copy_indirect_to_history()
        c.voxelize(cams["voxel"]); c.gtao(cams["main"]); c.trace_indirect(k); c.blur_indirect(k); c.lighting_deferred(k); c.composite(k)

    for f in range(3):
        frame(f)
    c.sync()
    c.stage_time_reset(True)
    for f in range(3, 3 + frames):
        frame(f)
    c.sync()
    st = {}
    for s in range(A.STAGE_COUNT):
        tot, runs = c.stage_total_ms(s)
        if runs:
            st[A.STAGE_NAMES[s]] = round(tot / frames, 4)
    print(json.dumps({"workload": "C1 reference-faithful mode, Sponza 128^3, 1280x720", "frames": frames, "ms_per_frame": round(sum(st.values()), 4),
                      "stages_ms": st, "fragments": c.counter(A.COUNTER_FRAGMENTS), "march_steps": c.counter(A.COUNTER_MARCH_STEPS)}))


if __name__ == "__main__":
    main()
