#!/usr/bin/env python
"""Per-stage device times (CUDA events inside libf184) of one frame in both modes.  usage: stage_timing.py N W H [iters]"""
import json
import sys

sys.path.insert(0, '.')
from final184_b200 import api as A, scene as S
from final184_b200.fixture import frame_inputs

N, W, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 6
sc = S.get_scene()
main, shadow, voxel = (S.fixture_constants(n) for n in ('main', 'shadow', 'voxel'))
fi = frame_inputs(sc, main, shadow, W, H, 2048)
for mode in (A.MODE_NORTHSTAR, A.MODE_REFERENCE):
    g = A.VoxelGI(N, W, H, mode)
    g.upload_scene(sc)
    for s, k in ((A.SLOT_DEPTH, 'depth'), (A.SLOT_NORMALS, 'normals'), (A.SLOT_SHADOW, 'shadow'), (A.SLOT_MATERIAL, 'material')):
        g.upload(s, fi[k])
    for it in range(iters):
        k = A.trace_constants_c(main, shadow, voxel, W, H, it, it == 0)
        if it == 2:
            g.stage_time_reset(True)
        if it:
            g.copy_indirect_to_history()
        g.voxelize(voxel)
        if mode == A.MODE_NORTHSTAR:
            g.inject(k); g.build_mips()
        g.trace_indirect(k); g.gtao(main); g.blur_indirect(k)
    g.sync()
    r = {}
    for s in range(A.STAGE_COUNT):
        tot, runs = g.stage_total_ms(s)
        if runs:
            r[A.STAGE_NAMES[s]] = round(tot / (iters - 2), 4)
    r['frags'] = g.counter(A.COUNTER_FRAGMENTS); r['samples'] = g.counter(A.COUNTER_MARCH_STEPS)
    print(json.dumps({'mode': 'N' if mode == A.MODE_NORTHSTAR else 'R', 'grid': N, 'W': W, 'H': H, **r}), flush=True)
    g.close()
