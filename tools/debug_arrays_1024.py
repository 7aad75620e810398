#!/usr/bin/env python
"""Does f184_debug_read_array return what the linear MIPS slot holds at 1024^3?  (one GPU)"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from final184_b200 import api as A, scene as S
from final184_b200.fixture import Fixture

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sp = S.load_sponza()
sc = S.tile_scene(sp, S.C4_OFFSETS) if N == 1024 else sp
cam = S.fixture_constants("voxel_c4" if N == 1024 else "voxel")
main_, shadow_ = S.fixture_constants("main"), S.fixture_constants("shadow")
k = A.trace_constants_c(main_, shadow_, cam, 64, 36, 0, True)
g = A.VoxelGI(grid_n=N, width=64, height=36, mode=A.MODE_NORTHSTAR, flags=A.FLAG_NO_OVERLAP)
g.upload_scene(sc)
g.upload(A.SLOT_SHADOW, Fixture(sc).shadow(shadow_, 2048))
g.voxelize(cam); g.inject(k); g.build_mips()
mips = g.readback(A.SLOT_MIPS).reshape(-1, 4)
off, n, lvl = 0, N // 2, 0
while n >= 16:
    for d in range(6):
        lin = mips[off + d * n ** 3: off + (d + 1) * n ** 3].reshape(n, n, n, 4)
        arr = g.read_array(d, lvl, n, copy_out=True)
        tex = g.read_array(d, lvl, n)
        print(f"level {lvl + 1} (n={n}) dir {d}: copy-out differs in {int((lin != arr).any(-1).sum())} texels, texture fetch differs in {int((lin != tex).any(-1).sum())} of {int((lin[..., 3] != 0).sum())} occupied", flush=True)
    off += 6 * n ** 3
    n //= 2; lvl += 1
