#!/usr/bin/env python
"""Where does the fast GTAO differ from the pinned one?  (GPU; prints error statistics per 4x4 interleave cell and the worst pixels.)"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from final184_b200 import api as A, scene as S
from final184_b200.fixture import frame_inputs

sc = S.procedural_scene(seed=1)
cams = {n: S.fixture_constants(n) for n in ("main", "shadow", "voxel")}
w, h = 320, 184
fi = frame_inputs(sc, cams["main"], cams["shadow"], w, h, 512, 0, cache=False)
outs = []
for flags in (0, A.FLAG_EXACT_SECONDARY):
    c = A.VoxelGI(64, w, h, A.MODE_NORTHSTAR, shadow_res=512, flags=flags)
    for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals")):
        c.upload(slot, fi[key])
    c.gtao(cams["main"])
    outs.append(c.readback(A.SLOT_AO_RAW).astype(np.float32)[..., 0])
    c.close()
f, e = outs
print("nan fast", np.isnan(f).sum(), "nan exact", np.isnan(e).sum(), "mean", f[np.isfinite(f)].mean(), e[np.isfinite(e)].mean())
d = np.nan_to_num(f) - np.nan_to_num(e)
print("rel l2", np.linalg.norm(d) / np.linalg.norm(np.nan_to_num(e)), "max abs", np.abs(d).max(), "frac |d|>1e-2", (np.abs(d) > 1e-2).mean())
yy, xx = np.mgrid[0:h, 0:w]
for cy in range(4):
    print(" ".join(f"{np.sqrt((d[(yy & 3) == cy][:, None][((xx & 3) == cx)[(yy & 3) == cy].reshape(-1)] ** 2).mean()):.4f}" if False else f"{np.sqrt((d[((yy & 3) == cy) & ((xx & 3) == cx)] ** 2).mean()):.4f}" for cx in range(4)))
idx = np.argsort(-np.abs(d).ravel())[:12]
for i in idx:
    y, x = divmod(int(i), w)
    print("pixel", x, y, "fast", f[y, x], "exact", e[y, x], "depth", fi["depth"][y, x])
