#!/usr/bin/env python
"""Golden vectors for BASELINE configs[3] (8 x tiled Sponza, 2.1 M triangles, 1024^3) from the CPU oracle.

The oracle needs ~48 GB of host memory and minutes of CPU at 1024^3, too much for every test run, so its answers are pinned
here once: the order-independent counters of the whole volume (fragments, occupied voxels, touched bricks) and a 128^3 sub-cube —
the densest of a fixed set of candidates — of every volume the path produces (mean albedo, mean normal, injected radiance, and
the matching region of every level of the six-direction mip chain).  tests/test_gpu_fullsize.py holds the CUDA path to them.

    python tools/gen_c4_golden.py            ->  tests/golden/c4_oracle.npz
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from final184_b200 import api as A, scene as S          # noqa: E402
from final184_b200.fixture import Fixture               # noqa: E402

N, SUB = 1024, 128


def main():
    olib = A.Library(os.path.join(REPO, "oracle", "_build", "libf184_oracle.so"), "f184o_", product=False)
    sc = S.tile_scene(S.load_sponza(), S.C4_OFFSETS)
    cams = {n: S.fixture_constants(n) for n in ("main", "shadow")}
    cams["voxel"] = S.fixture_constants("voxel_c4")
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], 64, 36, 0, True)
    shadow = Fixture(sc).shadow(cams["shadow"], 2048)
    o = A.VoxelGI(grid_n=N, width=64, height=36, mode=A.MODE_NORTHSTAR, shadow_res=2048, lib=olib)
    o.upload_scene(sc)
    o.upload(A.SLOT_SHADOW, shadow)
    t0 = time.time()
    o.voxelize(cams["voxel"])
    counters = {"fragments": o.counter(A.COUNTER_FRAGMENTS), "occupied": o.counter(A.COUNTER_OCCUPIED), "bricks": o.counter(A.COUNTER_BRICKS)}
    print("voxelize", round(time.time() - t0, 1), "s", counters, flush=True)
    alb = o.readback(A.SLOT_VOX_ALBEDO)
    occ = (alb[..., 3] != 0)
    blocks = occ.reshape(N // SUB, SUB, N // SUB, SUB, N // SUB, SUB).sum((1, 3, 5))
    bz, by, bx = np.unravel_index(int(blocks.argmax()), blocks.shape)
    z0, y0, x0 = int(bz) * SUB, int(by) * SUB, int(bx) * SUB
    print("densest sub-cube at (x, y, z) =", (x0, y0, z0), "with", int(blocks.max()), "occupied voxels", flush=True)
    cut = lambda v, n, s: np.ascontiguousarray(v[z0 * n // N:z0 * n // N + s, y0 * n // N:y0 * n // N + s, x0 * n // N:x0 * n // N + s])
    out = {"origin": np.array([x0, y0, z0]), "counters": np.array([counters["fragments"], counters["occupied"], counters["bricks"]], np.int64),
           "albedo": cut(alb, N, SUB)}
    del alb, occ
    out["normal"] = cut(o.readback(A.SLOT_VOX_NORMAL), N, SUB)
    o.inject(k)
    print("inject", round(time.time() - t0, 1), "s", flush=True)
    out["radiance"] = cut(o.readback(A.SLOT_RADIANCE), N, SUB)
    o.build_mips()
    print("mips", round(time.time() - t0, 1), "s", flush=True)
    mips = o.readback(A.SLOT_MIPS).reshape(-1, 4)
    off, n, lvl = 0, N // 2, 1
    while n >= 1:
        s = max(1, SUB * n // N)
        for d in range(6):
            v = mips[off + d * n ** 3: off + (d + 1) * n ** 3].reshape(n, n, n, 4)
            out[f"mip{lvl}_d{d}"] = cut(v, n, s)
        off += 6 * n ** 3
        n //= 2; lvl += 1
    o.close()
    path = os.path.join(REPO, "tests", "golden", "c4_oracle.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
