"""Shared drivers for the parity tests: the same host code runs the CUDA library and the CPU oracle."""
import numpy as np

from final184_b200 import api as A
from final184_b200 import scene as S
from final184_b200.fixture import frame_inputs


def make_pair(cuda_lib, oracle_lib, scene, **kw):
    g = A.VoxelGI(lib=cuda_lib, **kw)
    o = A.VoxelGI(lib=oracle_lib, **kw)
    g.upload_scene(scene)
    o.upload_scene(scene)
    return g, o


def upload_frame(ctx, fi):
    ctx.upload(A.SLOT_DEPTH, fi["depth"])
    ctx.upload(A.SLOT_NORMALS, fi["normals"])
    ctx.upload(A.SLOT_SHADOW, fi["shadow"])
    if "material" in fi:
        ctx.upload(A.SLOT_MATERIAL, fi["material"])


def rel_l2(a, b):
    a = np.nan_to_num(np.asarray(a, np.float64))
    b = np.nan_to_num(np.asarray(b, np.float64))
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def bits16(a):
    return np.ascontiguousarray(a).view(np.uint16)
