"""Tiny hand-built scenes for the CPU (oracle) tests."""
import numpy as np

from final184_b200 import scene as S


def quad_scene(quads, tex=None, name="hand"):
    """quads: list of (p0, e1, e2, normal) in WORLD units (identity model matrix).  One white 4x4 texture."""
    pos, nrm, uv, idx = [], [], [], []
    for p0, e1, e2, n in quads:
        b = len(pos)
        p0, e1, e2 = (np.asarray(v, np.float64) for v in (p0, e1, e2))
        for s, t in ((0, 0), (1, 0), (1, 1), (0, 1)):
            pos.append(p0 + s * e1 + t * e2); nrm.append(n); uv.append((s, t))
        idx += [(b, b + 1, b + 2), (b, b + 2, b + 3)]
    tex = tex if tex is not None else np.full((4, 4, 4), 255, np.uint8)
    T = len(idx)
    return S.Scene(pos=np.asarray(pos, np.float32), nrm=np.asarray(nrm, np.float32), uv=np.asarray(uv, np.float32),
                   idx=np.asarray(idx, np.uint32), tri_mat=np.zeros(T, np.uint16), tri_model=np.zeros(T, np.uint16),
                   model_mats=np.eye(4, dtype=np.float32)[None], mat_tex=np.zeros(1, np.int32), mat_factor=np.ones((1, 4), np.float32),
                   textures=[tex], name=name)


def tri_scene(tris, normal=(0, 1, 0), name="tris"):
    """tris: (T,3,3) world-space vertices."""
    tris = np.asarray(tris, np.float32)
    T = len(tris)
    return S.Scene(pos=tris.reshape(-1, 3), nrm=np.tile(np.asarray(normal, np.float32), (3 * T, 1)), uv=np.zeros((3 * T, 2), np.float32),
                   idx=np.arange(3 * T, dtype=np.uint32).reshape(T, 3), tri_mat=np.zeros(T, np.uint16), tri_model=np.zeros(T, np.uint16),
                   model_mats=np.eye(4, dtype=np.float32)[None], mat_tex=np.zeros(1, np.int32), mat_factor=np.ones((1, 4), np.float32),
                   textures=[np.full((4, 4, 4), 255, np.uint8)], name=name)
