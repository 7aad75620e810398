"""Input synthesiser for tests and the benchmark (libf184_fixture.so, CPU): renders the G-buffer and the
shadow map the Vulkan renderer would hand to the hot path (see gbuffer_raster.cpp).  It is neither the
product hot path nor the oracle; both consume its output unchanged.  Results are cached under
assets/_cache/ (git-ignored) keyed by scene, camera and resolution."""
from __future__ import annotations

import ctypes as C
import hashlib
import os

import numpy as np

from .. import scene as S
from ..api import SceneDescC, ViewConstantsC, view_constants_c

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libf184_fixture.so")
CACHE = os.path.join(S.REPO, "assets", "_cache")


class Fixture:
    def __init__(self, sc: S.Scene):
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built (run __graft_entry__.build())")
        self.dll = C.CDLL(LIB_PATH)
        if hasattr(self.dll, "f184fx_set_threads"):
            self.dll.f184fx_set_threads(C.c_int(os.cpu_count() or 1))
        self.h = C.c_void_p()
        self.dll.f184fx_create(C.byref(self.h))
        self.scene = sc
        a = dict(pos=np.ascontiguousarray(sc.pos, np.float32), nrm=np.ascontiguousarray(sc.nrm, np.float32),
                 uv=np.ascontiguousarray(sc.uv, np.float32), idx=np.ascontiguousarray(sc.idx, np.uint32),
                 tm=np.ascontiguousarray(sc.tri_mat, np.uint16), tmod=np.ascontiguousarray(sc.tri_model, np.uint16),
                 mm=np.ascontiguousarray(np.stack([S.to_glsl(m) for m in sc.model_mats]), np.float32))
        d = SceneDescC(*[x.ctypes.data for x in a.values()], len(a["pos"]), len(a["idx"]), len(a["mm"]))
        self.dll.f184fx_scene_upload(self.h, C.byref(d))
        for k, t in enumerate(sc.textures):
            t = np.ascontiguousarray(t, np.uint8)
            self.dll.f184fx_texture_upload(self.h, C.c_uint32(k), C.c_void_p(t.ctypes.data), C.c_uint32(t.shape[1]), C.c_uint32(t.shape[0]))
        for k in range(len(sc.mat_tex)):
            self.dll.f184fx_material_set(self.h, C.c_uint32(k), (C.c_float * 4)(*sc.mat_factor[k]), C.c_int32(int(sc.mat_tex[k])), C.c_uint32(1))

    def __del__(self):
        try:
            self.dll.f184fx_destroy(self.h)
        except Exception:
            pass

    def gbuffer(self, view: S.ViewConstants, width, height, frame_count=0):
        v = view_constants_c(view)
        depth = np.empty((height, width), np.float32)
        normals = np.empty((height, width, 4), np.uint16)
        albedo = np.empty((height, width, 4), np.uint8)
        material = np.empty((height, width, 4), np.uint8)
        self.dll.f184fx_render_gbuffer(self.h, C.byref(v), C.c_uint32(frame_count), C.c_uint32(width), C.c_uint32(height),
                                       C.c_void_p(depth.ctypes.data), C.c_void_p(normals.ctypes.data),
                                       C.c_void_p(albedo.ctypes.data), C.c_void_p(material.ctypes.data))
        return dict(depth=depth, normals=normals, albedo=albedo, material=material)

    def shadow(self, view: S.ViewConstants, res=2048):
        v = view_constants_c(view)
        depth = np.empty((res, res), np.float32)
        self.dll.f184fx_render_shadow(self.h, C.byref(v), C.c_uint32(res), C.c_void_p(depth.ctypes.data))
        return depth


def _key(sc: S.Scene, *parts):
    h = hashlib.sha1()
    h.update(sc.name.encode())
    h.update(str(sc.n_tris).encode())
    for p in parts:
        h.update(np.ascontiguousarray(p).tobytes() if isinstance(p, np.ndarray) else str(p).encode())
    return h.hexdigest()[:16]


def frame_inputs(sc: S.Scene, main: S.ViewConstants, shadow: S.ViewConstants, width, height, shadow_res=2048,
                 frame_count=0, cache=True):
    """G-buffer + shadow map for one frame, cached on disk (a 4K Sponza G-buffer takes a few seconds of CPU)."""
    key = _key(sc, main.view, main.proj, shadow.view, shadow.proj, width, height, shadow_res, frame_count)
    path = os.path.join(CACHE, f"frame_{key}.npz")
    if cache and os.path.exists(path):
        z = np.load(path)
        return {k: z[k] for k in z.files}
    fx = Fixture(sc)
    out = fx.gbuffer(main, width, height, frame_count)
    out["shadow"] = fx.shadow(shadow, shadow_res)
    if cache:
        os.makedirs(CACHE, exist_ok=True)
        np.savez(path, **out)
    return out
