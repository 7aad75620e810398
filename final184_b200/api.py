"""ctypes binding of the libf184 C-ABI (include/f184.h) and the host-side mirror of the reference's
frame section that drives it.

`VoxelGI` is deliberately written against a *(library, symbol prefix)* pair: the product is
(`libf184.so`, "f184_").  The test-suite points the same class at the CPU oracle's mirror API
(`libf184_oracle.so`, "f184o_") so both sides are driven by identical host code; nothing in this
package ever loads the oracle.

There is no CPU fallback: `load_library()` raises if the CUDA library has not been built, and
`f184_create` fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import scene as S

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libf184.so")

# ---- enums (include/f184.h) ------------------------------------------------------------------
MODE_REFERENCE, MODE_NORTHSTAR = 0, 1
(SLOT_DEPTH, SLOT_NORMALS, SLOT_ALBEDO, SLOT_MATERIAL, SLOT_SHADOW, SLOT_VOXELS, SLOT_INDIRECT_OUT,
 SLOT_INDIRECT_HISTORY, SLOT_AO_RAW, SLOT_AO_OUT, SLOT_INDIRECT_BLUR_X, SLOT_INDIRECT_FINAL,
 SLOT_ACCUM_COLOR, SLOT_ACCUM_NORMAL, SLOT_VOX_ALBEDO, SLOT_VOX_NORMAL, SLOT_RADIANCE, SLOT_MIPS,
 SLOT_BRICK_FLAGS, SLOT_LIGHTING, SLOT_TAA_HISTORY, SLOT_TAA_OUT, SLOT_COLOR_OUT, SLOT_COUNT) = range(24)
(STAGE_CLEAR, STAGE_VOXELIZE, STAGE_NORMALISE, STAGE_INJECT, STAGE_MIPS, STAGE_TRACE, STAGE_GTAO,
 STAGE_BLUR, STAGE_EXCHANGE, STAGE_LIGHTING, STAGE_COMPOSITE, STAGE_BARRIER, STAGE_APPLY, STAGE_NEED, STAGE_TAIL, STAGE_COUNT) = range(16)
STAGE_NAMES = ["clear", "voxelize", "normalise", "inject", "mips", "trace", "gtao", "blur", "exchange", "lighting", "composite", "barrier",
               "apply", "need", "tail"]
(COUNTER_FRAGMENTS, COUNTER_MARCH_STEPS, COUNTER_OCCUPIED, COUNTER_KERNEL_LAUNCHES, COUNTER_BRICKS, COUNTER_GATHER_BYTES) = range(6)
(IPC_ACCUM_COLOR, IPC_ACCUM_NORMAL, IPC_BRICK_FLAGS, IPC_EXPORT, IPC_COUNTERS, IPC_BRICK_LIST, IPC_SYNC, IPC_FRAG_QUEUE, IPC_FRAG_COUNTS,
 IPC_COUNT) = range(10)
FLAG_EXTERNAL_RANDS = 1
FLAG_NO_TMA = 2
FLAG_DENSE_MIPS = 4
FLAG_GATHER_LINEAR = 8
FLAG_NO_OVERLAP = 16
FLAG_SPEC_APPENDIX_B = 32
FLAG_EXACT_SECONDARY = 64

(FMT_UNDEFINED, FMT_R32_SFLOAT, FMT_R16G16B16A16_UNORM, FMT_R8G8B8A8_UNORM, FMT_R16G16B16A16_SFLOAT,
 FMT_R16G16_UINT, FMT_R32G32B32A32_SFLOAT, FMT_R8G8B8A8_SNORM, FMT_R32_UINT) = range(9)
_FMT_NP = {FMT_R32_SFLOAT: (np.float32, 1), FMT_R16G16B16A16_UNORM: (np.uint16, 4),
           FMT_R8G8B8A8_UNORM: (np.uint8, 4), FMT_R16G16B16A16_SFLOAT: (np.float16, 4),
           FMT_R16G16_UINT: (np.uint16, 2), FMT_R32G32B32A32_SFLOAT: (np.float32, 4),
           FMT_R8G8B8A8_SNORM: (np.int8, 4), FMT_R32_UINT: (np.uint32, 1)}


# ---- POD structs -------------------------------------------------------------------------------
class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("mode", C.c_uint32), ("grid_n", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32), ("shadow_res", C.c_uint32),
                ("march_steps", C.c_uint32), ("step_size", C.c_float), ("cone_max_distance", C.c_float),
                ("radiance_exposure", C.c_float), ("rank", C.c_uint32), ("nranks", C.c_uint32),
                ("flags", C.c_uint32)]


class ImageDesc(C.Structure):
    _fields_ = [("device_ptr", C.c_void_p), ("format", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("depth", C.c_uint32), ("row_pitch_bytes", C.c_uint32), ("size_bytes", C.c_uint64)]


class ViewConstantsC(C.Structure):
    _fields_ = [("CameraPos", C.c_float * 4), ("ViewMat", C.c_float * 16), ("ProjMat", C.c_float * 16),
                ("InvProj", C.c_float * 16)]


class ExtendedMatricesC(C.Structure):
    _fields_ = [("InvModelView", C.c_float * 16), ("ShadowView", C.c_float * 16), ("ShadowProj", C.c_float * 16),
                ("VoxelView", C.c_float * 16), ("VoxelProj", C.c_float * 16)]


class PrevProjC(C.Structure):
    _fields_ = [("PrevProjection", C.c_float * 16), ("PrevModelView", C.c_float * 16)]


class SunC(C.Structure):
    _fields_ = [("luminance", C.c_float * 3), ("_pad0", C.c_float), ("position", C.c_float * 3), ("_pad1", C.c_float)]


class EngineMiscsC(C.Structure):
    _fields_ = [("resolution", C.c_float * 2), ("frameCount", C.c_uint32), ("frameTime", C.c_float)]


class LightListC(C.Structure):
    """LightLists, MegaPipeline.cpp:101-105"""
    _fields_ = [("lights", SunC * 100), ("numLights", C.c_int32), ("_pad", C.c_int32 * 3)]


def light_list_c(lights):
    """lights: iterable of (luminance rgb, position-or-direction xyz)"""
    ll = LightListC()
    n = 0
    for lum, pos in lights:
        ll.lights[n] = SunC((C.c_float * 3)(*lum), 0.0, (C.c_float * 3)(*pos), 0.0)
        n += 1
    ll.numLights = n
    return ll


class TraceConstantsC(C.Structure):
    _fields_ = [("view", ViewConstantsC), ("ext", ExtendedMatricesC), ("prev", PrevProjC), ("sun", SunC),
                ("miscs", EngineMiscsC), ("reset_history", C.c_uint32), ("_pad", C.c_uint32 * 3)]


class SceneDescC(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("normals", C.c_void_p), ("uvs", C.c_void_p), ("indices", C.c_void_p),
                ("tri_material", C.c_void_p), ("tri_model", C.c_void_p), ("model_mats", C.c_void_p),
                ("n_verts", C.c_uint32), ("n_tris", C.c_uint32), ("n_models", C.c_uint32)]


assert C.sizeof(ViewConstantsC) == 208 and C.sizeof(ExtendedMatricesC) == 320 and C.sizeof(PrevProjC) == 128
assert C.sizeof(SunC) == 32 and C.sizeof(EngineMiscsC) == 16


def _m16(m):
    return (C.c_float * 16)(*S.to_glsl(m))


def view_constants_c(v: S.ViewConstants) -> ViewConstantsC:
    return ViewConstantsC((C.c_float * 4)(*np.asarray(v.camera_pos, np.float32)), _m16(v.view), _m16(v.proj), _m16(v.inv_proj))


def trace_constants_c(main: S.ViewConstants, shadow: S.ViewConstants, voxel: S.ViewConstants, width, height,
                      frame_count=0, reset_history=True, prev: S.ViewConstants | None = None,
                      sun_luminance=S.SUN_LUMINANCE) -> TraceConstantsC:
    """The five uniform blocks lighting_indirect binds (MegaPipeline.cpp:241-250, 259-266)."""
    prev = prev or main
    k = TraceConstantsC()
    k.view = view_constants_c(main)
    k.ext = ExtendedMatricesC(_m16(main.inv_view), _m16(shadow.view), _m16(shadow.proj), _m16(voxel.view), _m16(voxel.proj))
    k.prev = PrevProjC(_m16(prev.proj), _m16(prev.view))
    k.sun = SunC((C.c_float * 3)(*sun_luminance), 0.0, (C.c_float * 3)(*np.asarray(shadow.forward, np.float32)), 0.0)
    k.miscs = EngineMiscsC((C.c_float * 2)(float(width), float(height)), int(frame_count), 0.0)
    k.reset_history = 1 if reset_history else 0
    return k


# ---- library -----------------------------------------------------------------------------------
_SIGS = {
    "abi_version": (C.c_int, []),
    "create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "destroy": (None, [C.c_void_p]),
    "last_error": (C.c_char_p, [C.c_void_p]),
    "scene_upload": (C.c_int, [C.c_void_p, C.POINTER(SceneDescC)]),
    "texture_upload": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]),
    "material_set": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.c_int32, C.c_uint32]),
    "texture_readback": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]),
    "bind_image": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(ImageDesc)]),
    "image_info": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(ImageDesc)]),
    "upload_image": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "readback": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "readback_async": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "upload_image_rows": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "readback_async_rows": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "readback_wait": (C.c_int, [C.c_void_p, C.c_uint32]),
    "set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sync": (C.c_int, [C.c_void_p]),
    "frame_begin": (C.c_int, [C.c_void_p]),
    "frame_end": (C.c_int, [C.c_void_p]),
    "voxelize": (C.c_int, [C.c_void_p, C.POINTER(ViewConstantsC)]),
    "voxelize_accumulate": (C.c_int, [C.c_void_p, C.POINTER(ViewConstantsC)]),
    "normalise": (C.c_int, [C.c_void_p]),
    "static_cache_capture": (C.c_int, [C.c_void_p]),
    "static_cache_clear": (C.c_int, [C.c_void_p]),
    "inject": (C.c_int, [C.c_void_p, C.POINTER(SunC), C.POINTER(ExtendedMatricesC)]),
    "build_mips": (C.c_int, [C.c_void_p]),
    "trace_indirect": (C.c_int, [C.c_void_p, C.POINTER(TraceConstantsC)]),
    "trace_views": (C.c_int, [C.c_void_p, C.POINTER(TraceConstantsC), C.c_uint32, C.c_uint32, C.c_uint32]),
    "gtao": (C.c_int, [C.c_void_p, C.POINTER(ViewConstantsC)]),
    "blur_indirect": (C.c_int, [C.c_void_p, C.POINTER(EngineMiscsC)]),
    "copy_indirect_to_history": (C.c_int, [C.c_void_p]),
    "copy_taa_to_history": (C.c_int, [C.c_void_p]),
    "composite": (C.c_int, [C.c_void_p, C.POINTER(TraceConstantsC)]),
    "lighting_deferred": (C.c_int, [C.c_void_p, C.POINTER(ViewConstantsC), C.POINTER(ExtendedMatricesC), C.POINTER(LightListC), C.POINTER(LightListC)]),
    "bind_rands": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "set_triangle_range": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "set_triangle_chunks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "set_trace_rows": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "set_trace_tiles": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "stage_time_ms": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]),
    "counter_get": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]),
}
# product-only entry points (no oracle mirror)
_PRODUCT_ONLY = {
    "import_external_memory_fd": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(ImageDesc)]),
    "import_semaphores_fd": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ipc_export": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "ipc_import": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "peer_barrier": (C.c_int, [C.c_void_p]),
    "gather_volume": (C.c_int, [C.c_void_p]),
    "gather_volume_view": (C.c_int, [C.c_void_p, C.POINTER(TraceConstantsC)]),
    "microbench_peer": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_double)]),
    "stage_time_reset": (C.c_int, [C.c_void_p, C.c_uint32]),
    "stage_time_total": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "microbench": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_double)]),
    "debug_read_array": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_size_t]),
    "debug_detmath": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "debug_set_peer": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "debug_get_ipc_ptr": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
}
EXPORTS = sorted(list(_SIGS) + list(_PRODUCT_ONLY))


class F184Error(RuntimeError):
    pass


class Library:
    def __init__(self, path: str, prefix: str = "f184_", product: bool = True):
        if not os.path.exists(path):
            raise F184Error(f"{path} not found: build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                            "There is no CPU fallback for the CUDA path.")
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        sigs = dict(_SIGS)
        if product:
            sigs.update(_PRODUCT_ONLY)
        for name, (res, args) in sigs.items():
            fn = getattr(self.dll, prefix + name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)


_lib = None


def load_library() -> Library:
    """The CUDA library.  Raises if it is not built — never falls back to anything else."""
    global _lib
    if _lib is None:
        _lib = Library(LIB_PATH, "f184_", product=True)
    return _lib


class VoxelGI:
    """One voxel-GI context = the voxel/indirect/GTAO section of CMegaPipeline (MegaPipeline.cpp:195-284)."""

    def __init__(self, grid_n=128, width=1280, height=720, mode=MODE_REFERENCE, shadow_res=2048, device=0,
                 march_steps=60, step_size=0.2, flags=0, rank=0, nranks=1, lib: Library | None = None):
        self.lib = lib or load_library()
        self.cfg = Config(C.sizeof(Config), device, mode, grid_n, width, height, shadow_res, march_steps, step_size,
                          32.0, 0.0, rank, nranks, flags)
        self.h = C.c_void_p()
        rc = self.lib.create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            raise F184Error(f"{self.lib.prefix}create failed ({rc}): {self.lib.last_error(None).decode()}")
        self._keep = {}

    # -- helpers
    def _ck(self, rc, what):
        if rc != 0:
            raise F184Error(f"{self.lib.prefix}{what} failed ({rc}): {self.lib.last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.lib.destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene
    def upload_scene(self, sc: S.Scene):
        arrs = dict(pos=np.ascontiguousarray(sc.pos, np.float32), nrm=np.ascontiguousarray(sc.nrm, np.float32),
                    uv=np.ascontiguousarray(sc.uv, np.float32), idx=np.ascontiguousarray(sc.idx, np.uint32),
                    tm=np.ascontiguousarray(sc.tri_mat, np.uint16), tmod=np.ascontiguousarray(sc.tri_model, np.uint16),
                    mm=np.ascontiguousarray(np.stack([S.to_glsl(m) for m in sc.model_mats]), np.float32))
        d = SceneDescC(*[a.ctypes.data for a in arrs.values()], len(arrs["pos"]), len(arrs["idx"]), len(arrs["mm"]))
        self._ck(self.lib.scene_upload(self.h, C.byref(d)), "scene_upload")
        for k, t in enumerate(sc.textures):
            t = np.ascontiguousarray(t, np.uint8)
            self._ck(self.lib.texture_upload(self.h, k, t.ctypes.data, t.shape[1], t.shape[0]), "texture_upload")
        for k in range(len(sc.mat_tex)):
            f = (C.c_float * 4)(*sc.mat_factor[k])
            self._ck(self.lib.material_set(self.h, k, f, int(sc.mat_tex[k]), 1), "material_set")
        self.n_tris = sc.n_tris

    def texture_level(self, tex_id, level, w, h):
        out = np.empty((h, w, 4), np.uint8)
        self._ck(self.lib.texture_readback(self.h, tex_id, level, out.ctypes.data, out.nbytes), "texture_readback")
        return out

    # -- images
    def image_info(self, slot) -> ImageDesc:
        d = ImageDesc()
        self._ck(self.lib.image_info(self.h, slot, C.byref(d)), "image_info")
        return d

    def upload(self, slot, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        # the copy is asynchronous: the array must outlive it.  Two generations per slot (uploads double-buffer) are enough.
        self._keep[slot] = (self._keep.get(slot, ()) + (arr,))[-2:]
        self._ck(self.lib.upload_image(self.h, slot, arr.ctypes.data, arr.nbytes), "upload_image")

    def upload_ptr(self, slot, host_ptr, nbytes, rows=False):
        """rows=True: only the rows this rank traces travel (f184_upload_image_rows)"""
        fn = self.lib.upload_image_rows if rows else self.lib.upload_image
        self._ck(fn(self.h, slot, host_ptr, nbytes), "upload_image")

    def readback(self, slot) -> np.ndarray:
        d = self.image_info(slot)
        dt, nc = _FMT_NP[d.format]
        shape = [d.depth, d.height, d.width, nc] if d.depth > 1 else [d.height, d.width, nc]
        if nc == 1:
            shape = shape[:-1]
        out = np.empty(shape, dt)
        assert out.nbytes == d.size_bytes, (out.nbytes, d.size_bytes)
        self._ck(self.lib.readback(self.h, slot, out.ctypes.data, out.nbytes), "readback")
        return out

    def readback_async_ptr(self, slot, host_ptr, nbytes, rows=False):
        fn = self.lib.readback_async_rows if rows else self.lib.readback_async
        self._ck(fn(self.h, slot, host_ptr, nbytes), "readback_async")

    def readback_wait(self, age=0):
        self._ck(self.lib.readback_wait(self.h, age), "readback_wait")

    def read_array(self, direction, level, n, copy_out=False):
        """Texture-side storage (what the tracer samples): direction < 0 -> level-0 radiance array (copied out), else mip `level`+1 of
        that direction as the texture units return it at the texel centres.  copy_out: cudaMemcpy3D from the level's array instead
        (wrong for atlases beyond 4 GiB, i.e. 1024^3 — kept to show it)."""
        out = np.empty((n, n, n, 4), np.uint8)
        self._ck(self.lib.debug_read_array(self.h, direction, level | (0x40000000 if copy_out and direction >= 0 else 0), out.ctypes.data, out.nbytes), "debug_read_array")
        return out

    def bind(self, slot, device_ptr, fmt, width, height, depth=1):
        d = ImageDesc(device_ptr, fmt, width, height, depth, 0, 0)
        self._ck(self.lib.bind_image(self.h, slot, C.byref(d)), "bind_image")

    def set_stream(self, stream_ptr):
        self._ck(self.lib.set_stream(self.h, stream_ptr), "set_stream")

    def sync(self):
        self._ck(self.lib.sync(self.h), "sync")

    # -- passes
    def voxelize(self, voxel_cam: S.ViewConstants | ViewConstantsC):
        v = voxel_cam if isinstance(voxel_cam, ViewConstantsC) else view_constants_c(voxel_cam)
        self._ck(self.lib.voxelize(self.h, C.byref(v)), "voxelize")

    def voxelize_accumulate(self, voxel_cam: S.ViewConstants | ViewConstantsC):
        v = voxel_cam if isinstance(voxel_cam, ViewConstantsC) else view_constants_c(voxel_cam)
        self._ck(self.lib.voxelize_accumulate(self.h, C.byref(v)), "voxelize_accumulate")

    def normalise(self):
        self._ck(self.lib.normalise(self.h), "normalise")

    def static_cache_capture(self):
        """keep what the accumulators hold (the static triangles, just accumulated) as the static cache — include/f184.h"""
        self._ck(self.lib.static_cache_capture(self.h), "static_cache_capture")

    def static_cache_clear(self):
        self._ck(self.lib.static_cache_clear(self.h), "static_cache_clear")

    # -- one NVLink box (product library only)
    def ipc_export(self, buffer) -> bytes:
        h = (C.c_uint8 * 64)()
        self._ck(self.lib.ipc_export(self.h, buffer, h), "ipc_export")
        return bytes(h)

    def ipc_import(self, peer_rank, buffer, handle: bytes):
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        self._ck(self.lib.ipc_import(self.h, peer_rank, buffer, h), "ipc_import")

    def ipc_ptr(self, buffer) -> int:
        """device pointer of one of this context's shareable buffers (test hook: loopback ranks inside one process)"""
        p = C.c_void_p()
        self._ck(self.lib.debug_get_ipc_ptr(self.h, buffer, C.byref(p)), "debug_get_ipc_ptr")
        return p.value

    def set_peer(self, peer_rank, buffer, device_ptr):
        self._ck(self.lib.debug_set_peer(self.h, peer_rank, buffer, device_ptr), "debug_set_peer")

    def peer_barrier(self):
        self._ck(self.lib.peer_barrier(self.h), "peer_barrier")

    def gather_volume(self, k: TraceConstantsC | None = None):
        """k: the constants of the trace that follows — level 1 then travels only where its cones sample it"""
        if k is None:
            self._ck(self.lib.gather_volume(self.h), "gather_volume")
        else:
            self._ck(self.lib.gather_volume_view(self.h, C.byref(k)), "gather_volume_view")

    def inject(self, k: TraceConstantsC):
        self._ck(self.lib.inject(self.h, C.byref(k.sun), C.byref(k.ext)), "inject")

    def build_mips(self):
        self._ck(self.lib.build_mips(self.h), "build_mips")

    def trace_indirect(self, k: TraceConstantsC):
        self._ck(self.lib.trace_indirect(self.h, C.byref(k)), "trace_indirect")

    def trace_views(self, ks, view_height, first=0, count=None):
        """Probe batch: the images hold len(ks) views of `view_height` rows stacked top to bottom; view v uses ks[v]."""
        arr = (TraceConstantsC * len(ks))(*ks)
        count = len(ks) - first if count is None else count
        self._ck(self.lib.trace_views(self.h, arr, view_height, first, count), "trace_views")

    def gtao(self, view: S.ViewConstants | ViewConstantsC):
        v = view if isinstance(view, ViewConstantsC) else view_constants_c(view)
        self._ck(self.lib.gtao(self.h, C.byref(v)), "gtao")

    def blur_indirect(self, k: TraceConstantsC):
        self._ck(self.lib.blur_indirect(self.h, C.byref(k.miscs)), "blur_indirect")

    def lighting_deferred(self, k: TraceConstantsC, point: LightListC | None = None, directional: LightListC | None = None):
        """lighting_deferred pass; by default one directional light = the sun of `k` (MegaPipeline.cpp:106-125)."""
        if directional is None:
            directional = light_list_c([(tuple(k.sun.luminance), tuple(k.sun.position))])
        point = point if point is not None else LightListC()
        self._ck(self.lib.lighting_deferred(self.h, C.byref(k.view), C.byref(k.ext), C.byref(point), C.byref(directional)), "lighting_deferred")

    def composite(self, k: TraceConstantsC):
        self._ck(self.lib.composite(self.h, C.byref(k)), "composite")

    def copy_taa_to_history(self):
        self._ck(self.lib.copy_taa_to_history(self.h), "copy_taa_to_history")

    def copy_indirect_to_history(self):
        self._ck(self.lib.copy_indirect_to_history(self.h), "copy_indirect_to_history")

    def bind_rands(self, ptr, count):
        self._ck(self.lib.bind_rands(self.h, ptr, count), "bind_rands")

    def set_triangle_range(self, first, count):
        self._ck(self.lib.set_triangle_range(self.h, first, count), "set_triangle_range")

    def set_triangle_chunks(self, chunk_ids):
        """128-triangle chunks this context voxelizes (None / empty: every triangle of the range)"""
        a = np.ascontiguousarray(chunk_ids if chunk_ids is not None else [], np.uint32)
        self._ck(self.lib.set_triangle_chunks(self.h, a.ctypes.data if len(a) else None, len(a)), "set_triangle_chunks")

    def set_trace_rows(self, y0, y1):
        self._ck(self.lib.set_trace_rows(self.h, y0, y1), "set_trace_rows")

    def set_trace_tiles(self, first, stride):
        self._ck(self.lib.set_trace_tiles(self.h, first, stride), "set_trace_tiles")

    # -- measurement
    def stage_ms(self, stage) -> float:
        v = C.c_float()
        self._ck(self.lib.stage_time_ms(self.h, stage, C.byref(v)), "stage_time_ms")
        return v.value

    def stage_time_reset(self, accumulate=True):
        self._ck(self.lib.stage_time_reset(self.h, 1 if accumulate else 0), "stage_time_reset")

    def stage_total_ms(self, stage):
        """(sum of the stage's device time, runs) since stage_time_reset(True)."""
        v, n = C.c_float(), C.c_uint32()
        self._ck(self.lib.stage_time_total(self.h, stage, C.byref(v), C.byref(n)), "stage_time_total")
        return v.value, n.value

    def microbench(self, which) -> float:
        """0: trilinear RGBA8 3D fetches / s; 1: scattered 16-byte vector reductions / s."""
        v = C.c_double()
        self._ck(self.lib.microbench(self.h, which, C.byref(v)), "microbench")
        return v.value

    def microbench_peer(self, peer, mode, copy_bytes, depth, ctas, total_bytes) -> float:
        """GB/s of reading rank `peer`'s export buffer over NVLink (mode 0/1: bulk copies from 1/32 lanes per CTA, 2: per-lane loads)"""
        v = C.c_double()
        self._ck(self.lib.microbench_peer(self.h, peer, mode, copy_bytes, depth, ctas, total_bytes, C.byref(v)), "microbench_peer")
        return v.value

    def counter(self, which) -> int:
        v = C.c_uint64()
        self._ck(self.lib.counter_get(self.h, which, C.byref(v)), "counter_get")
        return v.value
