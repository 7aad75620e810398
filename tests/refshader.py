"""Driver for oracle/_ref/libf184_refshaders.so — the reference's OWN shader text (indirect.frag, gtao.frag, GTAO/blur.frag,
blurX/blurY.frag + bilateralBlur.inc + math.inc, and VoxelGS / BasicMaterial / VoxelPS from Pipelang/Internal/main.lua)
compiled by g++ over oracle/glsl_shim.h.  Test infrastructure: used by tests/test_refshader_pin.py and by
tools/gen_refshader_golden.py, which writes tests/golden/refshader_golden.npz for the machines without /root/reference.

The shaders hard-code volume 128, shadow map 2048 and 60 steps of 0.2, so the pinned case runs at exactly that; the
screen is a power of two so that texel-centre uv's are exact in fp32 (the restatement short-cuts texel-centre fetches)."""
import ctypes as C
import hashlib
import os

import numpy as np

from final184_b200 import api as A
from final184_b200 import scene as S
from final184_b200.fixture import frame_inputs

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(REPO, "oracle", "_ref", "libf184_refshaders.so")
N, SH, W, H = 128, 2048, 128, 64
FRAMES = 2
# procedural atrium (always available); Sponza as the reference loads it (= BASELINE C1's scene); an open floor with one wall, so
# that half the screen is sky (the atmosphere branch of color.frag) and the sun reaches the ground
CASES = ("atrium", "sponza", "open")


def golden_path(case):
    return os.path.join(REPO, "tests", "golden", f"refshader_{case}.npz")


def case_available(case):
    return case != "sponza" or S.sponza_available()


def available():
    return os.path.exists(SO)


def load():
    dll = C.CDLL(SO)
    vp = C.c_void_p
    dll.refsh_indirect.argtypes = [C.POINTER(A.TraceConstantsC), vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]
    dll.refsh_gtao.argtypes = [C.POINTER(A.ViewConstantsC), vp, vp, C.c_int, C.c_int, vp]
    dll.refsh_gtao_blur.argtypes = [vp, C.c_int, C.c_int, vp]
    dll.refsh_blur.argtypes = [C.c_int, C.POINTER(A.EngineMiscsC), vp, vp, C.c_int, C.c_int, vp]
    dll.refsh_lighting_deferred.argtypes = [C.POINTER(A.ViewConstantsC), C.POINTER(A.ExtendedMatricesC), C.POINTER(A.LightListC),
                                            C.POINTER(A.LightListC), vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]
    dll.refsh_composite.argtypes = [C.POINTER(A.TraceConstantsC), vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
    return dll


def light_lists(k):
    """The fixture's lights (App/MainBehaviour.cpp:34-45): one point light of luminance Color(0.1, 1.0, 0.6) * 1.0 at (0, 2, 0) and
    the sun as the one directional light (the same PerLightConstants lighting_indirect binds as `Sun`)."""
    return A.light_list_c([((0.1, 1.0, 0.6), (0.0, 2.0, 0.0))]), A.light_list_c([(tuple(k.sun.luminance), tuple(k.sun.position))])


def case_inputs(case="atrium", size=None):
    """A pinned case: the scene under the reference's fixture cameras (App/MainBehaviour.cpp:19-76)."""
    if case == "open":
        import cpu_helpers as Hc
        tex = (np.arange(16 * 16 * 4, dtype=np.uint32).reshape(16, 16, 4) * 37 % 256).astype(np.uint8); tex[..., 3] = 255
        sc = Hc.quad_scene([((-14, 0, -14), (28, 0, 0), (0, 0, 28), (0, 1, 0)), ((-6, 0, -9), (12, 0, 0), (0, 5, 0), (0, 0, 1))], tex=tex, name="open")
    else:
        sc = S.procedural_scene(seed=1) if case == "atrium" else S.load_sponza()
    cams = {n: S.fixture_constants(n) for n in ("main", "shadow", "voxel")}
    W, H = size or (globals()["W"], globals()["H"])          # a power of two each (see the module docstring)
    fis = [frame_inputs(sc, cams["main"], cams["shadow"], W, H, SH, f, cache=False) for f in range(FRAMES)]
    # every Sponza material is (roughness, metallic) = (1, 1), which zeroes the diffuse term of the deferred
    # lighting: give the pinned case a material image that sweeps both, so that whole BRDF is exercised
    yy, xx = np.mgrid[0:H, 0:W]
    for fi in fis:
        m = np.zeros((H, W, 4), np.uint8)
        m[..., 1] = (xx * 7 + yy * 3) % 256
        m[..., 2] = (xx * 5 + yy * 11) % 256
        fi["material"] = m
    return sc, cams, fis


def input_digest(sc, fis):
    h = hashlib.sha256()
    for a in (sc.pos, sc.nrm, sc.uv, sc.idx):
        h.update(np.ascontiguousarray(a).tobytes())
    for fi in fis:
        for k in ("depth", "normals", "shadow", "albedo", "material"):
            h.update(np.ascontiguousarray(fi[k]).tobytes())
    return h.hexdigest()


def run_reference_shaders(oracle_lib, sc, cams, fis):
    """Mode R frames through the reference's shader text.  The voxel pass needs a rasteriser between VoxelGS and VoxelPS:
    that fixed-function stage is the oracle's (f184o_debug_set_voxel_stage_hooks); everything programmable is the reference's."""
    dll = load()
    H, W = fis[0]["depth"].shape
    set_hooks = getattr(oracle_lib.dll, "f184o_debug_set_voxel_stage_hooks")
    set_hooks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    o = A.VoxelGI(grid_n=N, width=W, height=H, mode=A.MODE_REFERENCE, shadow_res=SH, lib=oracle_lib)
    o.upload_scene(sc)
    assert set_hooks(o.h, C.cast(dll.refsh_voxel_gs, C.c_void_p), C.cast(dll.refsh_voxel_ps, C.c_void_p)) == 0
    o.voxelize(cams["voxel"])
    vox = o.readback(A.SLOT_VOXELS).copy()
    frags = o.counter(A.COUNTER_FRAGMENTS)
    o.close()
    out = dict(voxels=vox, fragments=np.int64(frags))
    hist = np.zeros((H, W, 4), np.uint16)
    taa_hist = np.zeros((H, W, 4), np.uint16)
    ptr = lambda a: a.ctypes.data
    for f, fi in enumerate(fis):
        depth, normals, shadow = (np.ascontiguousarray(fi[k]) for k in ("depth", "normals", "shadow"))
        k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, f, f == 0)
        ind = np.zeros((H, W, 4), np.uint16)
        dll.refsh_indirect(C.byref(k), ptr(depth), ptr(normals), ptr(shadow), ptr(vox), ptr(hist), W, H, ptr(ind))
        raw, ao = np.zeros((H, W, 4), np.uint16), np.zeros((H, W, 4), np.uint16)
        vc = A.view_constants_c(cams["main"])
        dll.refsh_gtao(C.byref(vc), ptr(depth), ptr(normals), W, H, ptr(raw))
        dll.refsh_gtao_blur(ptr(raw), W, H, ptr(ao))
        bx, by = np.zeros((H, W, 4), np.uint16), np.zeros((H, W, 4), np.uint16)
        dll.refsh_blur(0, C.byref(k.miscs), ptr(ind), ptr(depth), W, H, ptr(bx))
        dll.refsh_blur(1, C.byref(k.miscs), ptr(bx), ptr(depth), W, H, ptr(by))
        lit = np.zeros((H, W, 4), np.uint16)
        albedo, material = (np.ascontiguousarray(fi[k_]) for k_ in ("albedo", "material"))
        pl, dl = light_lists(k)
        dll.refsh_lighting_deferred(C.byref(k.view), C.byref(k.ext), C.byref(pl), C.byref(dl), ptr(albedo), ptr(normals), ptr(depth),
                                    ptr(shadow), ptr(material), W, H, ptr(lit))
        col, taa = np.zeros((H, W, 4), np.uint16), np.zeros((H, W, 4), np.uint16)
        dll.refsh_composite(C.byref(k), ptr(albedo), ptr(ao), ptr(depth), ptr(lit), ptr(shadow), ptr(by), ptr(taa_hist), W, H, ptr(col), ptr(taa))
        out.update({f"indirect{f}": ind, f"ao_raw{f}": raw, f"ao{f}": ao, f"blur_x{f}": bx, f"blur{f}": by, f"lighting{f}": lit,
                    f"color{f}": col, f"taa{f}": taa})
        taa_hist = taa.copy()                  # CopyImage(taaImageA -> taaImageB), MegaPipeline.cpp:207-210
        hist = ind.copy()                      # CopyImage(indirectImage -> indirectTemporalImage), MegaPipeline.cpp:211-214
    return out


def run_library(lib, sc, cams, fis):
    """The same frames through a libf184-shaped library (the CUDA product or the CPU oracle), reference-faithful mode."""
    H, W = fis[0]["depth"].shape
    c = A.VoxelGI(grid_n=N, width=W, height=H, mode=A.MODE_REFERENCE, shadow_res=SH, lib=lib)
    c.upload_scene(sc)
    c.voxelize(cams["voxel"])
    out = dict(voxels=c.readback(A.SLOT_VOXELS).copy(), fragments=np.int64(c.counter(A.COUNTER_FRAGMENTS)))
    u16 = lambda a: np.ascontiguousarray(a).view(np.uint16).reshape(H, W, 4).copy()
    for f, fi in enumerate(fis):
        for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_SHADOW, "shadow"), (A.SLOT_ALBEDO, "albedo"),
                          (A.SLOT_MATERIAL, "material")):
            c.upload(slot, fi[key])
        k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, f, f == 0)
        if f > 0:
            c.copy_indirect_to_history()
            c.copy_taa_to_history()
        c.trace_indirect(k)
        c.gtao(cams["main"])
        c.blur_indirect(k)
        c.lighting_deferred(k, *light_lists(k))
        c.composite(k)
        out.update({f"lighting{f}": u16(c.readback(A.SLOT_LIGHTING)), f"color{f}": u16(c.readback(A.SLOT_COLOR_OUT)),
                    f"taa{f}": u16(c.readback(A.SLOT_TAA_OUT)),f"indirect{f}": u16(c.readback(A.SLOT_INDIRECT_OUT)), f"ao_raw{f}": u16(c.readback(A.SLOT_AO_RAW)),
                    f"ao{f}": u16(c.readback(A.SLOT_AO_OUT)), f"blur_x{f}": u16(c.readback(A.SLOT_INDIRECT_BLUR_X)),
                    f"blur{f}": u16(c.readback(A.SLOT_INDIRECT_FINAL))})
    c.close()
    return out


def pack_golden(out, digest):
    vox = out["voxels"].reshape(-1, 2)
    nz = np.flatnonzero(vox.any(axis=1)).astype(np.uint32)
    g = {k: v for k, v in out.items() if k != "voxels"}
    g.update(vox_index=nz, vox_value=vox[nz].copy(), input_sha256=np.frombuffer(bytes.fromhex(digest), np.uint8))
    return g


def unpack_golden(case):
    z = np.load(golden_path(case))
    g = {k: z[k] for k in z.files}
    vox = np.zeros((N * N * N, 2), np.uint16)
    vox[g.pop("vox_index")] = g.pop("vox_value")
    g["voxels"] = vox.reshape(N, N, N, 2)
    g["input_sha256"] = bytes(g["input_sha256"]).hex()
    return g


def compare(got, want, keys=None):
    """-> list of human-readable differences (empty = bit-identical)"""
    bad = []
    for k in (keys or [k for k in want if k not in ("input_sha256",)]):
        a, b = np.asarray(got[k]), np.asarray(want[k])
        if a.shape != b.shape:
            a = a.reshape(b.shape)
        if not np.array_equal(a, b):
            n = int((a != b).sum())
            bad.append(f"{k}: {n} of {b.size} values differ")
    return bad
