"""The C-ABI boundary: every entry point include/f184.h declares is exported by libf184.so (no compute calls here —
there is no GPU in the build container), the POD structs have the reference's byte layouts, and the CPU oracle
mirrors the same surface with an f184o_ prefix so one host driver serves both."""
import ctypes as C
import os
import re

import pytest

from final184_b200 import api as A

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "f184.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(f184_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_api_binds():
    names = declared_functions()
    assert "f184_voxelize" in names and "f184_trace_indirect" in names and "f184_build_mips" in names
    bound = {"f184_" + n for n in A.EXPORTS}
    assert bound == set(names), (sorted(bound - set(names)), sorted(set(names) - bound))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(A.LIB_PATH), "libf184.so not built: run __graft_entry__.build()"
    dll = C.CDLL(A.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(dll, n)]
    assert not missing, missing
    dll.f184_abi_version.restype = C.c_int
    assert dll.f184_abi_version() == 1


def test_create_fails_loudly_without_a_device_or_with_bad_config():
    """No CPU fallback: on a box without a GPU f184_create must fail with NO_DEVICE, never succeed."""
    import torch
    lib = A.load_library()
    cfg = A.Config(C.sizeof(A.Config), 0, A.MODE_NORTHSTAR, 100, 64, 64, 256, 60, 0.2, 32.0, 0.0, 0, 1, 0)   # 100 is not a power of two
    h = C.c_void_p()
    assert lib.create(C.byref(cfg), C.byref(h)) == -1
    assert b"power of two" in lib.last_error(None)
    if not torch.cuda.is_available():
        with pytest.raises(A.F184Error, match="no CUDA device|NO_DEVICE|-2"):
            A.VoxelGI(grid_n=32, width=8, height=8)


def test_struct_layouts_match_the_reference_uniform_blocks():
    # SceneView.h:8-14 (208 B), MegaPipeline.cpp:132-139 (320 B), MegaPipeline.h:16-20 (128 B),
    # MegaPipeline.cpp:95-99 (std140 vec3 @0, vec3 @16 = 32 B), MegaPipeline.cpp:141-146 (16 B)
    assert C.sizeof(A.ViewConstantsC) == 208
    assert C.sizeof(A.ExtendedMatricesC) == 320
    assert C.sizeof(A.PrevProjC) == 128
    assert C.sizeof(A.SunC) == 32 and A.SunC.position.offset == 16
    assert C.sizeof(A.EngineMiscsC) == 16 and A.EngineMiscsC.frameCount.offset == 8
    assert C.sizeof(A.TraceConstantsC) == 208 + 320 + 128 + 32 + 16 + 16
    assert A.ViewConstantsC.ViewMat.offset == 16 and A.ViewConstantsC.ProjMat.offset == 80 and A.ViewConstantsC.InvProj.offset == 144


def test_oracle_mirrors_the_boundary(oracle_lib):
    for name in A._SIGS:
        assert hasattr(oracle_lib.dll, "f184o_" + name), name
    assert oracle_lib.abi_version() == 1
