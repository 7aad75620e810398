"""integration/vulkan/: the Vulkan side of the boundary cannot be built or run here — no Vulkan SDK, loader or
driver in the image — but it is code, and it must at least be well-formed C++ against the Vulkan API as published and against
include/f184.h.  Compiled against tests/cpp/vk_stub (declarations only) and checked for the entry points INTEGRATION.md names."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VK = os.path.join(REPO, "integration", "vulkan")


def test_vulkan_interop_compiles_against_the_api_declarations(tmp_path):
    obj = tmp_path / "interop.o"
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-c", "-fPIC", "-I" + os.path.join(REPO, "tests", "cpp", "vk_stub"), "-I" + os.path.join(REPO, "include"),
                        os.path.join(VK, "f184_vk_interop.cpp"), "-o", str(obj)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    syms = subprocess.run(["nm", "-g", "--defined-only", str(obj)], capture_output=True, text=True).stdout
    for name in ("f184vk_create_image", "f184vk_create_depth_copy", "f184vk_create_semaphore", "f184vk_import", "f184vk_import_semaphores", "f184vk_cmd_copy_depth",
                 "f184vk_cmd_release", "f184vk_cmd_acquire", "f184vk_destroy", "kF184VkDeviceExtensions"):
        assert name in syms, name
    undefined = subprocess.run(["nm", "-g", "--undefined-only", str(obj)], capture_output=True, text=True).stdout
    assert "f184_import_external_memory_fd" in undefined and "f184_import_semaphores_fd" in undefined       # it drives the C-ABI, nothing else of ours


def test_the_patch_uses_only_entry_points_that_exist():
    """every f184_* / f184vk_* call the reference-side patch makes is declared in include/f184.h or f184_vk_interop.h"""
    import re
    patch = open(os.path.join(VK, "final184_rhi.patch")).read()
    declared = open(os.path.join(REPO, "include", "f184.h")).read() + open(os.path.join(VK, "f184_vk_interop.h")).read()
    called = set(re.findall(r"\b(f184(?:vk)?_[a-z0-9_]+)\s*\(", patch))
    assert len(called) > 15
    for name in called:
        assert re.search(r"\b" + name + r"\s*\(", declared), name
    for slot in set(re.findall(r"\bF184_SLOT_[A-Z_]+", patch)):
        assert slot in declared, slot
