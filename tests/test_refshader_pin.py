"""The pin of the oracle: the reference's OWN shader text, compiled here, against the restatement — bit for bit.

`make -C oracle ref` wraps Shader/Lighting/indirect.frag, Shader/GTAO/gtao.frag, Shader/GTAO/blur.frag,
Shader/Lighting/blurX.frag / blurY.frag (+ bilateralBlur.inc, math.inc, EngineCommon.h), the two consumers
Shader/Lighting/aggregateLights.frag and Shader/GTAO/color.frag, and the VoxelGS / BasicMaterial / VoxelPS strings of
Pipelang/Internal/main.lua — read where they lie under /root/reference — into C++ over
oracle/glsl_shim.h (purely lexical rewrite, oracle/make_ref_shaders.py) and builds oracle/_ref/libf184_refshaders.so.
tools/gen_refshader_golden.py ran it on three scenes (procedural atrium; Sponza as the reference loads it = BASELINE C1's
128^3 volume; an open floor with a wall, 22 % sky) for two frames each and committed the outputs as tests/golden/refshader_<case>.npz.

  * everywhere:            the oracle's restatement (oracle_mode_r.cpp) must reproduce the golden outputs exactly —
                           voxel volume, fragment count, lighting_indirect (frame 0 and the temporal frame 1),
                           gtao_visibility, gtao_blur, indirect_blurX, indirect_blurY, lighting_deferred, and both
                           targets of gtao_color (composite + TAA, second frame through the TAA history)
  * where the .so exists:  the live shader text must reproduce the goldens too (the fixture is not stale)
  * on the GPU:            tests/test_gpu_mode_r.py holds the CUDA path to the same goldens

What stays defined by the shim rather than by the reference (a GLSL compiler and a Vulkan driver would decide it): the
items of SURVEY.md §8(c) — transcendental polynomials, NaN rules, rasteriser, texture filtering — see glsl_shim.h."""
import pytest

import refshader as R


@pytest.fixture(scope="module", params=R.CASES)
def case(request):
    if not R.case_available(request.param):
        pytest.skip(f"{request.param}: scene not staged")
    sc, cams, fis = R.case_inputs(request.param)
    golden = R.unpack_golden(request.param)
    assert R.input_digest(sc, fis) == golden["input_sha256"], "the inputs differ from those the golden file was made from"
    return request.param, sc, cams, fis, golden


def test_restatement_matches_reference_shader_text(oracle_lib, case):
    name, sc, cams, fis, golden = case
    got = R.run_library(oracle_lib, sc, cams, fis)
    assert int(got["fragments"]) == int(golden["fragments"]) > 10000
    assert R.compare(got, golden) == []
    # the case is not degenerate: most pixels receive light, and the temporal frame differs from the first
    assert (golden["indirect0"][..., :3] != 0).mean() > 0.5 and (golden["indirect0"] != golden["indirect1"]).any()
    assert (golden["blur_x0"] != golden["blur0"]).any() and (golden["ao_raw0"] != golden["ao0"]).any()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libf184_refshaders.so not built (needs /root/reference: make -C oracle ref)")
def test_live_reference_shader_text_matches_golden(oracle_lib, case):
    name, sc, cams, fis, golden = case
    got = R.run_reference_shaders(oracle_lib, sc, cams, fis)
    assert int(got["fragments"]) == int(golden["fragments"])
    assert R.compare(got, golden) == []


def test_blur_pass_order_is_the_references(oracle_lib, case):
    """indirect_blurX steps along y (blurX.frag:5: DIR(x) = vec2(0.0, x)), indirect_blurY along x: a row-constant image is
    unchanged by the first pass only if that pass is vertical.  (The restatement had the order swapped until the shader
    text was run.)"""
    import numpy as np
    from final184_b200 import api as A
    name, sc, cams, fis, golden = case
    W, H = R.W, R.H
    c = A.VoxelGI(grid_n=32, width=W, height=H, mode=A.MODE_REFERENCE, shadow_res=64, lib=oracle_lib)
    img = np.zeros((H, W, 4), np.float16)
    img[..., :3] = (np.arange(W, dtype=np.float32) % 7)[None, :, None]          # varies along x only
    c.upload(A.SLOT_INDIRECT_OUT, img.view(np.uint16))
    c.upload(A.SLOT_DEPTH, np.full((H, W), 0.5, np.float32))
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    c.blur_indirect(k)
    bx = np.ascontiguousarray(c.readback(A.SLOT_INDIRECT_BLUR_X)).view(np.float16).reshape(H, W, 4)
    by = np.ascontiguousarray(c.readback(A.SLOT_INDIRECT_FINAL)).view(np.float16).reshape(H, W, 4)
    assert np.array_equal(bx[..., :3], img[..., :3]) and not np.array_equal(by[..., :3], img[..., :3])
    c.close()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libf184_refshaders.so not built (needs /root/reference: make -C oracle ref)")
def test_restatement_matches_live_shader_text_at_another_size(oracle_lib):
    """No golden file in between: the live shader text against the restatement at a size the fixtures do not hold (512 x 256,
    16 x the pixels), so agreement is not an artefact of the committed cases."""
    sc, cams, fis = R.case_inputs("atrium", size=(512, 256))
    ref = R.run_reference_shaders(oracle_lib, sc, cams, fis)
    got = R.run_library(oracle_lib, sc, cams, fis)
    assert R.compare(got, ref) == []
