"""The GLSL-over-C++ machinery that pins the oracle (oracle/make_ref_shaders.py + oracle/glsl_shim.h) on a SYNTHETIC shader —
nothing from /root/reference: the lexical wrapper must not change arithmetic (fp32 literals, left-to-right swizzles, uniform
blocks, out/inout parameters, array constructors, the pinned operand order around inout calls) and the shim must follow the
conventions it documents (NaN-aware min/max, saturating uint(), fp32-weight bilinear, out-of-range fetches)."""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("make_ref_shaders", os.path.join(REPO, "oracle", "make_ref_shaders.py"))
M = importlib.util.module_from_spec(spec)
spec.loader.exec_module(M)

GLSL = """
#version 450
layout(location = 0) in vec2 inUV;
layout(location = 0) out vec4 outColor;
layout(set = 1, binding = 0) uniform sampler s;
layout(set = 1, binding = 1) uniform texture2D t_img;
layout(set = 1, binding = 2) uniform Params {
    mat4 M;
    vec2 scale;
    uint count;
};
layout(set = 1, binding = 3) uniform Light {
    vec3 luminance;
    vec3 position;
} sun;
const int TAPS = 3;
const vec2 offs[3] = vec2 [] (
    vec2(-1.0, 0.5),
    vec2(0.25, 0.0),
    vec2(1.0, -0.5)
);
struct Acc { vec3 sum; bool hit; };
float third(float x) { return x * 0.1 + 1.0 / 3.0; }
void split(in vec3 v, out float a, out vec2 bc) { a = v.x; bc = v.yz; }
vec3 bump(vec3 v, inout Acc acc) { acc.sum = acc.sum * 2.0 + v; acc.hit = true; return acc.sum.zyx; }
void main() {
    vec4 p = M * vec4(inUV * scale, 0.5, 1.0);
    p.xy = p.xy * 0.5 + 0.5;
    vec3 c = texture(sampler2D(t_img, s), p.xy).rgb;
    for (int i = 0; i < TAPS; i++) c += texelFetch(sampler2D(t_img, s), ivec2(inUV * textureSize(t_img, 0)) + ivec2(offs[i] * 2.0), 0).rgb;
    float a; vec2 bc;
    split(c, a, bc);
    Acc acc;
    acc.sum = vec3(1.0, 2.0, 3.0); acc.hit = false;
    vec3 total = vec3(0.0);
    total += acc.sum * bump(sun.luminance, acc);
    float m = max(a, sqrt(-1.0)) + min(bc.x, bc.y);
    uint q = uint(-3.5) + uint(2.9) + count;
    outColor = vec4(total * third(m), float(q) + float(int(sqrt(-1.0))));
}
"""

DRIVER = r"""
#include "glsl_shim.h"
namespace glsl { struct synth { vec4 gl_FragCoord;
#include "synth.inc"
}; }
using namespace glsl;
extern "C" void run(const float* mat, const float* img, int w, int h, float u, float v, float* out)
{
    typedef synth S;
    S::s.wrap = 0; S::t_img.data = img; S::t_img.w = w; S::t_img.h = h; S::t_img.format = TEX_R32F;
    S::M = mat4(mat); S::scale = vec2(0.75f, 1.25f); S::count = 7u;
    S::sun.luminance = vec3(0.3f, 0.6f, 0.9f); S::sun.position = vec3(0.0f);
    S sh; sh.inUV = vec2(u, v); sh.main();
    out[0] = sh.outColor.x; out[1] = sh.outColor.y; out[2] = sh.outColor.z; out[3] = sh.outColor.w;
}
"""


def test_lexical_rewrites():
    t = M.declarations(M.lexical(GLSL, ["bump"]))
    assert "layout" not in t and "#version" not in t
    assert "1.0f / 3.0f" in t and "0.1f" in t and "0x" not in t
    assert "inline static mat4 M;" in t and "struct Light_t" in t and "inline static Light_t sun;" in t
    assert "static constexpr int TAPS = 3;" in t and "offs[3] = {" in t and "vec2 [] (" not in t
    assert "p.set_xy(p.xy() * 0.5f + 0.5f);" in t
    assert "float& a, vec2& bc" in t and "Acc& acc" in t and "(vec3 v," in t
    assert "{ auto l_ = acc.sum; total += l_ * bump(sun.luminance, acc); }" in t
    assert "to_uint(-3.5f)" in t and "to_int(sqrt(-1.0f))" in t
    for lit, want in (("1.", "1.f"), (".5", ".5f"), ("2e3", "2e3f"), ("3.25f", "3.25f"), ("0x1F", "0x1F"), ("vec2", "vec2"), ("7", "7")):
        assert M.lexical(f"x = {lit};").strip() == f"x = {want};"
    with pytest.raises(SystemExit):
        M.lexical("v.zw *= 2.0;")                    # an unhandled swizzle assignment must stop the build, not mis-compile


def test_synthetic_shader_matches_numpy(tmp_path):
    (tmp_path / "synth.inc").write_text(M.declarations(M.lexical(GLSL, ["bump"])))
    (tmp_path / "driver.cpp").write_text(DRIVER)
    so = tmp_path / "synth.so"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w", "-shared", "-I", os.path.join(REPO, "oracle"),
                        "-I", str(tmp_path), str(tmp_path / "driver.cpp"), "-o", str(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    dll = C.CDLL(str(so))
    f32 = np.float32
    rng = np.random.default_rng(1)
    w, h = 8, 4
    img = rng.random((h, w)).astype(f32)
    mat = np.array([[0.5, 0, 0, 0.125], [0, 0.25, 0, 0.5], [0, 0, 1, 0], [0, 0, 0, 1]], f32)      # row-major; uploaded column-major
    out = np.zeros(4, f32)
    u, v = f32(0.5625), f32(0.375)
    dll.run(C.c_void_p(np.ascontiguousarray(mat.T).ctypes.data), C.c_void_p(img.ctypes.data), w, h, C.c_float(u), C.c_float(v), C.c_void_p(out.ctypes.data))

    # the same arithmetic in numpy float32, operation by operation
    x, y = f32(u * f32(0.75)), f32(v * f32(1.25))
    px = f32(f32(f32(f32(mat[0, 0] * x) + f32(mat[0, 1] * y)) + f32(mat[0, 2] * f32(0.5))) + f32(mat[0, 3] * f32(1)))
    py = f32(f32(f32(f32(mat[1, 0] * x) + f32(mat[1, 1] * y)) + f32(mat[1, 2] * f32(0.5))) + f32(mat[1, 3] * f32(1)))
    px, py = f32(f32(px * f32(0.5)) + f32(0.5)), f32(f32(py * f32(0.5)) + f32(0.5))

    def texel(i, j, oob_zero=False):
        if oob_zero and not (0 <= i < w and 0 <= j < h):
            return f32(0)
        return img[min(max(j, 0), h - 1), min(max(i, 0), w - 1)]
    fx, fy = f32(f32(px * f32(w)) - f32(0.5)), f32(f32(py * f32(h)) - f32(0.5))
    x0, y0 = int(np.floor(fx)), int(np.floor(fy))
    wx, wy = f32(fx - f32(x0)), f32(fy - f32(y0))
    top = f32(f32(texel(x0, y0) * f32(f32(1) - wx)) + f32(texel(x0 + 1, y0) * wx))
    bot = f32(f32(texel(x0, y0 + 1) * f32(f32(1) - wx)) + f32(texel(x0 + 1, y0 + 1) * wx))
    c = np.array([f32(f32(top * f32(f32(1) - wy)) + f32(bot * wy)), f32(0), f32(0)], f32)         # R32F: (r, 0, 0, 1)
    base = (int(f32(u * f32(w))), int(f32(v * f32(h))))
    for ox, oy in ((-1.0, 0.5), (0.25, 0.0), (1.0, -0.5)):
        c[0] = f32(c[0] + texel(base[0] + int(f32(ox) * f32(2)), base[1] + int(f32(oy) * f32(2)), oob_zero=True))
    a, bc = c[0], c[1:]
    s0 = np.array([1, 2, 3], f32)                        # acc.sum before the call: the pinned left operand
    s1 = (s0 * f32(2) + np.array([0.3, 0.6, 0.9], f32)).astype(f32)
    total = (s0 * s1[::-1]).astype(f32)
    m = f32(a + min(bc[0], bc[1]))                       # max(a, NaN) = a
    third = f32(f32(m * f32(0.1)) + f32(f32(1) / f32(3)))
    want = np.array([*(total * third).astype(f32), f32(0 + 2 + 7) + f32(0)], f32)       # uint(-3.5) = 0, uint(2.9) = 2, int(NaN) = 0
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), (out, want)
