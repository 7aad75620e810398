// glsl_shim.h — TEST INFRASTRUCTURE.  A GLSL 4.50 vocabulary for g++, so that the reference's OWN shader text
// (Shader/Lighting/indirect.frag, Shader/GTAO/gtao.frag, Shader/GTAO/blur.frag, Shader/Lighting/bilateralBlur.inc,
// Shader/math.inc, the GLSL strings of Pipelang/Internal/main.lua) can be compiled where it lies under /root/reference
// and run on the CPU (oracle/make_ref_shaders.py writes the wrapped text into oracle/_ref/gen/, never into the repo).
// The result, oracle/_ref/libf184_refshaders.so, is "the reference itself run here" for the shader arithmetic: the
// restatement in oracle_mode_r.cpp is checked against it bit for bit (tests/test_refshader_pin.py) and its outputs are
// committed as golden fixtures (tests/golden/refshader_*.npz) for the GPU box, where /root/reference does not exist.
//
// What a GLSL compiler + Vulkan driver decide, and the text does not, is decided HERE, once, with the same conventions
// SURVEY.md §8(c) lists (they are the conventions of f184_detmath.h, shared with the CUDA kernels):
//   * +,-,*,/ and sqrt are IEEE fp32, never contracted (-ffp-contract=off); fma() is fused
//   * sin/cos/log/exp2/pow are the dm_* polynomials; min/max return the non-NaN operand; int(NaN) = 0
//   * dot() adds left to right; M*v accumulates column by column; normalize() divides by sqrt(dot)
//   * texture(): bilinear with fp32 weights, top row first; texelFetch/imageLoad out of range return 0, imageStore
//     out of range is dropped
// Nothing in here is taken from the reference; nothing under final184_b200/ includes it.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../final184_b200/csrc/f184_detmath.h"

namespace glsl {

typedef uint32_t uint;

struct vec2; struct vec3; struct vec4; struct ivec2; struct ivec3; struct uvec2; struct uvec4;

// ---------------------------------------------------------------------------------------------- integer vectors
struct ivec2
{
    int x, y;
    ivec2() : x(0), y(0) {}
    explicit ivec2(int s) : x(s), y(s) {}
    ivec2(int a, int b) : x(a), y(b) {}
    explicit ivec2(const vec2& v);
    explicit ivec2(const vec4& v);
    ivec2 xy() const { return *this; }
};
struct ivec3
{
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
    explicit ivec3(const vec3& v);
    ivec2 xy() const { return ivec2(x, y); }
};
inline bool operator==(const ivec3& a, const ivec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const ivec3& a, const ivec3& b) { return !(a == b); }
inline ivec2 operator>>(const ivec2& a, int s) { return ivec2(a.x >> s, a.y >> s); }
inline ivec2 operator+(int s, const ivec2& a) { return ivec2(s + a.x, s + a.y); }
inline ivec2 operator+(const ivec2& a, const ivec2& b) { return ivec2(a.x + b.x, a.y + b.y); }

struct uvec2
{
    union { uint x; uint r; };
    union { uint y; uint g; };
    uvec2() : x(0), y(0) {}
    uvec2(uint a, uint b) : x(a), y(b) {}
};
struct uvec4
{
    union { uint x; uint r; };
    union { uint y; uint g; };
    union { uint z; uint b; };
    union { uint w; uint a; };
    uvec4() : x(0), y(0), z(0), w(0) {}
    uvec4(uint a_, uint b_, uint c_, uint d_) : x(a_), y(b_), z(c_), w(d_) {}
    uvec2 rg() const { return uvec2(x, y); }
    uvec2 xy() const { return uvec2(x, y); }
};

// ---------------------------------------------------------------------------------------------- float vectors
struct vec2
{
    union { float x; float r; float s; };
    union { float y; float g; float t; };
    vec2() : x(0), y(0) {}
    explicit vec2(float v) : x(v), y(v) {}
    vec2(float a, float b) : x(a), y(b) {}
    vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}          // GLSL converts ivec2 -> vec2 implicitly
    vec2 xy() const { return *this; }
    vec2 st() const { return *this; }
    vec2 rg() const { return *this; }
    vec2 yx() const { return vec2(y, x); }
    vec3 xyx() const;
    void set_xy(const vec2& v) { x = v.x; y = v.y; }
    void set_st(const vec2& v) { x = v.x; y = v.y; }
    vec2& operator+=(const vec2& b) { x = x + b.x; y = y + b.y; return *this; }
    vec2& operator+=(float b) { x = x + b; y = y + b; return *this; }
    vec2& operator-=(const vec2& b) { x = x - b.x; y = y - b.y; return *this; }
    vec2& operator*=(const vec2& b) { x = x * b.x; y = y * b.y; return *this; }
    vec2& operator*=(float b) { x = x * b; y = y * b; return *this; }
    vec2& operator/=(float b) { x = x / b; y = y / b; return *this; }
    void add_xy(const vec2& v) { x = x + v.x; y = y + v.y; }
    void add_st(const vec2& v) { x = x + v.x; y = y + v.y; }
};
struct vec3
{
    union { float x; float r; float s; };
    union { float y; float g; float t; };
    union { float z; float b; float p; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float v) : x(v), y(v), z(v) {}
    vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
    vec3(const vec2& v, float c) : x(v.x), y(v.y), z(c) {}
    vec3(float a, const vec2& v) : x(a), y(v.x), z(v.y) {}
    explicit vec3(const vec4& v);
    vec2 xy() const { return vec2(x, y); }
    vec2 st() const { return vec2(x, y); }
    vec2 yz() const { return vec2(y, z); }
    vec2 zy() const { return vec2(z, y); }
    vec2 xz() const { return vec2(x, z); }
    vec3 xyz() const { return *this; }
    vec3 rgb() const { return *this; }
    vec3 yzx() const { return vec3(y, z, x); }
    vec3 zxy() const { return vec3(z, x, y); }
    vec3 zyx() const { return vec3(z, y, x); }
    vec3 rrr() const { return vec3(x, x, x); }
    vec2 zz() const { return vec2(z, z); }
    void set_xy(const vec2& v) { x = v.x; y = v.y; }
    void set_st(const vec2& v) { x = v.x; y = v.y; }
    void set_xz(const vec2& v) { x = v.x; z = v.y; }
    void set_yz(const vec2& v) { y = v.x; z = v.y; }
    vec3& operator+=(const vec3& o) { x = x + o.x; y = y + o.y; z = z + o.z; return *this; }
    vec3& operator+=(float o) { x = x + o; y = y + o; z = z + o; return *this; }
    vec3& operator-=(const vec3& o) { x = x - o.x; y = y - o.y; z = z - o.z; return *this; }
    vec3& operator*=(const vec3& o) { x = x * o.x; y = y * o.y; z = z * o.z; return *this; }
    vec3& operator*=(float o) { x = x * o; y = y * o; z = z * o; return *this; }
    vec3& operator/=(float o) { x = x / o; y = y / o; z = z / o; return *this; }
    void add_xy(const vec2& v) { x = x + v.x; y = y + v.y; }
    void add_rgb(const vec3& v) { x = x + v.x; y = y + v.y; z = z + v.z; }
    void add_xyz(const vec3& v) { x = x + v.x; y = y + v.y; z = z + v.z; }
};
struct vec4
{
    union { float x; float r; float s; };
    union { float y; float g; float t; };
    union { float z; float b; float p; };
    union { float w; float a; float q; };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float v) : x(v), y(v), z(v), w(v) {}
    vec4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
    vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    vec4(const vec2& v, float c, float d) : x(v.x), y(v.y), z(c), w(d) {}
    vec4(const vec2& u, const vec2& v) : x(u.x), y(u.y), z(v.x), w(v.y) {}
    vec2 xy() const { return vec2(x, y); }
    vec2 st() const { return vec2(x, y); }
    vec2 xw() const { return vec2(x, w); }
    vec2 yw() const { return vec2(y, w); }
    vec2 zw() const { return vec2(z, w); }
    vec2 xz() const { return vec2(x, z); }
    vec2 yz() const { return vec2(y, z); }
    vec3 xyz() const { return vec3(x, y, z); }
    vec3 rgb() const { return vec3(x, y, z); }
    vec4 wzxy() const { return vec4(w, z, x, y); }
    vec4 zywx() const { return vec4(z, y, w, x); }
    vec4 xxyz() const { return vec4(x, x, y, z); }
    vec4 yzzw() const { return vec4(y, z, z, w); }
    void set_xy(const vec2& v) { x = v.x; y = v.y; }
    void set_st(const vec2& v) { x = v.x; y = v.y; }
    void set_xz(const vec2& v) { x = v.x; z = v.y; }
    void set_yz(const vec2& v) { y = v.x; z = v.y; }
    void add_st(const vec2& v) { x = x + v.x; y = y + v.y; }
    void add_xy(const vec2& v) { x = x + v.x; y = y + v.y; }
    void add_rgb(const vec3& v);
    void add_xyz(const vec3& v);
    vec3 rrr() const;
    vec4& operator+=(const vec4& o) { x = x + o.x; y = y + o.y; z = z + o.z; w = w + o.w; return *this; }
    vec4& operator*=(float o) { x = x * o; y = y * o; z = z * o; w = w * o; return *this; }
    vec4& operator/=(float o) { x = x / o; y = y / o; z = z / o; w = w / o; return *this; }
};
inline vec3 vec2::xyx() const { return vec3(x, y, x); }
inline void vec4::add_rgb(const vec3& v) { x = x + v.x; y = y + v.y; z = z + v.z; }
inline void vec4::add_xyz(const vec3& v) { x = x + v.x; y = y + v.y; z = z + v.z; }
inline vec3 vec4::rrr() const { return vec3(x, x, x); }
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}
// float -> int conversion: truncation, NaN -> 0 (PINNED)
inline ivec2::ivec2(const vec2& v) : x(dm_f2i(v.x)), y(dm_f2i(v.y)) {}
inline ivec2::ivec2(const vec4& v) : x(dm_f2i(v.x)), y(dm_f2i(v.y)) {}
inline ivec3::ivec3(const vec3& v) : x(dm_f2i(v.x)), y(dm_f2i(v.y)), z(dm_f2i(v.z)) {}

// GLSL scalar constructors written as calls: float(x), int(x), uint(x).  float(...)/int(...) are C++ functional casts
// already; uint(float) must saturate at 0 (PINNED, item 3), so the generator rewrites `uint(` to `to_uint(`.
inline uint to_uint(float f) { return dm_f2uint(f); }
inline uint to_uint(int i) { return (uint)i; }
inline uint to_uint(uint u) { return u; }
inline int to_int(float f) { return dm_f2i(f); }
inline int to_int(int i) { return i; }
inline int to_int(uint u) { return (int)u; }

#define GLSL_VEC_OPS(V, EXPAND)                                                                                        \
    inline V operator+(const V& a, const V& b) { return EXPAND(a., +, b.); }                                           \
    inline V operator-(const V& a, const V& b) { return EXPAND(a., -, b.); }                                           \
    inline V operator*(const V& a, const V& b) { return EXPAND(a., *, b.); }                                           \
    inline V operator/(const V& a, const V& b) { return EXPAND(a., /, b.); }
#define GLSL_E2(A, OP, B) vec2(A x OP B x, A y OP B y)
#define GLSL_E3(A, OP, B) vec3(A x OP B x, A y OP B y, A z OP B z)
#define GLSL_E4(A, OP, B) vec4(A x OP B x, A y OP B y, A z OP B z, A w OP B w)
GLSL_VEC_OPS(vec2, GLSL_E2)
GLSL_VEC_OPS(vec3, GLSL_E3)
GLSL_VEC_OPS(vec4, GLSL_E4)
#define GLSL_SCALAR_OPS(V, N)                                                                                          \
    inline V operator+(const V& a, float s) { return a + V(s); }                                                      \
    inline V operator+(float s, const V& a) { return V(s) + a; }                                                      \
    inline V operator-(const V& a, float s) { return a - V(s); }                                                      \
    inline V operator-(float s, const V& a) { return V(s) - a; }                                                      \
    inline V operator*(const V& a, float s) { return a * V(s); }                                                      \
    inline V operator*(float s, const V& a) { return V(s) * a; }                                                      \
    inline V operator/(const V& a, float s) { return a / V(s); }                                                      \
    inline V operator/(float s, const V& a) { return V(s) / a; }
GLSL_SCALAR_OPS(vec2, 2)
GLSL_SCALAR_OPS(vec3, 3)
GLSL_SCALAR_OPS(vec4, 4)
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }
inline bool operator==(const vec2& a, const vec2& b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(const vec2& a, const vec2& b) { return !(a == b); }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const vec3& a, const vec3& b) { return !(a == b); }

// ---------------------------------------------------------------------------------------------- matrices (column-major)
struct mat4
{
    vec4 c[4];
    mat4() {}
    explicit mat4(const float* p) { for (int j = 0; j < 4; j++) c[j] = vec4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]); }
};
struct mat3
{
    vec3 c[3];
    mat3() {}
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    explicit mat3(const mat4& m) { for (int j = 0; j < 3; j++) c[j] = m.c[j].xyz(); }
};
// PINNED: linear combination of columns, accumulated left to right
inline vec4 operator*(const mat4& M, const vec4& v)
{
    return vec4(((M.c[0].x * v.x + M.c[1].x * v.y) + M.c[2].x * v.z) + M.c[3].x * v.w,
                ((M.c[0].y * v.x + M.c[1].y * v.y) + M.c[2].y * v.z) + M.c[3].y * v.w,
                ((M.c[0].z * v.x + M.c[1].z * v.y) + M.c[2].z * v.z) + M.c[3].z * v.w,
                ((M.c[0].w * v.x + M.c[1].w * v.y) + M.c[2].w * v.z) + M.c[3].w * v.w);
}
inline vec3 operator*(const mat3& M, const vec3& v)
{
    return vec3((M.c[0].x * v.x + M.c[1].x * v.y) + M.c[2].x * v.z, (M.c[0].y * v.x + M.c[1].y * v.y) + M.c[2].y * v.z,
                (M.c[0].z * v.x + M.c[1].z * v.y) + M.c[2].z * v.z);
}
inline mat4 operator*(const mat4& A, const mat4& B)
{
    mat4 C;
    for (int j = 0; j < 4; j++) C.c[j] = A * B.c[j];
    return C;
}
inline mat3 operator*(const mat3& A, const mat3& B)
{
    mat3 C;
    for (int j = 0; j < 3; j++) C.c[j] = A * B.c[j];
    return C;
}

// ---------------------------------------------------------------------------------------------- builtins
inline float abs(float x) { return fabsf(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float sqrt(float x) { return sqrtf(x); }
inline float inversesqrt(float x) { return 1.0f / sqrtf(x); }
inline float floor(float x) { return floorf(x); }
inline float fract(float x) { return dm_fract(x); }
inline float sin(float x) { return dm_sin(x); }
inline float cos(float x) { return dm_cos(x); }
inline float log(float x) { return dm_log(x); }
inline float log2(float x) { return dm_log2(x); }
inline float exp2(float x) { return dm_exp2(x); }
inline float exp2(int x) { return dm_exp2((float)x); }
// PINNED: exp(x) = exp2(x * log2(e)), the way GPU ISAs lower it (one multiply, then the exp2 unit)
inline float exp(float x) { return dm_exp2(x * 1.4426950408889634f); }
inline float pow(float x, float y) { return dm_pow(x, y); }
inline float min(float a, float b) { return dm_min(a, b); }
inline float max(float a, float b) { return dm_max(a, b); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline float clamp(float x, float lo, float hi) { return dm_clamp(x, lo, hi); }
inline float mix(float a, float b, float t) { return dm_mix(a, b, t); }
inline float step(float edge, float x) { return dm_step(edge, x); }
inline float smoothstep(float e0, float e1, float x) { return dm_smoothstep(e0, e1, x); }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float fma(float a, float b, float c) { return fmaf(a, b, c); }
inline int floatBitsToInt(float x) { return (int)dm_f2u(x); }
inline float intBitsToFloat(int i) { return dm_u2f((uint32_t)i); }
inline ivec2 floatBitsToInt(const vec2& v) { return ivec2(floatBitsToInt(v.x), floatBitsToInt(v.y)); }
inline vec2 intBitsToFloat(const ivec2& v) { return vec2(intBitsToFloat(v.x), intBitsToFloat(v.y)); }

#define GLSL_MAP1(F)                                                                                                   \
    inline vec2 F(const vec2& a) { return vec2(F(a.x), F(a.y)); }                                                      \
    inline vec3 F(const vec3& a) { return vec3(F(a.x), F(a.y), F(a.z)); }                                              \
    inline vec4 F(const vec4& a) { return vec4(F(a.x), F(a.y), F(a.z), F(a.w)); }
GLSL_MAP1(exp) GLSL_MAP1(abs) GLSL_MAP1(sqrt) GLSL_MAP1(floor) GLSL_MAP1(fract) GLSL_MAP1(sin) GLSL_MAP1(cos) GLSL_MAP1(sign)
#define GLSL_MAP2(F)                                                                                                   \
    inline vec2 F(const vec2& a, const vec2& b) { return vec2(F(a.x, b.x), F(a.y, b.y)); }                             \
    inline vec3 F(const vec3& a, const vec3& b) { return vec3(F(a.x, b.x), F(a.y, b.y), F(a.z, b.z)); }                \
    inline vec4 F(const vec4& a, const vec4& b) { return vec4(F(a.x, b.x), F(a.y, b.y), F(a.z, b.z), F(a.w, b.w)); }
GLSL_MAP2(min) GLSL_MAP2(max) GLSL_MAP2(pow) GLSL_MAP2(step)
#define GLSL_MAP3(F)                                                                                                   \
    inline vec2 F(const vec2& a, const vec2& b, const vec2& c) { return vec2(F(a.x, b.x, c.x), F(a.y, b.y, c.y)); }    \
    inline vec3 F(const vec3& a, const vec3& b, const vec3& c) { return vec3(F(a.x, b.x, c.x), F(a.y, b.y, c.y), F(a.z, b.z, c.z)); } \
    inline vec4 F(const vec4& a, const vec4& b, const vec4& c) { return vec4(F(a.x, b.x, c.x), F(a.y, b.y, c.y), F(a.z, b.z, c.z), F(a.w, b.w, c.w)); }
GLSL_MAP3(clamp) GLSL_MAP3(mix) GLSL_MAP3(fma)
inline vec2 mix(const vec2& a, const vec2& b, float t) { return vec2(mix(a.x, b.x, t), mix(a.y, b.y, t)); }
inline vec3 mix(const vec3& a, const vec3& b, float t) { return vec3(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)); }
inline vec4 mix(const vec4& a, const vec4& b, float t) { return vec4(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t), mix(a.w, b.w, t)); }
inline vec3 max(const vec3& a, float b) { return max(a, vec3(b)); }

// PINNED: products added left to right
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(const vec2& a) { return sqrtf(dot(a, a)); }
inline float length(const vec3& a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(const vec3& a) { float l = length(a); return vec3(a.x / l, a.y / l, a.z / l); }
inline vec2 normalize(const vec2& a) { float l = length(a); return vec2(a.x / l, a.y / l); }

// ---------------------------------------------------------------------------------------------- resources
enum tex_format { TEX_R32F = 1, TEX_RGBA16_UNORM = 2, TEX_RGBA8_UNORM = 3, TEX_RGBA16F = 4 };
struct sampler { int wrap; };        // 1 = repeat (GlobalLinearSampler), 0 = clamp to edge (GlobalLinearSamplerClamped)
struct texture2D
{
    const void* data; int w, h; int format; int levels; const void* const* level_data;
    texture2D() : data(nullptr), w(0), h(0), format(0), levels(1), level_data(nullptr) {}
};
struct sampler2D
{
    const texture2D& t; const sampler& s;
    sampler2D(const texture2D& t_, const sampler& s_) : t(t_), s(s_) {}
};
inline ivec2 textureSize(const texture2D& t, int) { return ivec2(t.w, t.h); }
inline vec4 glsl_load_texel(const texture2D& t, const void* base, int w, int x, int y)
{
    size_t o = (size_t)y * w + x;
    switch (t.format)
    {
    case TEX_R32F: return vec4(((const float*)base)[o], 0.0f, 0.0f, 1.0f);
    case TEX_RGBA16_UNORM: { const uint16_t* p = (const uint16_t*)base + 4 * o;
        return vec4((float)p[0] / 65535.0f, (float)p[1] / 65535.0f, (float)p[2] / 65535.0f, (float)p[3] / 65535.0f); }
    case TEX_RGBA8_UNORM: { const uint8_t* p = (const uint8_t*)base + 4 * o;
        return vec4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f); }
    default: { const uint16_t* p = (const uint16_t*)base + 4 * o;
        return vec4(dm_f16_to_f32(p[0]), dm_f16_to_f32(p[1]), dm_f16_to_f32(p[2]), dm_f16_to_f32(p[3])); }
    }
}
inline int glsl_addr(int i, int n, int wrap)
{
    if (wrap) { int m = i % n; return m < 0 ? m + n : m; }
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}
inline vec4 texelFetch(const sampler2D& st, const ivec2& p, int)
{
    if (p.x < 0 || p.y < 0 || p.x >= st.t.w || p.y >= st.t.h) return vec4(0.0f);      // PINNED: out of range -> 0
    return glsl_load_texel(st.t, st.t.data, st.t.w, p.x, p.y);
}
// bilinear, fp32 weights, top row first (PINNED, item 5)
inline vec4 glsl_bilinear(const texture2D& t, const void* base, int w, int h, int wrap, float u, float v)
{
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int xi = dm_f2i(x0f), yi = dm_f2i(y0f);
    int xa = glsl_addr(xi, w, wrap), xb = glsl_addr(xi + 1, w, wrap), ya = glsl_addr(yi, h, wrap), yb = glsl_addr(yi + 1, h, wrap);
    vec4 a = glsl_load_texel(t, base, w, xa, ya), b = glsl_load_texel(t, base, w, xb, ya);
    vec4 c = glsl_load_texel(t, base, w, xa, yb), d = glsl_load_texel(t, base, w, xb, yb);
    return (a * (1.0f - fx) + b * fx) * (1.0f - fy) + (c * (1.0f - fx) + d * fx) * fy;
}
inline vec4 texture(const sampler2D& st, const vec2& uv)
{
    return glsl_bilinear(st.t, st.t.data, st.t.w, st.t.h, st.s.wrap, uv.x, uv.y);
}
// explicit-derivative form used for the voxel pass's base-colour fetch (the rasteriser supplies per-triangle derivatives)
inline vec4 textureGrad(const sampler2D& st, const vec2& uv, const vec2& ddx, const vec2& ddy)
{
    const texture2D& t = st.t;
    float ax = ddx.x * (float)t.w, ay = ddx.y * (float)t.h, bx = ddy.x * (float)t.w, by = ddy.y * (float)t.h;
    float mx = sqrtf(ax * ax + ay * ay), my = sqrtf(bx * bx + by * by);
    float rho = mx > my ? mx : my;
    float maxlod = (float)((t.levels - 1) < 4 ? (t.levels - 1) : 4);          // MaxLod 4 (MegaPipeline.cpp:39-43)
    float lod = 0.0f;
    if (rho > 1.0f) lod = dm_log2(rho);
    if (!(lod < maxlod)) lod = maxlod;
    float lf = floorf(lod);
    int l0 = (int)lf;
    float f = lod - lf;
    auto lvl = [&](int l) { int w = t.w >> l, h = t.h >> l; if (w < 1) w = 1; if (h < 1) h = 1;
                            return glsl_bilinear(t, t.level_data[l], w, h, st.s.wrap, uv.x, uv.y); };
    vec4 c0 = lvl(l0);
    if (f == 0.0f) return c0;
    vec4 c1 = lvl(l0 + 1);
    return c0 * (1.0f - f) + c1 * f;
}
// gather of component 0: (i0,j1), (i1,j1), (i1,j0), (i0,j0)
inline vec4 textureGatherOffset(const sampler2D& st, const vec2& uv, const ivec2& off)
{
    const texture2D& t = st.t;
    float x = uv.x * (float)t.w - 0.5f, y = uv.y * (float)t.h - 0.5f;
    int i0 = dm_f2i(floorf(x)) + off.x, j0 = dm_f2i(floorf(y)) + off.y;
    auto R = [&](int i, int j) { return glsl_load_texel(t, t.data, t.w, glsl_addr(i, t.w, st.s.wrap), glsl_addr(j, t.h, st.s.wrap)).x; };
    return vec4(R(i0, j0 + 1), R(i0 + 1, j0 + 1), R(i0 + 1, j0), R(i0, j0));
}

struct uimage3D
{
    uint16_t* data; int n;           // rg16ui, n^3
    uint64_t stores;
    uimage3D() : data(nullptr), n(0), stores(0) {}
};
inline ivec3 imageSize(const uimage3D& im) { return ivec3(im.n, im.n, im.n); }
inline uvec4 imageLoad(const uimage3D& im, const ivec3& p)
{
    if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= im.n || p.y >= im.n || p.z >= im.n) return uvec4();   // PINNED item 4
    size_t o = ((size_t)p.z * im.n + p.y) * im.n + p.x;
    return uvec4(im.data[2 * o], im.data[2 * o + 1], 0, 0);
}
inline void imageStore(uimage3D& im, const ivec3& p, const uvec4& v)
{
    if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= im.n || p.y >= im.n || p.z >= im.n) return;
    size_t o = ((size_t)p.z * im.n + p.y) * im.n + p.x;
    im.data[2 * o] = (uint16_t)v.x; im.data[2 * o + 1] = (uint16_t)v.y;
    im.stores++;
}

struct glsl_discard {};      // `discard` is rewritten to `throw glsl_discard()`

}  // namespace glsl
