// oracle_common.h — TEST INFRASTRUCTURE.  Shared plumbing of the CPU oracle (libf184_oracle.so).
//
// The oracle is a CPU restatement of the reference's GPU programs for the voxel-GI hot path, one
// function per shader stage, each citing the reference file:line it follows.  It exists to CHECK the
// CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it.  Nothing under final184_b200/ links, imports or executes it.
//
// Parity status: mode R (the shipped shaders) is PINNED against the reference itself run here — the reference's own
// shader text is compiled where it lies into oracle/_ref/libf184_refshaders.so (glsl_shim.h, make_ref_shaders.py,
// ref_shaders.cpp) and oracle_mode_r.cpp reproduces its voxel volume and images bit for bit
// (tests/test_refshader_pin.py, goldens tests/golden/refshader_*.npz); camera/view constants are pinned by the
// reference Math library (oracle/_ref/ref_math_probe -> tests/golden/ref_constants.json) and texture decode by the
// reference stb_image.h (oracle/_ref/stb_decode).  NOT pinned by any reference code, because the reference leaves
// them to a GLSL compiler / Vulkan driver that is absent here: the items marked "PINNED:" in oracle_mode_r.cpp
// (rasteriser, texture unit, transcendentals, NaN rules) — and mode N as a whole (the reference has no such stages):
// for those this oracle is the definition ("parity unpinned").
//
// It mirrors the product's C-ABI (include/f184.h) one-to-one with an `f184o_` prefix so the same test
// code drives both; "device pointers" are host pointers here.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../include/f184.h"
#include "../final184_b200/csrc/f184_detmath.h"

#ifdef F184_ORACLE_LIBM
// variant used to quantify how much a different driver's sin/cos/log/pow would move the image
#define O_SIN(x) sinf(x)
#define O_COS(x) cosf(x)
#define O_LOG(x) logf(x)
#define O_LOG2(x) log2f(x)
#define O_EXP2(x) exp2f(x)
#define O_POW(x, y) powf(x, y)
#else
#define O_SIN(x) dm_sin(x)
#define O_COS(x) dm_cos(x)
#define O_LOG(x) dm_log(x)
#define O_LOG2(x) dm_log2(x)
#define O_EXP2(x) dm_exp2(x)
#define O_POW(x, y) dm_pow(x, y)
#endif

namespace orc {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline V3 normalize(V3 a) { float l = length(a); return {a.x / l, a.y / l, a.z / l}; }
inline V3 vabs(V3 a) { return {fabsf(a.x), fabsf(a.y), fabsf(a.z)}; }

// mat4 in upload order: m[c*4+r]
struct M4 { float m[16]; };
inline M4 load_m4(const float* p) { M4 r; memcpy(r.m, p, 64); return r; }
inline V4 mul(const M4& M, V4 v)
{
    V4 r;
    r.x = ((M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z) + M.m[12] * v.w;
    r.y = ((M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z) + M.m[13] * v.w;
    r.z = ((M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z) + M.m[14] * v.w;
    r.w = ((M.m[3] * v.x + M.m[7] * v.y) + M.m[11] * v.z) + M.m[15] * v.w;
    return r;
}
inline M4 matmul(const M4& A, const M4& B)
{
    M4 C;
    for (int j = 0; j < 4; j++)
    {
        V4 c = mul(A, V4{B.m[j * 4 + 0], B.m[j * 4 + 1], B.m[j * 4 + 2], B.m[j * 4 + 3]});
        C.m[j * 4 + 0] = c.x; C.m[j * 4 + 1] = c.y; C.m[j * 4 + 2] = c.z; C.m[j * 4 + 3] = c.w;
    }
    return C;
}
// mat3(M) * v
inline V3 mul3(const M4& M, V3 v)
{
    return {(M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z,
            (M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z,
            (M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z};
}

struct Texture
{
    uint32_t w = 0, h = 0;
    std::vector<std::vector<uint8_t>> levels;   // RGBA8, level l is (w>>l) x (h>>l)
};
struct Material
{
    float factor[4] = {1, 1, 1, 1};
    int32_t tex = -1;
    uint32_t use_textures = 1;
};
struct Image
{
    std::vector<uint8_t> own;
    void* ptr = nullptr;
    f184_image_desc desc{};
};

}  // namespace orc

// Programmable stages of the voxel pass as callbacks (see f184o_ctx::gs_hook).  GS: one triangle in, three vertices out.
// PS: one fragment in; returns 1 if it stored a texel, 0 if it discarded or the store was out of range.
struct f184o_ps_material
{
    float factor[4];
    uint32_t use_textures;
    uint32_t tex_w, tex_h, tex_levels;           // 0 levels = no texture bound
    const uint8_t* const* level_data;            // RGBA8, level l is (w>>l) x (h>>l)
};
typedef void (*f184o_voxel_gs_hook)(const float* view16, const float* proj16, const float* model16, const float* pos9,
                                    const float* nrm9, const float* uv6, float* out_clip12, float* out_nrm9, float* out_uv6,
                                    uint32_t* out_orientation);
typedef int (*f184o_voxel_ps_hook)(const float* fragcoord4, const float* normal3, const float* uv2, const float* duvdx2,
                                   const float* duvdy2, uint32_t orientation, const f184o_ps_material* material,
                                   uint16_t* voxels, uint32_t grid_n);

struct f184o_ctx
{
    f184_config cfg{};
    std::string err;
    // scene
    std::vector<float> pos, nrm, uv, model_mats;
    std::vector<uint32_t> idx;
    std::vector<uint16_t> tri_mat, tri_model;
    uint32_t n_verts = 0, n_tris = 0, n_models = 0;
    std::vector<orc::Texture> textures;
    std::vector<orc::Material> materials;
    orc::Image img[F184_SLOT_COUNT];
    // mode N mip chain: level l>=1, direction d: mips[l][d] is RGBA8 (N>>l)^3
    std::vector<std::vector<std::vector<uint8_t>>> mips;
    uint32_t tri_first = 0, tri_count = 0xffffffffu;
    std::vector<uint8_t> chunk_mask;      // f184o_set_triangle_chunks: chunk_mask[t / 128] != 0 selects triangle t (empty = all)
    uint32_t row0 = 0, row1 = 0xffffffffu;
    uint32_t tile_first = 0, tile_stride = 1;   // of those rows, only 8-row tile rows t with t % stride == first
    uint32_t view_y0 = 0, view_h = 0;           // f184o_trace_views: rows [view_y0, view_y0 + view_h) are one view (0 = whole image)
    bool keep_samples = false;                   // f184o_trace_views: the cone-sample counter accumulates over the views
    const float* rands = nullptr;
    size_t n_rands = 0;
    // test hook (tests/test_refshader_pin.py): run the reference's own GS / PS text (oracle/_ref/libf184_refshaders.so)
    // between this file's fixed-function stages instead of the restated stages
    f184o_voxel_gs_hook gs_hook = nullptr;
    f184o_voxel_ps_hook ps_hook = nullptr;
    // static / dynamic split (f184o_static_cache_capture): the accumulator sums of the static geometry, dense here
    std::vector<float> cacheC, cacheN;
    bool cache_held = false;
    uint64_t cache_fragments = 0;
    float cache_cam[32] = {0}, last_vox_cam[32] = {0};
    uint64_t counters[F184_COUNTER_COUNT] = {0};
    float stage_ms[F184_STAGE_COUNT] = {0};
};

namespace orc {
template <class T> inline T* image_ptr(f184o_ctx* c, int slot) { return reinterpret_cast<T*>(c->img[slot].ptr); }
int ensure_image(f184o_ctx* c, int slot);
double now_ms();
}  // namespace orc
