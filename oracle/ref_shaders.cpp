// ref_shaders.cpp — TEST INFRASTRUCTURE.  Runs the reference's OWN shader text on the CPU.
//
// oracle/make_ref_shaders.py wraps Shader/Lighting/indirect.frag, Shader/GTAO/gtao.frag, Shader/GTAO/blur.frag,
// Shader/Lighting/blurX.frag / blurY.frag (+ bilateralBlur.inc, math.inc, EngineCommon.h) and the GLSL strings of
// Pipelang/Internal/main.lua (VoxelGS, BasicMaterial, VoxelPS) into oracle/_ref/gen/*.inc, read from /root/reference
// where they lie; this file includes them as struct bodies over oracle/glsl_shim.h and exposes one C entry point per
// pass.  Built by `make -C oracle ref` into oracle/_ref/libf184_refshaders.so (git-ignored, travels to the GPU box).
// The screen-space entry points take the same images the reference binds (MegaPipeline.cpp:225-284); the rasteriser
// that feeds VoxelGS/VoxelPS is the fixed-function stage defined in oracle_mode_r.cpp (f184o_debug_set_voxel_stage_hooks).
//
// Literals the reference hard-codes stay hard-coded: volume 128, shadow map 2048, 60 steps of 0.2 (indirect.frag:109,
// 133-146, 164), so comparisons against the parametrised restatement run at exactly that configuration.
#include <cstdio>
#include <vector>

#include "../include/f184.h"
#include "glsl_shim.h"

namespace glsl {

struct gl_PerVertex { vec4 gl_Position; };
struct gl_in_array
{
    gl_PerVertex v[3];
    int length() const { return 3; }
    const gl_PerVertex& operator[](int i) const { return v[i]; }
};

struct indirect_frag
{
#include "_ref/gen/indirect_frag.inc"
};
struct gtao_frag
{
    vec4 gl_FragCoord;
#include "_ref/gen/gtao_frag.inc"
};
struct gtao_blur_frag
{
#include "_ref/gen/gtao_blur_frag.inc"
};
struct blurX_frag
{
#include "_ref/gen/blurX_frag.inc"
};
struct blurY_frag
{
#include "_ref/gen/blurY_frag.inc"
};
struct aggregateLights_frag
{
#include "_ref/gen/aggregateLights_frag.inc"
};
struct color_frag
{
    vec4 gl_FragCoord;
#include "_ref/gen/color_frag.inc"
};
struct voxel_gs
{
    gl_in_array gl_in;
    vec4 gl_Position;
    // what EmitVertex() captures: the stage's outputs at the time of the call
    struct emitted { vec4 pos; vec3 normal; vec2 uv; } out[3];
    int n_emitted = 0;
    void EmitVertex();
    void EndPrimitive() {}
#include "_ref/gen/voxel_gs.inc"
};
inline void voxel_gs::EmitVertex()
{
    if (n_emitted < 3) { out[n_emitted].pos = gl_Position; out[n_emitted].normal = iNormal; out[n_emitted].uv = iTexCoord0; }
    n_emitted++;
}
// implicit-derivative texture() of the pixel stage: the rasteriser hands the per-triangle uv derivatives in
struct frag_derivs { bool on; vec2 ddx, ddy; };
static thread_local frag_derivs g_derivs = {false, vec2(), vec2()};
struct voxel_ps
{
    vec4 gl_FragCoord;
    static vec4 texture(const sampler2D& st, const vec2& uv)
    {
        if (g_derivs.on && st.t.level_data) return textureGrad(st, uv, g_derivs.ddx, g_derivs.ddy);
        return glsl::texture(st, uv);
    }
#include "_ref/gen/voxel_ps.inc"
};

static mat4 M(const float* p) { return mat4(p); }
static texture2D tex2d(const void* data, int w, int h, int format)
{
    texture2D t; t.data = data; t.w = w; t.h = h; t.format = format; return t;
}
static void store_rgba16f(uint16_t* o, const vec4& c)
{
    o[0] = dm_f32_to_f16(c.x); o[1] = dm_f32_to_f16(c.y); o[2] = dm_f32_to_f16(c.z); o[3] = dm_f32_to_f16(c.w);
}

}  // namespace glsl

using namespace glsl;

// bench.py's bounded CPU sample: the screen passes run rows [g_row0, g_row1) only (default: all)
static int g_row0 = 0, g_row1 = 1 << 30;
static inline int row_begin() { return g_row0 < 0 ? 0 : g_row0; }
static inline int row_end(int H) { return g_row1 < H ? g_row1 : H; }

// indirect_blurX (dir 0) / indirect_blurY (dir 1), MegaPipeline.cpp:270-284
template <class S> static void run_blur(const f184_engine_miscs* m, const uint16_t* src, const float* depth, int W, int H, uint16_t* out)
{
    S::s.wrap = 0;                                                    // GlobalLinearSamplerClamped, MegaPipeline.cpp:271,279
    S::t_indirect = tex2d(src, W, H, TEX_RGBA16F); S::t_depth = tex2d(depth, W, H, TEX_R32F);
    S::resolution = vec2(m->resolution[0], m->resolution[1]); S::frameCount = m->frameCount; S::frameTime = m->frameTime;
#pragma omp parallel for schedule(static)
    for (int y = row_begin(); y < row_end(H); y++)
        for (int x = 0; x < W; x++)
        {
            S sh;
            sh.inUV = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
            sh.main();
            store_rgba16f(&out[4 * ((size_t)y * W + x)], sh.outColor);
        }
}
extern "C" {

void refsh_set_rows(int y0, int y1) { g_row0 = y0; g_row1 = y1; }

// lighting_indirect, MegaPipeline.cpp:252-268.  history = indirectTemporalImage (all zero on the first frame, :197-204).
int refsh_indirect(const f184_trace_constants* k, const float* depth, const uint16_t* normals, const float* shadow,
                   const uint16_t* voxels, const uint16_t* history, int W, int H, uint16_t* out)
{
    typedef indirect_frag S;
    S::s.wrap = 1;                                                    // GlobalLinearSampler, MegaPipeline.cpp:253
    S::t_depth = tex2d(depth, W, H, TEX_R32F); S::t_normals = tex2d(normals, W, H, TEX_RGBA16_UNORM);
    S::t_shadow = tex2d(shadow, 2048, 2048, TEX_R32F); S::temporal = tex2d(history, W, H, TEX_RGBA16F);
    S::voxels.data = const_cast<uint16_t*>(voxels); S::voxels.n = 128;
    S::InvProj = M(k->view.InvProj); S::ViewMat = M(k->view.ViewMat); S::ProjMat = M(k->view.ProjMat);
    S::InvModelView = M(k->ext.InvModelView); S::ShadowView = M(k->ext.ShadowView); S::ShadowProj = M(k->ext.ShadowProj);
    S::VoxelView = M(k->ext.VoxelView); S::VoxelProj = M(k->ext.VoxelProj);
    S::sun.luminance = vec3(k->sun.luminance[0], k->sun.luminance[1], k->sun.luminance[2]);
    S::sun.position = vec3(k->sun.position[0], k->sun.position[1], k->sun.position[2]);
    S::prevProjection = M(k->prev.PrevProjection); S::prevModelView = M(k->prev.PrevModelView);
    S::resolution = vec2(k->miscs.resolution[0], k->miscs.resolution[1]);
    S::frameCount = k->miscs.frameCount; S::frameTime = k->miscs.frameTime;
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = row_begin(); y < row_end(H); y++)
        for (int x = 0; x < W; x++)
        {
            S sh;
            sh.inUV = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);     // Shader/Common/Quad.vert at the pixel centre
            sh.main();
            store_rgba16f(&out[4 * ((size_t)y * W + x)], sh.outColor);
        }
    return 0;
}

// gtao_visibility, MegaPipeline.cpp:225-233
int refsh_gtao(const f184_view_constants* view, const float* depth, const uint16_t* normals, int W, int H, uint16_t* out)
{
    typedef gtao_frag S;
    S::s.wrap = 1;
    S::t_depth = tex2d(depth, W, H, TEX_R32F); S::t_normals = tex2d(normals, W, H, TEX_RGBA16_UNORM);
    S::InvProj = M(view->InvProj); S::ViewMat = M(view->ViewMat); S::ProjMat = M(view->ProjMat);
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = row_begin(); y < row_end(H); y++)
        for (int x = 0; x < W; x++)
        {
            S sh;
            sh.gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
            sh.inUV = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
            sh.main();
            store_rgba16f(&out[4 * ((size_t)y * W + x)], sh.outColor);
        }
    return 0;
}

// gtao_blur, MegaPipeline.cpp:235-239
int refsh_gtao_blur(const uint16_t* ao_raw, int W, int H, uint16_t* out)
{
    typedef gtao_blur_frag S;
    S::s.wrap = 1;
    S::t_ao = tex2d(ao_raw, W, H, TEX_RGBA16F);
#pragma omp parallel for schedule(static)
    for (int y = row_begin(); y < row_end(H); y++)
        for (int x = 0; x < W; x++)
        {
            S sh;
            sh.inUV = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
            sh.main();
            store_rgba16f(&out[4 * ((size_t)y * W + x)], sh.outColor);
        }
    return 0;
}

int refsh_blur(int dir, const f184_engine_miscs* m, const uint16_t* src, const float* depth, int W, int H, uint16_t* out)
{
    if (dir == 0) run_blur<blurX_frag>(m, src, depth, W, H, out); else run_blur<blurY_frag>(m, src, depth, W, H, out);
    return 0;
}

// lighting_deferred, MegaPipeline.cpp:286-300.  Light lists as the reference uploads them (LightLists, :101-105).
int refsh_lighting_deferred(const f184_view_constants* view, const f184_extended_matrices* m, const f184_light_list* point,
                            const f184_light_list* directional, const uint8_t* albedo, const uint16_t* normals, const float* depth,
                            const float* shadow, const uint8_t* material, int W, int H, uint16_t* out)
{
    typedef aggregateLights_frag S;
    S::s.wrap = 1;                                                    // GlobalLinearSampler, MegaPipeline.cpp:287
    S::t_albedo = tex2d(albedo, W, H, TEX_RGBA8_UNORM); S::t_normals = tex2d(normals, W, H, TEX_RGBA16_UNORM);
    S::t_depth = tex2d(depth, W, H, TEX_R32F); S::t_shadow = tex2d(shadow, 2048, 2048, TEX_R32F);
    S::t_material = tex2d(material, W, H, TEX_RGBA8_UNORM);
    S::InvProj = M(view->InvProj); S::ViewMat = M(view->ViewMat); S::ProjMat = M(view->ProjMat);
    S::InvModelView = M(m->InvModelView); S::ShadowView = M(m->ShadowView); S::ShadowProj = M(m->ShadowProj);
    S::VoxelView = M(m->VoxelView); S::VoxelProj = M(m->VoxelProj);
    auto fill = [](auto& dst, const f184_light_list* src) {
        dst.numLights = src ? src->numLights : 0;
        for (int i = 0; i < dst.numLights; i++)
        {
            dst.lights[i].luminance = vec3(src->lights[i].luminance[0], src->lights[i].luminance[1], src->lights[i].luminance[2]);
            dst.lights[i].position = vec3(src->lights[i].position[0], src->lights[i].position[1], src->lights[i].position[2]);
        }
    };
    fill(S::Point, point); fill(S::Directional, directional);
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = row_begin(); y < row_end(H); y++)
        for (int x = 0; x < W; x++)
        {
            S sh;
            sh.inUV = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
            sh.main();
            store_rgba16f(&out[4 * ((size_t)y * W + x)], sh.lightingBuffer);
        }
    return 0;
}

// gtao_color, MegaPipeline.cpp:302-319: target 0 -> out_color, target 1 (taaImageA) -> out_taa
int refsh_composite(const f184_trace_constants* k, const uint8_t* albedo, const uint16_t* ao, const float* depth, const uint16_t* lighting,
                    const float* shadow, const uint16_t* indirect, const uint16_t* taa, int W, int H, uint16_t* out_color, uint16_t* out_taa)
{
    typedef color_frag S;
    S::s.wrap = 0;                                                    // GlobalLinearSamplerClamped, MegaPipeline.cpp:303
    S::t_albedo = tex2d(albedo, W, H, TEX_RGBA8_UNORM); S::t_ao = tex2d(ao, W, H, TEX_RGBA16F); S::t_depth = tex2d(depth, W, H, TEX_R32F);
    S::t_lighting = tex2d(lighting, W, H, TEX_RGBA16F); S::t_shadow = tex2d(shadow, 2048, 2048, TEX_R32F);
    S::t_indirect = tex2d(indirect, W, H, TEX_RGBA16F); S::taaBuffer = tex2d(taa, W, H, TEX_RGBA16F);
    S::InvProj = M(k->view.InvProj); S::ViewMat = M(k->view.ViewMat); S::ProjMat = M(k->view.ProjMat);
    S::InvModelView = M(k->ext.InvModelView); S::ShadowView = M(k->ext.ShadowView); S::ShadowProj = M(k->ext.ShadowProj);
    S::VoxelView = M(k->ext.VoxelView); S::VoxelProj = M(k->ext.VoxelProj);
    S::sun.luminance = vec3(k->sun.luminance[0], k->sun.luminance[1], k->sun.luminance[2]);
    S::sun.position = vec3(k->sun.position[0], k->sun.position[1], k->sun.position[2]);
    S::prevProjection = M(k->prev.PrevProjection); S::prevModelView = M(k->prev.PrevModelView);
    S::resolution = vec2(k->miscs.resolution[0], k->miscs.resolution[1]); S::frameCount = k->miscs.frameCount; S::frameTime = k->miscs.frameTime;
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = row_begin(); y < row_end(H); y++)
        for (int x = 0; x < W; x++)
        {
            S sh;
            sh.gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
            sh.inUV = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
            sh.main();
            store_rgba16f(&out_color[4 * ((size_t)y * W + x)], sh.outColor);
            store_rgba16f(&out_taa[4 * ((size_t)y * W + x)], sh.outTAA);
        }
    return 0;
}

// ---- voxel pass stages, signatures = f184o_voxel_gs_hook / f184o_voxel_ps_hook (oracle_common.h)
void refsh_voxel_gs(const float* view16, const float* proj16, const float* model16, const float* pos9, const float* nrm9,
                    const float* uv6, float* out_clip12, float* out_nrm9, float* out_uv6, uint32_t* out_orientation)
{
    typedef voxel_gs S;
    S::ViewMat = M(view16); S::ProjMat = M(proj16); S::ModelMat = M(model16);
    S sh;
    for (int i = 0; i < 3; i++)
    {
        // StaticMeshPassThruVS, main.lua:60-75: gl_Position = vec4(Position, 1), attributes passed through
        sh.gl_in.v[i].gl_Position = vec4(pos9[3 * i], pos9[3 * i + 1], pos9[3 * i + 2], 1.0f);
        sh.vgNormal[i] = vec3(nrm9[3 * i], nrm9[3 * i + 1], nrm9[3 * i + 2]);
        sh.vgTexCoord0[i] = vec2(uv6[2 * i], uv6[2 * i + 1]);
    }
    sh.main();
    for (int i = 0; i < 3; i++)
    {
        const vec4& p = sh.out[i].pos;
        out_clip12[4 * i] = p.x; out_clip12[4 * i + 1] = p.y; out_clip12[4 * i + 2] = p.z; out_clip12[4 * i + 3] = p.w;
        out_nrm9[3 * i] = sh.out[i].normal.x; out_nrm9[3 * i + 1] = sh.out[i].normal.y; out_nrm9[3 * i + 2] = sh.out[i].normal.z;
        out_uv6[2 * i] = sh.out[i].uv.x; out_uv6[2 * i + 1] = sh.out[i].uv.y;
    }
    *out_orientation = sh.iOrientation;
}

struct refsh_ps_material
{
    float factor[4];
    uint32_t use_textures;
    uint32_t tex_w, tex_h, tex_levels;
    const uint8_t* const* level_data;
};
int refsh_voxel_ps(const float* fragcoord4, const float* normal3, const float* uv2, const float* duvdx2, const float* duvdy2,
                   uint32_t orientation, const refsh_ps_material* m, uint16_t* voxels, uint32_t grid_n)
{
    typedef voxel_ps S;
    static const uint8_t zero_texel[4] = {0, 0, 0, 0};
    static const void* zero_levels[1] = {zero_texel};
    S::GlobalLinearSampler.wrap = 1;                                  // MegaPipeline.cpp:39-43
    texture2D t;
    if (m->tex_levels) { t = tex2d(m->level_data[0], (int)m->tex_w, (int)m->tex_h, TEX_RGBA8_UNORM); t.levels = (int)m->tex_levels; t.level_data = (const void* const*)m->level_data; }
    else { t = tex2d(zero_texel, 1, 1, TEX_RGBA8_UNORM); t.levels = 1; t.level_data = zero_levels; }
    S::BaseColorTex = t; S::MetallicRoughnessTex = t;                 // the second fetch does not reach the stored texel
    S::BaseColorFactor = vec4(m->factor[0], m->factor[1], m->factor[2], m->factor[3]);
    S::MetallicRoughness = vec4(1.0f); S::UseTextures = m->use_textures != 0;
    S::voxels.data = voxels; S::voxels.n = (int)grid_n; S::voxels.stores = 0;
    g_derivs.on = true; g_derivs.ddx = vec2(duvdx2[0], duvdx2[1]); g_derivs.ddy = vec2(duvdy2[0], duvdy2[1]);
    S sh;
    sh.gl_FragCoord = vec4(fragcoord4[0], fragcoord4[1], fragcoord4[2], fragcoord4[3]);
    sh.iNormal = vec3(normal3[0], normal3[1], normal3[2]);
    sh.iTexCoord0 = vec2(uv2[0], uv2[1]);
    sh.iOrientation = orientation;
    try { sh.main(); } catch (const glsl_discard&) { return 0; }
    return (int)S::voxels.stores;
}

}  // extern "C"
