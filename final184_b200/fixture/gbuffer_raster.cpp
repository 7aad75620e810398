// gbuffer_raster.cpp — INPUT SYNTHESISER (libf184_fixture.so), not part of the hot path and not the oracle.
//
// The hot path consumes images the Vulkan renderer produces before it: the G-buffer
// (CGBufferRenderer, Foreground/Renderer/GBufferRenderer.cpp:43-46; shaders Pipelang/Internal/main.lua
// :32-58 StaticMeshVS, :179-206 BasicMaterial, :227-240 GBufferPS; targets MegaPipeline.cpp:412-468) and the
// shadow depth map (CZOnlyRenderer; main.lua:146-156, 208-221; MegaPipeline.cpp:388-410).  Those passes
// stay Vulkan and are out of scope, but tests and the benchmark need their outputs,
// and there is no Vulkan here.  This is a plain CPU software rasteriser that stands in for them: it
// feeds the SAME arrays to the CUDA path and to the CPU oracle, so it cannot bias a parity result.
//
// Semantics followed: cull none, depth test less, clear depth 1 / colour 0, alpha discard < 0.05,
// TAA jitter of main.lua:44-53, view-space normal n*0.5+0.5 in RGBA16_UNORM, material (0, rough, metal, 0)
// with the importer's defaults roughness = metallic = 1 (glTFSceneImporter.cpp:186-196).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/f184.h"

namespace {

struct V4 { float x, y, z, w; };
struct M4 { float m[16]; };
inline M4 load(const float* p) { M4 r; memcpy(r.m, p, 64); return r; }
inline V4 mul(const M4& M, V4 v)
{
    return {M.m[0] * v.x + M.m[4] * v.y + M.m[8] * v.z + M.m[12] * v.w, M.m[1] * v.x + M.m[5] * v.y + M.m[9] * v.z + M.m[13] * v.w,
            M.m[2] * v.x + M.m[6] * v.y + M.m[10] * v.z + M.m[14] * v.w, M.m[3] * v.x + M.m[7] * v.y + M.m[11] * v.z + M.m[15] * v.w};
}
inline M4 matmul(const M4& A, const M4& B)
{
    M4 C;
    for (int j = 0; j < 4; j++)
    {
        V4 c = mul(A, V4{B.m[4 * j], B.m[4 * j + 1], B.m[4 * j + 2], B.m[4 * j + 3]});
        C.m[4 * j] = c.x; C.m[4 * j + 1] = c.y; C.m[4 * j + 2] = c.z; C.m[4 * j + 3] = c.w;
    }
    return C;
}

struct Tex { uint32_t w = 0, h = 0; std::vector<std::vector<uint8_t>> lv; };
struct Mat { float factor[4] = {1, 1, 1, 1}; int32_t tex = -1; uint32_t use = 1; };

struct Vtx { V4 clip; float nx, ny, nz, u, v; };

inline Vtx lerp(const Vtx& a, const Vtx& b, float t)
{
    Vtx r;
    r.clip = {a.clip.x + (b.clip.x - a.clip.x) * t, a.clip.y + (b.clip.y - a.clip.y) * t, a.clip.z + (b.clip.z - a.clip.z) * t,
              a.clip.w + (b.clip.w - a.clip.w) * t};
    r.nx = a.nx + (b.nx - a.nx) * t; r.ny = a.ny + (b.ny - a.ny) * t; r.nz = a.nz + (b.nz - a.nz) * t;
    r.u = a.u + (b.u - a.u) * t; r.v = a.v + (b.v - a.v) * t;
    return r;
}

}  // namespace

struct f184fx_ctx
{
    std::vector<float> pos, nrm, uv, model_mats;
    std::vector<uint32_t> idx;
    std::vector<uint16_t> tri_mat, tri_model;
    uint32_t n_verts = 0, n_tris = 0, n_models = 0;
    std::vector<Tex> tex;
    std::vector<Mat> mat;
};

static void sample(const Tex& t, float u, float v, float lod, float out[4])
{
    int l = (int)std::min<float>(std::max(lod + 0.5f, 0.0f), (float)std::min<size_t>(4, t.lv.size() - 1));
    uint32_t w = std::max(1u, t.w >> l), h = std::max(1u, t.h >> l);
    float x = u * w - 0.5f, y = v * h - 0.5f;
    float x0 = floorf(x), y0 = floorf(y), fx = x - x0, fy = y - y0;
    auto wr = [](int i, int n) { int m = i % n; return m < 0 ? m + n : m; };
    int xa = wr((int)x0, w), xb = wr((int)x0 + 1, w), ya = wr((int)y0, h), yb = wr((int)y0 + 1, h);
    const uint8_t* p = t.lv[l].data();
    for (int c = 0; c < 4; c++)
    {
        float a = p[4 * (ya * w + xa) + c], b = p[4 * (ya * w + xb) + c], cc = p[4 * (yb * w + xa) + c], d = p[4 * (yb * w + xb) + c];
        out[c] = ((a * (1 - fx) + b * fx) * (1 - fy) + (cc * (1 - fx) + d * fx) * fy) / 255.0f;
    }
}

// Rasterise every triangle into [y0,y1) of the targets.  `normals`/`albedo`/`material` may be null (shadow pass).
static void raster_band(const f184fx_ctx* c, const M4& View, const M4& Proj, float jx, float jy, uint32_t W, uint32_t H,
                        uint32_t y0, uint32_t y1, float* depth, uint16_t* normals, uint8_t* albedo, uint8_t* material)
{
    std::vector<M4> PVM(c->n_models), NM(c->n_models);
    M4 PV = matmul(Proj, View);
    for (uint32_t m = 0; m < c->n_models; m++)
    {
        M4 Mm = load(&c->model_mats[16 * m]);
        PVM[m] = matmul(PV, Mm);
        NM[m] = matmul(View, Mm);    // mat3(ViewMat) * mat3(ModelMat)
    }
    for (uint32_t t = 0; t < c->n_tris; t++)
    {
        const uint32_t* id = &c->idx[3 * t];
        const M4& pvm = PVM[c->tri_model[t]];
        const M4& nm = NM[c->tri_model[t]];
        Vtx poly[8];
        int np = 3;
        for (int i = 0; i < 3; i++)
        {
            const float* p = &c->pos[3 * id[i]];
            V4 cl = mul(pvm, V4{p[0], p[1], p[2], 1.0f});
            cl.x += jx; cl.y += jy;
            const float* n = &c->nrm[3 * id[i]];
            float nx = nm.m[0] * n[0] + nm.m[4] * n[1] + nm.m[8] * n[2];
            float ny = nm.m[1] * n[0] + nm.m[5] * n[1] + nm.m[9] * n[2];
            float nz = nm.m[2] * n[0] + nm.m[6] * n[1] + nm.m[10] * n[2];
            float l = sqrtf(nx * nx + ny * ny + nz * nz);
            if (l > 0) { nx /= l; ny /= l; nz /= l; }
            poly[i] = Vtx{cl, nx, ny, nz, c->uv[2 * id[i]], c->uv[2 * id[i] + 1]};
        }
        // clip against z >= 0 (Vulkan near plane) and w > eps
        for (int plane = 0; plane < 2; plane++)
        {
            Vtx out[8];
            int no = 0;
            auto dist = [&](const Vtx& v) { return plane == 0 ? v.clip.z : v.clip.w - 1e-6f; };
            for (int i = 0; i < np; i++)
            {
                const Vtx& a = poly[i]; const Vtx& b = poly[(i + 1) % np];
                float da = dist(a), db = dist(b);
                if (da >= 0) out[no++] = a;
                if ((da >= 0) != (db >= 0)) out[no++] = lerp(a, b, da / (da - db));
            }
            np = no;
            for (int i = 0; i < np; i++) poly[i] = out[i];
            if (np < 3) break;
        }
        if (np < 3) continue;
        const Mat& mat = c->mat[c->tri_mat[t]];
        const Tex* tex = (mat.use && mat.tex >= 0 && (size_t)mat.tex < c->tex.size()) ? &c->tex[mat.tex] : nullptr;
        for (int f = 1; f + 1 < np; f++)
        {
            const Vtx* v[3] = {&poly[0], &poly[f], &poly[f + 1]};
            double sx[3], sy[3], sz[3], iw[3];
            for (int i = 0; i < 3; i++)
            {
                iw[i] = 1.0 / v[i]->clip.w;
                sx[i] = (v[i]->clip.x * iw[i] * 0.5 + 0.5) * W;
                sy[i] = (v[i]->clip.y * iw[i] * 0.5 + 0.5) * H;
                sz[i] = v[i]->clip.z * iw[i];
            }
            double area = (sx[1] - sx[0]) * (sy[2] - sy[0]) - (sx[2] - sx[0]) * (sy[1] - sy[0]);
            if (area == 0 || !std::isfinite(area)) continue;
            double minx = std::min({sx[0], sx[1], sx[2]}), maxx = std::max({sx[0], sx[1], sx[2]});
            double miny = std::min({sy[0], sy[1], sy[2]}), maxy = std::max({sy[0], sy[1], sy[2]});
            int px0 = (int)std::max(0.0, std::ceil(minx - 0.5)), px1 = (int)std::min((double)W - 1, std::floor(maxx - 0.5));
            int py0 = (int)std::max((double)y0, std::ceil(miny - 0.5)), py1 = (int)std::min((double)y1 - 1, std::floor(maxy - 0.5));
            if (px0 > px1 || py0 > py1) continue;
            double inv_area = 1.0 / area;
            for (int py = py0; py <= py1; py++)
                for (int px = px0; px <= px1; px++)
                {
                    double cx = px + 0.5, cy = py + 0.5;
                    double w0 = ((sx[2] - sx[1]) * (cy - sy[1]) - (sy[2] - sy[1]) * (cx - sx[1])) * inv_area;
                    double w1 = ((sx[0] - sx[2]) * (cy - sy[2]) - (sy[0] - sy[2]) * (cx - sx[2])) * inv_area;
                    double w2 = 1.0 - w0 - w1;
                    if (w0 < 0 || w1 < 0 || w2 < 0) continue;
                    float z = (float)(w0 * sz[0] + w1 * sz[1] + w2 * sz[2]);
                    if (!(z >= 0.0f && z <= 1.0f)) continue;
                    size_t o = (size_t)py * W + px;
                    if (!(z < depth[o])) continue;
                    double pw = w0 * iw[0] + w1 * iw[1] + w2 * iw[2];
                    double b0 = w0 * iw[0] / pw, b1 = w1 * iw[1] / pw, b2 = w2 * iw[2] / pw;
                    float u = (float)(b0 * v[0]->u + b1 * v[1]->u + b2 * v[2]->u);
                    float vv = (float)(b0 * v[0]->v + b1 * v[1]->v + b2 * v[2]->v);
                    float base[4] = {mat.factor[0], mat.factor[1], mat.factor[2], mat.factor[3]};
                    if (mat.use)
                    {
                        float s[4] = {0, 0, 0, 0};
                        if (tex)
                        {
                            // LOD from a one-pixel finite difference of the (screen-affine) barycentrics
                            double w0x = w0 - (sy[2] - sy[1]) * inv_area, w1x = w1 - (sy[0] - sy[2]) * inv_area, w2x = 1.0 - w0x - w1x;
                            double pwx = w0x * iw[0] + w1x * iw[1] + w2x * iw[2];
                            float ux = (float)((w0x * iw[0] * v[0]->u + w1x * iw[1] * v[1]->u + w2x * iw[2] * v[2]->u) / pwx);
                            float vx = (float)((w0x * iw[0] * v[0]->v + w1x * iw[1] * v[1]->v + w2x * iw[2] * v[2]->v) / pwx);
                            float rho = std::max(fabsf(ux - u) * tex->w, fabsf(vx - vv) * tex->h);
                            float lod = rho > 1.0f ? log2f(rho) : 0.0f;
                            sample(*tex, u, vv, lod, s);
                        }
                        for (int k = 0; k < 4; k++) base[k] = s[k] * mat.factor[k];
                        if (base[3] < 0.05f) continue;     // discard (main.lua:199, :216)
                    }
                    depth[o] = z;
                    if (normals)
                    {
                        float nx = (float)(b0 * v[0]->nx + b1 * v[1]->nx + b2 * v[2]->nx);
                        float ny = (float)(b0 * v[0]->ny + b1 * v[1]->ny + b2 * v[2]->ny);
                        float nz = (float)(b0 * v[0]->nz + b1 * v[1]->nz + b2 * v[2]->nz);
                        auto q16 = [](float f) { float t = f * 0.5f + 0.5f; t = t < 0 ? 0 : (t > 1 ? 1 : t); return (uint16_t)lrintf(t * 65535.0f); };
                        normals[4 * o] = q16(nx); normals[4 * o + 1] = q16(ny); normals[4 * o + 2] = q16(nz); normals[4 * o + 3] = 0;
                    }
                    if (albedo)
                    {
                        auto q8 = [](float f) { f = f < 0 ? 0 : (f > 1 ? 1 : f); return (uint8_t)lrintf(f * 255.0f); };
                        albedo[4 * o] = q8(base[0]); albedo[4 * o + 1] = q8(base[1]); albedo[4 * o + 2] = q8(base[2]); albedo[4 * o + 3] = 255;
                    }
                    if (material) { material[4 * o] = 0; material[4 * o + 1] = 255; material[4 * o + 2] = 255; material[4 * o + 3] = 0; }
                }
        }
    }
}

#include <omp.h>
extern "C" {

// torchrun exports OMP_NUM_THREADS=1 to every rank; the caller says how many threads the synthesiser may use
void f184fx_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }


int f184fx_create(f184fx_ctx** out) { *out = new f184fx_ctx(); return 0; }
void f184fx_destroy(f184fx_ctx* c) { delete c; }

int f184fx_scene_upload(f184fx_ctx* c, const f184_scene_desc* s)
{
    c->n_verts = s->n_verts; c->n_tris = s->n_tris; c->n_models = s->n_models;
    c->pos.assign(s->positions, s->positions + 3ull * s->n_verts);
    c->nrm.assign(s->normals, s->normals + 3ull * s->n_verts);
    c->uv.assign(s->uvs, s->uvs + 2ull * s->n_verts);
    c->idx.assign(s->indices, s->indices + 3ull * s->n_tris);
    c->tri_mat.assign(s->tri_material, s->tri_material + s->n_tris);
    c->tri_model.assign(s->tri_model, s->tri_model + s->n_tris);
    c->model_mats.assign(s->model_mats, s->model_mats + 16ull * s->n_models);
    return 0;
}

int f184fx_texture_upload(f184fx_ctx* c, uint32_t id, const uint8_t* rgba, uint32_t w, uint32_t h)
{
    if (c->tex.size() <= id) c->tex.resize(id + 1);
    Tex& t = c->tex[id];
    t.w = w; t.h = h; t.lv.clear();
    t.lv.emplace_back(rgba, rgba + 4ull * w * h);
    uint32_t sw = w, sh = h;
    while (sw > 1 && sh > 1 && t.lv.size() < 5)
    {
        uint32_t dw = sw / 2, dh = sh / 2;
        const auto& s = t.lv.back();
        std::vector<uint8_t> d(4ull * dw * dh);
        for (uint32_t y = 0; y < dh; y++)
            for (uint32_t x = 0; x < dw; x++)
                for (int ch = 0; ch < 4; ch++)
                    d[4 * (y * dw + x) + ch] = (uint8_t)((s[4 * (2 * y * sw + 2 * x) + ch] + s[4 * (2 * y * sw + 2 * x + 1) + ch] +
                                                          s[4 * ((2 * y + 1) * sw + 2 * x) + ch] + s[4 * ((2 * y + 1) * sw + 2 * x + 1) + ch] + 2) >> 2);
        t.lv.push_back(std::move(d));
        sw = dw; sh = dh;
    }
    return 0;
}

int f184fx_material_set(f184fx_ctx* c, uint32_t id, const float factor[4], int32_t tex, uint32_t use)
{
    if (c->mat.size() <= id) c->mat.resize(id + 1);
    memcpy(c->mat[id].factor, factor, 16);
    c->mat[id].tex = tex; c->mat[id].use = use;
    return 0;
}

// G-buffer pass stand-in (MegaPipeline.cpp:178-186).  Outputs are W*H, row-major; depth cleared to 1, colour to 0.
int f184fx_render_gbuffer(f184fx_ctx* c, const f184_view_constants* view, uint32_t frame_count, uint32_t W, uint32_t H,
                          float* depth, uint16_t* normals, uint8_t* albedo, uint8_t* material)
{
    static const float taa[4][2] = {{-1, 0}, {1, 0}, {0, 1}, {0, -1}};     // main.lua:44-49
    float jx = taa[frame_count % 4][0] / (float)W, jy = taa[frame_count % 4][1] / (float)H;
    for (size_t i = 0; i < (size_t)W * H; i++) depth[i] = 1.0f;
    if (normals) memset(normals, 0, (size_t)W * H * 8);
    if (albedo) memset(albedo, 0, (size_t)W * H * 4);
    if (material) memset(material, 0, (size_t)W * H * 4);
    M4 View = load(view->ViewMat), Proj = load(view->ProjMat);
    const int bands = 64;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < bands; b++)
    {
        uint32_t y0 = (uint32_t)((uint64_t)H * b / bands), y1 = (uint32_t)((uint64_t)H * (b + 1) / bands);
        if (y0 < y1) raster_band(c, View, Proj, jx, jy, W, H, y0, y1, depth, normals, albedo, material);
    }
    return 0;
}

// Shadow Z-only pass stand-in (MegaPipeline.cpp:188-193).
int f184fx_render_shadow(f184fx_ctx* c, const f184_view_constants* view, uint32_t S, float* depth)
{
    for (size_t i = 0; i < (size_t)S * S; i++) depth[i] = 1.0f;
    M4 View = load(view->ViewMat), Proj = load(view->ProjMat);
    const int bands = 64;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < bands; b++)
    {
        uint32_t y0 = (uint32_t)((uint64_t)S * b / bands), y1 = (uint32_t)((uint64_t)S * (b + 1) / bands);
        if (y0 < y1) raster_band(c, View, Proj, 0.0f, 0.0f, S, S, y0, y1, depth, nullptr, nullptr, nullptr);
    }
    return 0;
}

}  // extern "C"
