// oracle_mode_r.cpp — TEST INFRASTRUCTURE (see oracle_common.h).
// CPU restatement of what the reference's shipped GPU programs compute ("mode R"):
//   voxelization   Pipelang/Internal/main.lua:60-75 (VS), 83-144 (GS), 179-206 (material), 242-275 (PS)
//                  + the fixed-function rasteriser state of Foreground/Renderer/VoxelizeRenderer.cpp:50-56
//   indirect march Shader/Lighting/indirect.frag (all), noise from Shader/math.inc
//   GTAO           Shader/GTAO/gtao.frag, Shader/GTAO/blur.frag, fastAcos/fastSqrt Shader/math.inc:14-33
//   blur tail      Shader/Lighting/bilateralBlur.inc, blurX.frag, blurY.frag
// Where Vulkan leaves behaviour to the driver the choice made here is the definition (SURVEY.md §8(c)
// items 1-8); each is marked "PINNED:" below.
#include <algorithm>
#include <cstdio>

#include "oracle_common.h"

using namespace orc;

// =================================================================================================
// texture sampling: GlobalLinearSampler — linear min/mag, linear mip, wrap, MaxLod 4
// (Foreground/Renderer/MegaPipeline.cpp:39-43)
// =================================================================================================
static inline int wrapi(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }   // n is a power of two

static V4 bilinear_level(const Texture& t, uint32_t level, float u, float v)
{
    uint32_t w = std::max(1u, t.w >> level), h = std::max(1u, t.h >> level);
    const uint8_t* px = t.levels[level].data();
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = wrapi(dm_f2i(x0f), (int)w), y0 = wrapi(dm_f2i(y0f), (int)h);
    int x1 = wrapi(x0 + 1, (int)w), y1 = wrapi(y0 + 1, (int)h);
    const uint8_t* t00 = px + 4 * ((size_t)y0 * w + x0);
    const uint8_t* t10 = px + 4 * ((size_t)y0 * w + x1);
    const uint8_t* t01 = px + 4 * ((size_t)y1 * w + x0);
    const uint8_t* t11 = px + 4 * ((size_t)y1 * w + x1);
    float r[4];
    for (int c = 0; c < 4; c++)
    {
        // PINNED: fp32 weights (no 8-bit weight quantisation), UNORM8 -> float by /255
        float a = (float)t00[c] / 255.0f, b = (float)t10[c] / 255.0f;
        float cc = (float)t01[c] / 255.0f, d = (float)t11[c] / 255.0f;
        float top = a * (1.0f - fx) + b * fx;
        float bot = cc * (1.0f - fx) + d * fx;
        r[c] = top * (1.0f - fy) + bot * fy;
    }
    return {r[0], r[1], r[2], r[3]};
}

// texture(sampler2D(BaseColorTex, GlobalLinearSampler), uv) with explicit derivatives
static V4 sample_trilinear(const Texture& t, float u, float v, float dudx, float dvdx, float dudy, float dvdy)
{
    // PINNED: Vulkan LOD: rho = max(|d(uv*size)/dx|, |d(uv*size)/dy|), lambda = log2(rho), clamp [0, MaxLod]
    float ax = dudx * (float)t.w, ay = dvdx * (float)t.h;
    float bx = dudy * (float)t.w, by = dvdy * (float)t.h;
    float mx = sqrtf(ax * ax + ay * ay), my = sqrtf(bx * bx + by * by);
    float rho = mx > my ? mx : my;
    float maxlod = (float)std::min<size_t>(4, t.levels.size() - 1);
    float lod = 0.0f;
    if (rho > 1.0f) lod = O_LOG2(rho);
    if (!(lod < maxlod)) lod = maxlod;
    float lf = floorf(lod);
    uint32_t l0 = (uint32_t)lf;
    float f = lod - lf;
    V4 c0 = bilinear_level(t, l0, u, v);
    if (f == 0.0f) return c0;
    V4 c1 = bilinear_level(t, l0 + 1, u, v);
    return {c0.x * (1.0f - f) + c1.x * f, c0.y * (1.0f - f) + c1.y * f, c0.z * (1.0f - f) + c1.z * f,
            c0.w * (1.0f - f) + c1.w * f};
}

// =================================================================================================
// voxelization
// =================================================================================================
struct VtxOut { float cx, cy, cz; V3 n; float u, v; };

extern "C" int orc_voxelize_r(f184o_ctx* c, const f184_view_constants* cam)
{
    double t0 = now_ms();
    int rc = ensure_image(c, F184_SLOT_VOXELS);
    if (rc) return rc;
    const uint32_t N = c->cfg.grid_n;
    uint16_t* vox = image_ptr<uint16_t>(c, F184_SLOT_VOXELS);
    // ClearImage(VoxelImage, 0): MegaPipeline.cpp:196
    memset(vox, 0, (size_t)N * N * N * 4);

    const M4 View = load_m4(cam->ViewMat), Proj = load_m4(cam->ProjMat);
    std::vector<M4> VM(c->n_models), MM(c->n_models);
    for (uint32_t m = 0; m < c->n_models; m++)
    {
        MM[m] = load_m4(&c->model_mats[16 * m]);
        VM[m] = matmul(View, MM[m]);                 // `ViewMat * ModelMat * p` associates left (main.lua:98)
    }
    const float Nf = (float)N, halfN = Nf * 0.5f, maxDepth = (float)(N - 1);
    uint64_t frags = 0;
    uint32_t first = c->tri_first, last = (uint32_t)std::min<uint64_t>((uint64_t)c->tri_first + c->tri_count, c->n_tris);

    for (uint32_t t = first; t < last; t++)
    {
        const uint32_t* id = &c->idx[3 * t];
        const M4& vm = VM[c->tri_model[t]];
        const M4& mm = MM[c->tri_model[t]];
        VtxOut vo[3];
        int orient;
        if (c->gs_hook)
        {
            float pos9[9], nrm9[9], uv6[6], clip[12], on[9], ouv[6];
            for (int i = 0; i < 3; i++)
            {
                memcpy(&pos9[3 * i], &c->pos[3 * id[i]], 12); memcpy(&nrm9[3 * i], &c->nrm[3 * id[i]], 12);
                memcpy(&uv6[2 * i], &c->uv[2 * id[i]], 8);
            }
            uint32_t o = 0;
            c->gs_hook(cam->ViewMat, cam->ProjMat, &c->model_mats[16 * c->tri_model[t]], pos9, nrm9, uv6, clip, on, ouv, &o);
            orient = (int)o;
            for (int i = 0; i < 3; i++)
            {
                vo[i].cx = clip[4 * i]; vo[i].cy = clip[4 * i + 1]; vo[i].cz = clip[4 * i + 2];     // w = 1 after the GS's divide
                vo[i].n = {on[3 * i], on[3 * i + 1], on[3 * i + 2]};
                vo[i].u = ouv[2 * i]; vo[i].v = ouv[2 * i + 1];
            }
        }
        else
        {
            // ---- VoxelGS, main.lua:94-115
            V3 vp[3];
            for (int i = 0; i < 3; i++)
            {
                const float* p = &c->pos[3 * id[i]];
                V4 q = mul(vm, V4{p[0], p[1], p[2], 1.0f});
                vp[i] = {q.x, q.y, q.z};
            }
            V3 fn = vabs(cross(vp[1] - vp[0], vp[2] - vp[0]));
            if (fn.x > fn.y) orient = (fn.x > fn.z) ? 1 : 0;
            else orient = (fn.y > fn.z) ? 2 : 0;
            // ---- main.lua:116-141
            for (int i = 0; i < 3; i++)
            {
                V4 g = mul(Proj, V4{vp[i].x, vp[i].y, vp[i].z, 1.0f});
                float gx = g.x / g.w, gy = g.y / g.w, gz = g.z / g.w;
                if (orient == 1) { float nx = gz * 2.0f - 1.0f; float nz = -gx / 2.0f + 0.5f; gx = nx; gz = nz; }
                else if (orient == 2) { float ny = 1.0f - gz * 2.0f; float nz = gy / 2.0f + 0.5f; gy = ny; gz = nz; }
                vo[i].cx = gx; vo[i].cy = gy; vo[i].cz = gz;
                const float* nn = &c->nrm[3 * id[i]];
                vo[i].n = normalize(mul3(mm, V3{nn[0], nn[1], nn[2]}));
                vo[i].u = c->uv[2 * id[i]]; vo[i].v = c->uv[2 * id[i] + 1];
            }
        }
        // ---- fixed-function rasteriser (viewport N x N, depth range [0,1], no cull, 1 sample)
        // PINNED: viewport transform x_f = x_ndc*N/2 + N/2, snap to 1/256 px, integer edge functions,
        // top-left rule, pixel-centre sample, depth-clip discards z outside [0,1] (no depth clamp).
        int64_t X[3], Y[3];
        bool bad = false;
        for (int i = 0; i < 3; i++)
        {
            float xf = vo[i].cx * halfN + halfN, yf = vo[i].cy * halfN + halfN;
            if (!(fabsf(xf) < 1048576.0f) || !(fabsf(yf) < 1048576.0f)) { bad = true; break; }   // also rejects NaN
            X[i] = (int64_t)rintf(xf * 256.0f);
            Y[i] = (int64_t)rintf(yf * 256.0f);
        }
        if (bad) continue;
        int64_t area = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0]);
        if (area == 0) continue;
        int i1 = 1, i2 = 2;
        if (area < 0) { i1 = 2; i2 = 1; area = -area; }
        const int ord[3] = {0, i1, i2};
        int64_t x[3], y[3];
        for (int k = 0; k < 3; k++) { x[k] = X[ord[k]]; y[k] = Y[ord[k]]; }
        int64_t minx = std::min({x[0], x[1], x[2]}), maxx = std::max({x[0], x[1], x[2]});
        int64_t miny = std::min({y[0], y[1], y[2]}), maxy = std::max({y[0], y[1], y[2]});
        // pixel px has its centre at 256*px + 128
        auto ceil_div = [](int64_t a, int64_t b) { return (a >= 0) ? (a + b - 1) / b : -((-a) / b); };
        auto floor_div = [](int64_t a, int64_t b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
        int64_t px0 = std::max<int64_t>(0, ceil_div(minx - 128, 256)), px1 = std::min<int64_t>(N - 1, floor_div(maxx - 128, 256));
        int64_t py0 = std::max<int64_t>(0, ceil_div(miny - 128, 256)), py1 = std::min<int64_t>(N - 1, floor_div(maxy - 128, 256));
        if (px0 > px1 || py0 > py1) continue;
        // edge k is opposite vertex k: from vertex (k+1)%3 to (k+2)%3; E = dx*(py-ay) - dy*(px-ax) >= 0 inside
        int64_t ex[3], ey[3], bias[3];
        for (int k = 0; k < 3; k++)
        {
            int a = (k + 1) % 3, b = (k + 2) % 3;
            ex[k] = x[b] - x[a]; ey[k] = y[b] - y[a];
            bool top_left = (ey[k] < 0) || (ey[k] == 0 && ex[k] > 0);
            bias[k] = top_left ? 0 : -1;
        }
        const float areaf = (float)area;
        // attribute gradients (affine: ortho camera, w = 1) for the implicit-derivative LOD
        float dbdx[3], dbdy[3];
        for (int k = 0; k < 3; k++) { dbdx[k] = (float)(-ey[k] * 256) / areaf; dbdy[k] = (float)(ex[k] * 256) / areaf; }
        const VtxOut& A = vo[ord[0]]; const VtxOut& B = vo[ord[1]]; const VtxOut& C = vo[ord[2]];
        float dudx = (A.u * dbdx[0] + B.u * dbdx[1]) + C.u * dbdx[2];
        float dvdx = (A.v * dbdx[0] + B.v * dbdx[1]) + C.v * dbdx[2];
        float dudy = (A.u * dbdy[0] + B.u * dbdy[1]) + C.u * dbdy[2];
        float dvdy = (A.v * dbdy[0] + B.v * dbdy[1]) + C.v * dbdy[2];
        const Material& mat = c->materials[c->tri_mat[t]];
        const Texture* tex = (mat.use_textures && mat.tex >= 0) ? &c->textures[mat.tex] : nullptr;

        for (int64_t py = py0; py <= py1; py++)
            for (int64_t px = px0; px <= px1; px++)
            {
                int64_t cxp = px * 256 + 128, cyp = py * 256 + 128;
                int64_t w[3];
                bool inside = true;
                for (int k = 0; k < 3; k++)
                {
                    int a = (k + 1) % 3;
                    w[k] = ex[k] * (cyp - y[a]) - ey[k] * (cxp - x[a]);
                    if (w[k] + bias[k] < 0) { inside = false; break; }
                }
                if (!inside) continue;
                float b0 = (float)w[0] / areaf, b1 = (float)w[1] / areaf, b2 = (float)w[2] / areaf;
                float z = (A.cz * b0 + B.cz * b1) + C.cz * b2;
                if (!(z >= 0.0f && z <= 1.0f)) continue;
                float u = (A.u * b0 + B.u * b1) + C.u * b2;
                float v = (A.v * b0 + B.v * b1) + C.v * b2;
                V3 n = {(A.n.x * b0 + B.n.x * b1) + C.n.x * b2, (A.n.y * b0 + B.n.y * b1) + C.n.y * b2,
                        (A.n.z * b0 + B.n.z * b1) + C.n.z * b2};
                if (c->ps_hook)
                {
                    f184o_ps_material pm{};
                    memcpy(pm.factor, mat.factor, 16); pm.use_textures = mat.use_textures;
                    const uint8_t* lv[16] = {nullptr};
                    if (tex)
                    {
                        pm.tex_w = tex->w; pm.tex_h = tex->h; pm.tex_levels = (uint32_t)std::min<size_t>(16, tex->levels.size());
                        for (uint32_t l = 0; l < pm.tex_levels; l++) lv[l] = tex->levels[l].data();
                        pm.level_data = lv;
                    }
                    const float fc[4] = {(float)px + 0.5f, (float)py + 0.5f, z, 1.0f}, nn[3] = {n.x, n.y, n.z}, uvv[2] = {u, v};
                    const float ddx[2] = {dudx, dvdx}, ddy[2] = {dudy, dvdy};
                    frags += (uint64_t)c->ps_hook(fc, nn, uvv, ddx, ddy, (uint32_t)orient, &pm, vox, N);
                    continue;
                }
                // ---- BasicMaterial, main.lua:188-205
                V4 base;
                if (!mat.use_textures) base = {mat.factor[0], mat.factor[1], mat.factor[2], mat.factor[3]};
                else
                {
                    V4 s = tex ? sample_trilinear(*tex, u, v, dudx, dvdx, dudy, dvdy) : V4{0, 0, 0, 0};
                    base = {s.x * mat.factor[0], s.y * mat.factor[1], s.z * mat.factor[2], s.w * mat.factor[3]};
                    if (base.w < 0.05f) continue;      // discard
                }
                // ---- VoxelPS, main.lua:249-273
                float fxc = (float)px + 0.5f, fyc = (float)py + 0.5f;
                float vx, vy, vz;
                if (orient == 0) { vx = fxc; vy = fyc; vz = z * maxDepth; }
                else if (orient == 1) { vx = (1.0f - z) * maxDepth; vy = fyc; vz = fxc; }
                else { vx = fxc; vy = z * maxDepth; vz = maxDepth - fyc; }
                int ix = dm_f2i(vx), iy = dm_f2i(vy), iz = dm_f2i(vz);
                // PINNED: out-of-range imageStore is dropped; uint() of a negative float saturates to 0
                if (ix < 0 || iy < 0 || iz < 0 || ix >= (int)N || iy >= (int)N || iz >= (int)N) continue;
                uint32_t pc = (dm_f2uint(base.x * 31.0f) << 11) | (dm_f2uint(base.y * 63.0f) << 5) | dm_f2uint(base.z * 31.0f);
                uint32_t pn = (dm_f2uint(n.x * 16.0f + 15.0f) << 11) | (dm_f2uint(n.y * 32.0f + 31.0f) << 5) |
                              dm_f2uint(n.z * 16.0f + 15.0f);
                // PINNED: the plain imageStore race (main.lua:273) resolves as "last writer in draw order wins"
                size_t o = ((size_t)iz * N + iy) * N + ix;
                vox[2 * o] = (uint16_t)pc;
                vox[2 * o + 1] = (uint16_t)pn;
                frags++;
            }
    }
    c->counters[F184_COUNTER_FRAGMENTS] = frags;
    c->stage_ms[F184_STAGE_VOXELIZE] = (float)(now_ms() - t0);
    return F184_OK;
}

// =================================================================================================
// indirect march
// =================================================================================================
namespace {

struct TraceCtx
{
    const f184_trace_constants* k;
    M4 InvProj, InvModelView, ShadowView, ShadowProj, w2voxel, prevModelView, prevProjection;
    const float* depth; const uint16_t* normals; const float* shadow; const uint16_t* vox; const uint16_t* hist;
    uint32_t W, H, S, N, steps;
    float step_size, iiTime;
    V2 res;
    V3 sunLum, sunPos;
    uint64_t march_steps;
};

// math.inc:83-87
inline float glsl_hash(float px, float py)
{
    float p3x = dm_fract(px * 0.2031f), p3y = dm_fract(py * 0.2031f), p3z = dm_fract(px * 0.2031f);
    float d = (p3x * (p3y + 19.19f) + p3y * (p3z + 19.19f)) + p3z * (p3x + 19.19f);
    p3x += d; p3y += d; p3z += d;
    return dm_fract((p3x + p3y) * p3z);
}
// math.inc:107-110
inline float nrand(float nx, float ny) { return dm_fract(O_SIN(nx * 12.9898f + ny * 78.233f) * 43758.5453f); }
// math.inc:189-194
inline float n4rand_ss(float nx, float ny, float iiTime)
{
    float t0 = 0.07f * dm_fract(iiTime);
    float t1 = 0.11f * dm_fract(iiTime + 0.573953f);
    float nrnd0 = nrand(nx + t0, ny + t0);
    float nrnd1 = nrand(nx + t1, ny + t1);
    return 0.23f * sqrtf(-O_LOG(nrnd0 + 0.00001f)) * O_COS(2.0f * 3.141592f * nrnd1) + 0.5f;
}
// math.inc:214-219
inline float blugausnoise2(float cx, float cy, float iiTime)
{
    float nrand1 = n4rand_ss(cx, cy, iiTime);
    float nrand0 = n4rand_ss(cx - 1.0f, cy, iiTime);
    float nrand2 = n4rand_ss(cx + 1.0f, cy, iiTime);
    return 2.0f * nrand1 - 0.5f * (nrand0 + nrand2);
}

struct Hit { V3 wpos, wnorm, brdf; bool hit; };

// indirect.frag:71-86
inline void make_coord_space(V3 n, V3& X, V3& Y, V3& Z)
{
    V3 z = n, h = n;
    if (fabsf(h.x) <= fabsf(h.y) && fabsf(h.x) <= fabsf(h.z)) h.x = 1.0f;
    else if (fabsf(h.y) <= fabsf(h.x) && fabsf(h.y) <= fabsf(h.z)) h.y = 1.0f;
    else h.z = 1.0f;
    z = normalize(z);
    V3 y = normalize(cross(h, z));
    V3 x = normalize(cross(z, y));
    X = x; Y = y; Z = z;
}

inline V3 voxel_pos(const TraceCtx& T, V3 p)
{
    V4 v = mul(T.w2voxel, V4{p.x, p.y, p.z, 1.0f});
    float Nf = (float)T.N;
    return {(v.x * 0.5f + 0.5f) * Nf, (v.y * 0.5f + 0.5f) * Nf, v.z * Nf};
}

// indirect.frag:104-184
V3 get_indirect(TraceCtx& T, V3 wpos, V3 wnorm, float seed, float uvx, float uvy, const float* ext_rands, Hit& hit)
{
    hit.hit = false;
    V3 Lo = {0, 0, 0};
    const float step_size = T.step_size;
    float rx, ry;
    if (ext_rands) { rx = ext_rands[0]; ry = ext_rands[1]; }
    else
    {
        float ra = glsl_hash(seed, seed);
        float ax = (uvx + ra) * T.res.x, ay = (uvy + ra) * T.res.y;
        rx = blugausnoise2(-ax, -ay, T.iiTime);
        ry = blugausnoise2(ax, ay, T.iiTime);
    }
    // hemisphereSample_cos, math.inc:76-81
    float phi = ry * 2.0f * 3.1415926f;
    float cosTheta = sqrtf(1.0f - rx);
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    V3 d = {O_COS(phi) * sinTheta, O_SIN(phi) * sinTheta, cosTheta};
    V3 X, Y, Z;
    make_coord_space(wnorm, X, Y, Z);
    V3 dir = {(X.x * d.x + Y.x * d.y) + Z.x * d.z, (X.y * d.x + Y.y * d.y) + Z.y * d.z, (X.z * d.x + Y.z * d.y) + Z.z * d.z};
    if (dot(dir, wnorm) < 0.0f) dir = neg(dir);
    float NdotD = dot(dir, wnorm);

    float dn = dot(dir, wnorm);
    float s1 = 1.0f + ry;
    V3 march_pos = {wpos.x + dir.x * s1 * step_size / dn, wpos.y + dir.y * s1 * step_size / dn, wpos.z + dir.z * s1 * step_size / dn};

    V3 sv = voxel_pos(T, wpos);
    int pvx = dm_f2i(sv.x), pvy = dm_f2i(sv.y), pvz = dm_f2i(sv.z);
    const float hi = (float)(T.N - 1);
    uint32_t i = 0;
    for (; i < T.steps; i++)
    {
        T.march_steps++;
        march_pos = {march_pos.x + dir.x * step_size, march_pos.y + dir.y * step_size, march_pos.z + dir.z * step_size};
        V3 vp = voxel_pos(T, march_pos);
        if (vp.x < 0 || vp.y < 0 || vp.z < 0 || vp.x > hi || vp.y > hi || vp.z > hi) break;
        // PINNED: ivec3(NaN) = 0; a NaN ray never leaves the loop early
        int ix = dm_f2i(vp.x), iy = dm_f2i(vp.y), iz = dm_f2i(vp.z);
        if (pvx != ix || pvy != iy || pvz != iz)
        {
            uint32_t r = 0, g = 0;
            if (ix >= 0 && iy >= 0 && iz >= 0 && ix < (int)T.N && iy < (int)T.N && iz < (int)T.N)
            {
                size_t o = ((size_t)iz * T.N + iy) * T.N + ix;
                r = T.vox[2 * o]; g = T.vox[2 * o + 1];
            }
            pvx = ix; pvy = iy; pvz = iz;
            if (r != 0)
            {
                // unpackColor565 + pow 2.2, indirect.frag:88-94,157
                V3 col = {O_POW((float)((r & 0xF800u) >> 11) / 31.0f, 2.2f), O_POW((float)((r & 0x7E0u) >> 5) / 63.0f, 2.2f),
                          O_POW((float)(r & 0x1Fu) / 31.0f, 2.2f)};
                // unpackNormal565, indirect.frag:96-102 (x from the low bits: the reference's swap, kept)
                V3 vn = normalize(V3{(float)(g & 0x1Fu) / 16.0f - 1.0f, (float)((g & 0x7E0u) >> 5) / 32.0f - 1.0f,
                                     (float)((g & 0xF800u) >> 11) / 16.0f - 1.0f});
                V3 sp = {march_pos.x + vn.x * 0.06f, march_pos.y + vn.y * 0.06f, march_pos.z + vn.z * 0.06f};
                V4 sv4 = mul(T.ShadowProj, mul(T.ShadowView, V4{sp.x, sp.y, sp.z, 1.0f}));
                float spx = sv4.x / sv4.w, spy = sv4.y / sv4.w, spz = sv4.z / sv4.w;
                spx = spx * 0.5f + 0.5f; spy = spy * 0.5f + 0.5f;
                int tx = dm_f2i(spx * (float)T.S), ty = dm_f2i(spy * (float)T.S);
                float shadowZ = 0.0f;     // PINNED: out-of-range texelFetch returns 0
                if (tx >= 0 && ty >= 0 && tx < (int)T.S && ty < (int)T.S) shadowZ = T.shadow[(size_t)ty * T.S + tx];
                float shade = dm_step(spz + 0.005f, shadowZ);
                float l = fabsf(dot(neg(T.sunPos), vn));
                float den = dm_max(0.01f, NdotD);
                V3 rr = {l * col.x / den, l * col.y / den, l * col.z / den};
                hit.brdf = rr;
                Lo = {Lo.x + T.sunLum.x * shade * rr.x, Lo.y + T.sunLum.y * shade * rr.y, Lo.z + T.sunLum.z * shade * rr.z};
                hit.wpos = march_pos; hit.wnorm = vn; hit.hit = true;
                break;
            }
        }
    }
    if (!hit.hit && i == T.steps)
    {
        float den = dm_max(0.01f, NdotD);
        float sm = dm_smoothstep(0.0f, 0.01f, NdotD);
        Lo = {Lo.x + 0.7f * 0.4f / den * sm, Lo.y + 0.8f * 0.4f / den * sm, Lo.z + 1.0f * 0.4f / den * sm};
    }
    return Lo;
}

inline float unorm16(uint16_t v) { return (float)v / 65535.0f; }

}  // namespace

extern "C" int orc_trace_r(f184o_ctx* c, const f184_trace_constants* k)
{
    double t0 = now_ms();
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_SHADOW, F184_SLOT_VOXELS, F184_SLOT_INDIRECT_OUT, F184_SLOT_INDIRECT_HISTORY})
    { int rc = ensure_image(c, s); if (rc) return rc; }
    TraceCtx T0;
    T0.k = k;
    T0.InvProj = load_m4(k->view.InvProj); T0.InvModelView = load_m4(k->ext.InvModelView);
    T0.ShadowView = load_m4(k->ext.ShadowView); T0.ShadowProj = load_m4(k->ext.ShadowProj);
    T0.w2voxel = matmul(load_m4(k->ext.VoxelProj), load_m4(k->ext.VoxelView));     // indirect.frag:127
    T0.prevModelView = load_m4(k->prev.PrevModelView); T0.prevProjection = load_m4(k->prev.PrevProjection);
    T0.depth = image_ptr<float>(c, F184_SLOT_DEPTH); T0.normals = image_ptr<uint16_t>(c, F184_SLOT_NORMALS);
    T0.shadow = image_ptr<float>(c, F184_SLOT_SHADOW); T0.vox = image_ptr<uint16_t>(c, F184_SLOT_VOXELS);
    uint16_t* hist = image_ptr<uint16_t>(c, F184_SLOT_INDIRECT_HISTORY);
    T0.hist = hist;
    T0.W = c->cfg.width; T0.H = c->cfg.height; T0.S = c->cfg.shadow_res; T0.N = c->cfg.grid_n; T0.steps = c->cfg.march_steps;
    T0.step_size = c->cfg.step_size;
    T0.iiTime = (float)k->miscs.frameCount * 0.03125f;                              // indirect.frag:111
    T0.res = {k->miscs.resolution[0], k->miscs.resolution[1]};
    T0.sunLum = {k->sun.luminance[0], k->sun.luminance[1], k->sun.luminance[2]};
    T0.sunPos = {k->sun.position[0], k->sun.position[1], k->sun.position[2]};
    T0.march_steps = 0;
    // first frame: the history image is cleared, MegaPipeline.cpp:197-204
    if (k->reset_history) memset(hist, 0, (size_t)T0.W * T0.H * 8);
    uint16_t* out = image_ptr<uint16_t>(c, F184_SLOT_INDIRECT_OUT);
    const uint32_t W = T0.W, H = T0.H;
    uint32_t y0 = c->row0, y1 = std::min(c->row1, H);
    const bool ext = (c->cfg.flags & F184_FLAG_EXTERNAL_RANDS) && c->rands;
    uint64_t total_steps = 0;

#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total_steps)
    for (int64_t y = y0; y < (int64_t)y1; y++)
    {
        if ((uint32_t)(y >> 3) % c->tile_stride != c->tile_first) continue;      // tile rows interleaved over the ranks
        TraceCtx T = T0;
        T.march_steps = 0;
        for (uint32_t x = 0; x < W; x++)
        {
            // full-screen triangle interpolant at the pixel centre (Shader/Common/Quad.vert)
            float uvx = ((float)x + 0.5f) / (float)W, uvy = ((float)y + 0.5f) / (float)H;
            // getCSpos, indirect.frag:44-53
            int dx = dm_f2i(uvx * (float)W), dy = dm_f2i(uvy * (float)H);
            float depth = (dx >= 0 && dy >= 0 && dx < (int)W && dy < (int)H) ? T.depth[(size_t)dy * W + dx] : 0.0f;
            V4 cp = mul(T.InvProj, V4{uvx * 2.0f - 1.0f, uvy * 2.0f - 1.0f, depth, 1.0f});
            V3 cspos = {cp.x / cp.w, cp.y / cp.w, cp.z / cp.w};
            V4 wp4 = mul(T.InvModelView, V4{cspos.x, cspos.y, cspos.z, 1.0f});
            V3 wpos = {wp4.x, wp4.y, wp4.z};
            // getNormal, indirect.frag:55-58.  PINNED: texture() at a texel centre returns that texel.
            const uint16_t* np = &T.normals[4 * ((size_t)y * W + x)];
            V3 raw = {fmaf(unorm16(np[0]), 2.0f, -1.0f), fmaf(unorm16(np[1]), 2.0f, -1.0f), fmaf(unorm16(np[2]), 2.0f, -1.0f)};
            V3 csnorm = normalize(raw);
            V3 wnorm = mul3(T.InvModelView, csnorm);

            Hit st{};
            V3 ind = {0, 0, 0};
            const float* er = ext ? &c->rands[16 * ((size_t)y * W + x)] : nullptr;
            for (int pair = 0; pair < 4; pair++)
            {
                V3 a = get_indirect(T, wpos, wnorm, (float)(2 * pair), uvx, uvy, er ? er + 4 * pair : nullptr, st);
                ind = ind + a;
                if (st.hit)
                {
                    V3 brdf = st.brdf;     // PINNED: left operand read before the call (indirect.frag:204)
                    V3 hw = st.wpos, hn = st.wnorm;
                    V3 b = get_indirect(T, hw, hn, (float)(2 * pair + 1), uvx, uvy, er ? er + 4 * pair + 2 : nullptr, st);
                    ind = ind + brdf * b;
                }
            }
            ind = ind * 0.25f;
            // temporal reprojection, indirect.frag:225-240
            V4 pc = mul(T.prevModelView, V4{wpos.x, wpos.y, wpos.z, 1.0f});
            V4 pp = mul(T.prevProjection, pc);
            float ru = pp.x / pp.w, rv = pp.y / pp.w;
            ru = ru * 0.5f + 0.5f; rv = rv * 0.5f + 0.5f;
            if (dm_clamp(ru, 0.0f, 1.0f) == ru && dm_clamp(rv, 0.0f, 1.0f) == rv)
            {
                // texture(temporal, uv): bilinear, wrap (GlobalLinearSampler), fp32 weights
                float fx = ru * (float)W - 0.5f, fy = rv * (float)H - 0.5f;
                float x0f = floorf(fx), y0f = floorf(fy);
                float wx = fx - x0f, wy = fy - y0f;
                int xi0 = dm_f2i(x0f), yi0 = dm_f2i(y0f);
                auto wrapn = [](int i, int n) { int m = i % n; return m < 0 ? m + n : m; };
                int xa = wrapn(xi0, (int)W), xb = wrapn(xi0 + 1, (int)W), ya = wrapn(yi0, (int)H), yb = wrapn(yi0 + 1, (int)H);
                float prev[4];
                for (int ch = 0; ch < 4; ch++)
                {
                    float a = dm_f16_to_f32(T.hist[4 * ((size_t)ya * W + xa) + ch]), b = dm_f16_to_f32(T.hist[4 * ((size_t)ya * W + xb) + ch]);
                    float cc = dm_f16_to_f32(T.hist[4 * ((size_t)yb * W + xa) + ch]), d = dm_f16_to_f32(T.hist[4 * ((size_t)yb * W + xb) + ch]);
                    prev[ch] = (a * (1.0f - wx) + b * wx) * (1.0f - wy) + (cc * (1.0f - wx) + d * wx) * wy;
                }
                float bw = 0.95f * dm_smoothstep(0.0f, 1.0f, 1.0f - fabsf(prev[3] + cspos.z));
                ind = {dm_clamp(dm_mix(ind.x, prev[0], bw), 0.0f, 16.0f), dm_clamp(dm_mix(ind.y, prev[1], bw), 0.0f, 16.0f),
                       dm_clamp(dm_mix(ind.z, prev[2], bw), 0.0f, 16.0f)};
            }
            uint16_t* o = &out[4 * ((size_t)y * W + x)];
            o[0] = dm_f32_to_f16(ind.x); o[1] = dm_f32_to_f16(ind.y); o[2] = dm_f32_to_f16(ind.z); o[3] = dm_f32_to_f16(-cspos.z);
        }
        total_steps += T.march_steps;
    }
    c->counters[F184_COUNTER_MARCH_STEPS] = total_steps;
    c->stage_ms[F184_STAGE_TRACE] = (float)(now_ms() - t0);
    return F184_OK;
}

// =================================================================================================
// GTAO
// =================================================================================================
namespace {
const float PI_ = 3.1415926f;
const float half_PI_ = 3.1415926f / 2.0f;

inline float fastSqrt(float x) { return dm_u2f((uint32_t)(0x1FBD1DF5 + ((int32_t)dm_f2u(x) >> 1))); }   // math.inc:14-16
inline float fastAcos1(float x)                                                                          // math.inc:22-26
{
    float res = -0.156583f * fabsf(x) + half_PI_;
    res *= fastSqrt(1.0f - fabsf(x));
    return x >= 0 ? res : PI_ - res;
}
inline float fastAcos2(float x)                                                                          // math.inc:28-33
{
    float res = -0.156583f * fabsf(x) + half_PI_;
    res *= fastSqrt(1.0f - fabsf(x));
    float flag = dm_step(x, 0.0f);
    return res * fmaf(-flag, 2.0f, 1.0f) + flag * PI_;
}
struct GtaoCtx { M4 InvProj; const float* depth; const uint16_t* normals; uint32_t W, H; };
inline V3 cs_pos(const GtaoCtx& G, float u, float v)      // gtao.frag:16-30
{
    int dx = dm_f2i(u * (float)G.W), dy = dm_f2i(v * (float)G.H);
    float depth = (dx >= 0 && dy >= 0 && dx < (int)G.W && dy < (int)G.H) ? G.depth[(size_t)dy * G.W + dx] : 0.0f;
    V4 p = mul(G.InvProj, V4{u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f});
    return {p.x / p.w, p.y / p.w, p.z / p.w};
}
inline float sgn(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
}  // namespace

extern "C" int f184o_gtao(f184o_ctx* c, const f184_view_constants* view)
{
    if (!c || !view) return F184_ERR_INVALID_ARGUMENT;
    double t0 = now_ms();
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_AO_RAW, F184_SLOT_AO_OUT})
    { int rc = ensure_image(c, s); if (rc) return rc; }
    GtaoCtx G{load_m4(view->InvProj), image_ptr<float>(c, F184_SLOT_DEPTH), image_ptr<uint16_t>(c, F184_SLOT_NORMALS), c->cfg.width, c->cfg.height};
    const uint32_t W = G.W, H = G.H;
    uint16_t* raw = image_ptr<uint16_t>(c, F184_SLOT_AO_RAW);
    uint16_t* out = image_ptr<uint16_t>(c, F184_SLOT_AO_OUT);
    const float cutoff = 32.0f;
    // gtao.frag:48-120
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++)
        {
            float u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;
            V3 cur = cs_pos(G, u, v);
            V3 Vv = neg(normalize(cur));
            float vis;
            if (-cur.z > cutoff) vis = 1.0f;
            else
            {
                const uint16_t* np = &G.normals[4 * ((size_t)y * W + x)];
                V3 cn = normalize(V3{fmaf(unorm16(np[0]), 2.0f, -1.0f), fmaf(unorm16(np[1]), 2.0f, -1.0f), fmaf(unorm16(np[2]), 2.0f, -1.0f)});
                float integral = 0.0f;
                float radius = (float)H * 0.5f / -cur.z;
                int cx = (int)x, cy = (int)y;
                float phi = -(1.0f / 16.0f) * (float)((((cx + cy) & 0x3) << 2) + (cx & 0x3)) * PI_;
                float rStep = radius / 2.0f;
                for (int samp = 0; samp < 4; samp++)
                {
                    float hx = -1.0f, hy = -1.0f;
                    float cph = O_COS(phi), sph = O_SIN(phi);
                    V3 sliceDir = {cph, sph, 0.0f};
                    float sdx = cph, sdy = -sph;
                    float r = rStep * (0.25f * (float)((cy - cx) & 0x3));
                    for (int j = 0; j < 2; j++)
                    {
                        float ox = r * sdx / (float)W, oy = r * sdy / (float)H;
                        r += rStep;
                        float u1 = u - ox, v1 = v - oy, u2 = u + ox, v2 = v + oy;
                        V3 ds = cs_pos(G, u1, v1) - cur, dt = cs_pos(G, u2, v2) - cur;
                        float hsx = dot(Vv, normalize(ds)), hsy = dot(Vv, normalize(dt));
                        if (dm_clamp(u1, 0.0f, 1.0f) != u1 || dm_clamp(v1, 0.0f, 1.0f) != v1) hsx = -1.0f;
                        if (dm_clamp(u2, 0.0f, 1.0f) != u2 || dm_clamp(v2, 0.0f, 1.0f) != v2) hsy = -1.0f;
                        float fx = dm_step(hsx, hx), fy = dm_step(hsy, hy);
                        hx = dm_mix(dm_mix(hx, hsx, 0.5f), dm_max(hx, hsx), fx);
                        hy = dm_mix(dm_mix(hy, hsy, 0.5f), dm_max(hy, hsy), fy);
                    }
                    hx = fastAcos2(hx); hy = fastAcos2(hy);
                    V3 sliceNormal = normalize(cross(Vv, sliceDir));
                    V3 sliceBitangent = normalize(cross(sliceNormal, Vv));
                    V3 projNorm = cn - sliceNormal * dot(cn, sliceNormal);
                    float weight = length(projNorm) + 1e-6f;
                    projNorm = projNorm / weight;
                    float cosn = dot(projNorm, Vv), sinn = dot(projNorm, sliceBitangent);
                    float n = fastAcos1(cosn) * sgn(sinn);
                    hx = n + dm_max(-hx - n, -half_PI_);
                    hy = n + dm_min(hy - n, half_PI_);
                    float ax = -O_COS(2.0f * hx - n) + O_COS(n) + 2.0f * hx * O_SIN(n);
                    float ay = -O_COS(2.0f * hy - n) + O_COS(n) + 2.0f * hy * O_SIN(n);
                    float a = 0.25f * (ax * 1.0f + ay * 1.0f);
                    integral += a * weight;
                    phi += PI_ / 4.0f;
                }
                vis = integral / 4.0f;
            }
            uint16_t* o = &raw[4 * ((size_t)y * W + x)];
            o[0] = dm_f32_to_f16(vis); o[1] = 0; o[2] = 0; o[3] = dm_f32_to_f16(1.0f);     // gtao.frag:125
        }
    // GTAO/blur.frag:12-27: four textureGatherOffset footprints = the 4x4 block x-1..x+2, y-1..y+2, wrap addressing.
    // PINNED: at a texel centre the gather footprint's base texel is the pixel itself.
#pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++)
        {
            auto wr = [](int i, int n) { int m = i % n; return m < 0 ? m + n : m; };
            auto R = [&](int xi, int yi) { return dm_f16_to_f32(raw[4 * ((size_t)wr(yi, (int)H) * W + wr(xi, (int)W))]); };
            const int offs[4][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
            float sum4[4];
            for (int g = 0; g < 4; g++)
            {
                int bx = (int)x + offs[g][0], by = (int)y + offs[g][1];
                // gather order: (i0,j1), (i1,j1), (i1,j0), (i0,j0)
                float gx = R(bx, by + 1), gy = R(bx + 1, by + 1), gz = R(bx + 1, by), gw = R(bx, by);
                sum4[g] = ((gx * 1.0f + gy * 1.0f) + gz * 1.0f) + gw * 1.0f;
            }
            float avg = (((sum4[0] * 1.0f + sum4[1] * 1.0f) + sum4[2] * 1.0f) + sum4[3] * 1.0f) / 16.0f;
            uint16_t hv = dm_f32_to_f16(avg);
            uint16_t* o = &out[4 * ((size_t)y * W + x)];
            o[0] = hv; o[1] = hv; o[2] = hv; o[3] = dm_f32_to_f16(1.0f);
        }
    c->stage_ms[F184_STAGE_GTAO] = (float)(now_ms() - t0);
    return F184_OK;
}

// =================================================================================================
// separable cross-bilateral blur of the indirect buffer (bilateralBlur.inc)
// =================================================================================================
namespace {
// All taps land on texel centres (offsets are 2*i and 2*i+1 pixels: `invres = 2/resolution`,
// bilateralBlur.inc:15) so the bilinear fetch degenerates to a texel read.  PINNED as such; sampler is
// GlobalLinearSamplerClamped (MegaPipeline.cpp:33-38) = clamp to edge.
void blur_pass(const uint16_t* src, const float* depth, uint16_t* dst, uint32_t W, uint32_t H, int dirx, int diry)
{
    const float BlurFalloff = 1.0f / (2.0f * 4.0f * 4.0f);
#pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++)
        {
            auto cl = [](int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); };
            auto tap = [&](int off, float* rgb, float* z) {
                int xi = cl((int)x + off * dirx, (int)W), yi = cl((int)y + off * diry, (int)H);
                const uint16_t* p = &src[4 * ((size_t)yi * W + xi)];
                rgb[0] = dm_f16_to_f32(p[0]); rgb[1] = dm_f16_to_f32(p[1]); rgb[2] = dm_f16_to_f32(p[2]);
                *z = depth[(size_t)yi * W + xi];
            };
            float rgb[3], z, cz;
            tap(0, rgb, &cz);
            float tc[3] = {rgb[0] * 1.0f, rgb[1] * 1.0f, rgb[2] * 1.0f};
            float tw = 1.0f;
            auto acc = [&](int off, float r) {
                tap(off, rgb, &z);
                float dz = (cz - z) * 512.0f;
                float w = O_EXP2(-r * r * BlurFalloff - dz * dz);
                tc[0] += rgb[0] * w; tc[1] += rgb[1] * w; tc[2] += rgb[2] * w;
                tw += w;
            };
            float i = 1.0f;
            for (; i <= 4.0f; i += 1.0f) { acc((int)(2.0f * i), i); acc(-(int)(2.0f * i), i); }
            for (; i <= 8.0f; i += 2.0f) { acc((int)(2.0f * (i + 0.5f)), i); acc(-(int)(2.0f * (0.5f + i)), i); }
            uint16_t* o = &dst[4 * ((size_t)y * W + x)];
            o[0] = dm_f32_to_f16(tc[0] / tw); o[1] = dm_f32_to_f16(tc[1] / tw); o[2] = dm_f32_to_f16(tc[2] / tw); o[3] = 0;
        }
}
}  // namespace

extern "C" int f184o_blur_indirect(f184o_ctx* c, const f184_engine_miscs* miscs)
{
    if (!c || !miscs) return F184_ERR_INVALID_ARGUMENT;
    double t0 = now_ms();
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_INDIRECT_OUT, F184_SLOT_INDIRECT_BLUR_X, F184_SLOT_INDIRECT_FINAL})
    { int rc = ensure_image(c, s); if (rc) return rc; }
    const uint32_t W = c->cfg.width, H = c->cfg.height;
    // The pass the reference NAMES indirect_blurX steps along y — blurX.frag:5 defines DIR(x) as vec2(0.0, x) — and
    // indirect_blurY steps along x (blurY.frag:5).  The bilateral weights do not commute, so the order is kept as
    // shipped: vertical first, then horizontal (found by running the reference's shader text, tests/test_refshader_pin.py).
    blur_pass(image_ptr<uint16_t>(c, F184_SLOT_INDIRECT_OUT), image_ptr<float>(c, F184_SLOT_DEPTH),
              image_ptr<uint16_t>(c, F184_SLOT_INDIRECT_BLUR_X), W, H, 0, 1);
    blur_pass(image_ptr<uint16_t>(c, F184_SLOT_INDIRECT_BLUR_X), image_ptr<float>(c, F184_SLOT_DEPTH),
              image_ptr<uint16_t>(c, F184_SLOT_INDIRECT_FINAL), W, H, 1, 0);
    c->stage_ms[F184_STAGE_BLUR] = (float)(now_ms() - t0);
    return F184_OK;
}

// =================================================================================================
// deferred direct lighting (Shader/Lighting/aggregateLights.frag; MegaPipeline.cpp:286-300)
// =================================================================================================
namespace {
inline float bw0(float a) { return (1.0f / 6.0f) * (a * (a * (-a + 3.0f) - 3.0f) + 1.0f); }      // math.inc:35-49
inline float bw1(float a) { return (1.0f / 6.0f) * (a * a * (3.0f * a - 6.0f) + 4.0f); }
inline float bw2(float a) { return (1.0f / 6.0f) * (a * (a * (-3.0f * a + 3.0f) + 3.0f) + 1.0f); }
inline float bw3(float a) { return (1.0f / 6.0f) * (a * a * a); }
inline float bg0(float a) { return bw0(a) + bw1(a); }                                              // math.inc:52-66
inline float bg1(float a) { return bw2(a) + bw3(a); }
inline float bh0(float a) { return -1.0f + bw1(a) / (bw0(a) + bw1(a)); }
inline float bh1(float a) { return 1.0f + bw3(a) / (bw2(a) + bw3(a)); }

struct LightCtx { const float* shadow; uint32_t S; };
inline float shadow_fetch(const LightCtx& L, float fx, float fy)
{
    const int x = dm_f2i(fx + 0.5f), y = dm_f2i(fy + 0.5f);
    return (x >= 0 && y >= 0 && x < (int)L.S && y < (int)L.S) ? L.shadow[(size_t)y * L.S + x] : 0.0f;   // PINNED: out of range -> 0
}
// shadowTexSmooth, aggregateLights.frag:83-117
inline float shadow_smooth(const LightCtx& L, float sx, float sy, float sz, float bias)
{
    const float res = (float)L.S;
    const float ux = sx * res - 1.0f, uy = sy * res - 1.0f;
    const float ix = floorf(ux), iy = floorf(uy);
    const float fx = ux - ix, fy = uy - iy;
    const float g0x = bg0(fx), g1x = bg1(fx);
    const float h0x = bh0(fx) * 0.75f, h1x = bh1(fx) * 0.75f, h0y = bh0(fy) * 0.75f, h1y = bh1(fy) * 0.75f;
    const float r0 = dm_step(sz, shadow_fetch(L, ix + h0x, iy + h0y) + bias);
    const float r1 = dm_step(sz, shadow_fetch(L, ix + h1x, iy + h0y) + bias);
    const float r2 = dm_step(sz, shadow_fetch(L, ix + h0x, iy + h1y) + bias);
    const float r3 = dm_step(sz, shadow_fetch(L, ix + h1x, iy + h1y) + bias);
    return bg0(fy) * (g0x * r0 + g1x * r1) + bg1(fy) * (g0x * r2 + g1x * r3);
}
inline float ggx_schlick(float NdotV, float roughness)          // :156-165
{
    const float r = roughness + 1.0f;
    const float k = r * r / 8.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
}
// illumination, aggregateLights.frag:176-205
inline V3 illumination(V3 lightVector, const float* lum, V3 cspos, V3 csnorm, float metallicity, float roughness)
{
    const V3 wi = normalize(lightVector), wo = normalize(neg(cspos)), halfvec = normalize(wi + wo);
    const float distSq = dot(lightVector, lightVector);
    const float dist = fastSqrt(distSq);
    const V3 radiance = {lum[0] / distSq, lum[1] / distSq, lum[2] / distSq};
    float r4 = roughness; r4 *= r4; r4 *= r4;                    // NDF :143-154
    float cTheta = dm_max(dot(csnorm, halfvec), 0.0f);
    cTheta *= cTheta;
    float nd = cTheta * (r4 - 1.0f) + 1.0f;
    nd *= nd;
    const float normalDist = r4 / (PI_ * nd);
    const float NdotV = dm_max(dot(csnorm, wo), 0.0f), NdotL0 = dm_max(dot(csnorm, wi), 0.0f);        // G :167-174
    const float g = ggx_schlick(NdotL0, roughness) * ggx_schlick(NdotV, roughness);
    const float cosTheta = dm_max(dot(wo, halfvec), 0.0f);       // metallicFresnel :134-141 (its albedo term is unused as shipped)
    const float fresnel = 0.04f + 0.96f * O_POW(1.0f - cosTheta, 5.0f);
    const float num = normalDist * g * fresnel;
    const float denom = 4.0f * dm_max(dot(csnorm, wo), 0.0f) * dm_max(dot(csnorm, wi), 0.0f);
    const float specular = num / dm_max(denom, 0.001f);
    float diffuse = 1.0f - fresnel;
    diffuse *= 1.0f - metallicity;
    const V3 wid = {wi.x / dist, wi.y / dist, wi.z / dist};
    const float NdotL = dm_max(dot(csnorm, wid), 0.0f);
    const float ds = diffuse + specular;
    return {ds * radiance.x * NdotL, ds * radiance.y * NdotL, ds * radiance.z * NdotL};
}
const float kPoisson12[12][2] = {
    {-0.326212f, -0.40581f}, {-0.840144f, -0.07358f}, {-0.695914f, 0.457137f}, {-0.203345f, 0.620716f},
    {0.96234f, -0.194983f},  {0.473434f, -0.480026f}, {0.519456f, 0.767022f},  {0.185461f, -0.893124f},
    {0.507431f, 0.064425f},  {0.89642f, 0.412458f},   {-0.32194f, -0.932615f}, {-0.791559f, -0.59771f}};   // :119-132
}  // namespace

extern "C" int f184o_lighting_deferred(f184o_ctx* c, const f184_view_constants* view, const f184_extended_matrices* m,
                                       const f184_light_list* point, const f184_light_list* directional)
{
    if (!c || !view || !m) return F184_ERR_INVALID_ARGUMENT;
    double t0 = now_ms();
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_MATERIAL, F184_SLOT_SHADOW, F184_SLOT_LIGHTING})
    { int rc = ensure_image(c, s); if (rc) return rc; }
    const int np = point ? point->numLights : 0, nd = directional ? directional->numLights : 0;
    if (np < 0 || np > 100 || nd < 0 || nd > 100) return F184_ERR_INVALID_ARGUMENT;
    const M4 InvProj = load_m4(view->InvProj), ViewMat = load_m4(view->ViewMat), InvModelView = load_m4(m->InvModelView);
    const M4 ShadowView = load_m4(m->ShadowView), ShadowProj = load_m4(m->ShadowProj);
    const uint32_t W = c->cfg.width, H = c->cfg.height;
    const float* depthp = image_ptr<float>(c, F184_SLOT_DEPTH);
    const uint16_t* normals = image_ptr<uint16_t>(c, F184_SLOT_NORMALS);
    const uint8_t* material = image_ptr<uint8_t>(c, F184_SLOT_MATERIAL);
    uint16_t* out = image_ptr<uint16_t>(c, F184_SLOT_LIGHTING);
    const LightCtx LC{image_ptr<float>(c, F184_SLOT_SHADOW), c->cfg.shadow_res};
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++)
        {
            const float u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;
            const int dx = dm_f2i(u * (float)W), dy = dm_f2i(v * (float)H);
            const float depth = (dx >= 0 && dy >= 0 && dx < (int)W && dy < (int)H) ? depthp[(size_t)dy * W + dx] : 0.0f;
            const V4 cp = mul(InvProj, V4{u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f});
            const V3 cspos = {cp.x / cp.w, cp.y / cp.w, cp.z / cp.w};
            // PINNED: texture() at a texel centre returns that texel (normals, material)
            const uint16_t* nq = &normals[4 * ((size_t)y * W + x)];
            const V3 csnorm = normalize(V3{fmaf(unorm16(nq[0]), 2.0f, -1.0f), fmaf(unorm16(nq[1]), 2.0f, -1.0f), fmaf(unorm16(nq[2]), 2.0f, -1.0f)});
            const uint8_t* mq = &material[4 * ((size_t)y * W + x)];
            const float roughness = (float)mq[1] / 255.0f, metallicity = (float)mq[2] / 255.0f;      // getMaterial = .yz, :60-70
            V3 result = {0, 0, 0};
            for (int i = 0; i < np; i++)
            {
                const f184_sun& L = point->lights[i];
                const V4 lp = mul(ViewMat, V4{L.position[0], L.position[1], L.position[2], 1.0f});
                result = result + illumination(V3{lp.x, lp.y, lp.z} - cspos, L.luminance, cspos, csnorm, metallicity, roughness);
            }
            for (int i = 0; i < nd; i++)
            {
                const f184_sun& L = directional->lights[i];
                const V3 lightVector = neg(normalize(mul3(ViewMat, V3{L.position[0], L.position[1], L.position[2]})));
                const V4 wp = mul(InvModelView, V4{cspos.x + csnorm.x * 0.01f, cspos.y + csnorm.y * 0.01f, cspos.z + csnorm.z * 0.01f, 1.0f});
                V4 spos = mul(ShadowProj, mul(ShadowView, V4{wp.x, wp.y, wp.z, 1.0f}));                 // no divide by w, :227-229
                spos.x = spos.x * 0.5f + 0.5f; spos.y = spos.y * 0.5f + 0.5f;
                const float pix = 1.0f / (float)LC.S;
                float shade = 0.0f;
                for (int j = 0; j < 12; j++)
                    shade += shadow_smooth(LC, spos.x + kPoisson12[j][0] * pix, spos.y + kPoisson12[j][1] * pix, spos.z + 0.0f, 0.002f);
                shade /= 12.0f;
                const V3 il = illumination(lightVector, L.luminance, cspos, csnorm, metallicity, roughness);
                result = {result.x + il.x * shade, result.y + il.y * shade, result.z + il.z * shade};
            }
            uint16_t* o = &out[4 * ((size_t)y * W + x)];
            o[0] = dm_f32_to_f16(result.x); o[1] = dm_f32_to_f16(result.y); o[2] = dm_f32_to_f16(result.z); o[3] = dm_f32_to_f16(1.0f);
        }
    (void)t0;
    return F184_OK;
}

// =================================================================================================
// composite (Shader/GTAO/color.frag; MegaPipeline.cpp:302-319): the per-pixel function is
// final184_b200/csrc/f184_composite.h, shared with the CUDA kernel the way f184_detmath.h is; what pins it is the
// reference's shader text (tests/test_refshader_pin.py), not this wrapper.
// =================================================================================================
#include "../final184_b200/csrc/f184_composite.h"

extern "C" int f184o_composite(f184o_ctx* c, const f184_trace_constants* k)
{
    if (!c || !k) return F184_ERR_INVALID_ARGUMENT;
    for (int s : {F184_SLOT_ALBEDO, F184_SLOT_AO_OUT, F184_SLOT_DEPTH, F184_SLOT_LIGHTING, F184_SLOT_SHADOW, F184_SLOT_INDIRECT_FINAL,
                  F184_SLOT_TAA_HISTORY, F184_SLOT_TAA_OUT, F184_SLOT_COLOR_OUT})
    { int rc = ensure_image(c, s); if (rc) return rc; }
    f184_composite_in I{};
    memcpy(I.InvProj.m, k->view.InvProj, 64); memcpy(I.InvModelView.m, k->ext.InvModelView, 64);
    memcpy(I.ShadowView.m, k->ext.ShadowView, 64); memcpy(I.ShadowProj.m, k->ext.ShadowProj, 64);
    memcpy(I.prevModelView.m, k->prev.PrevModelView, 64); memcpy(I.prevProjection.m, k->prev.PrevProjection, 64);
    memcpy(I.sun_luminance, k->sun.luminance, 12); memcpy(I.sun_position, k->sun.position, 12);
    I.albedo = image_ptr<uint8_t>(c, F184_SLOT_ALBEDO); I.ao = image_ptr<uint16_t>(c, F184_SLOT_AO_OUT);
    I.lighting = image_ptr<uint16_t>(c, F184_SLOT_LIGHTING); I.indirect = image_ptr<uint16_t>(c, F184_SLOT_INDIRECT_FINAL);
    I.taa = image_ptr<uint16_t>(c, F184_SLOT_TAA_HISTORY); I.depth = image_ptr<float>(c, F184_SLOT_DEPTH); I.shadow = image_ptr<float>(c, F184_SLOT_SHADOW);
    I.W = c->cfg.width; I.H = c->cfg.height; I.S = c->cfg.shadow_res;
    if (k->reset_history) memset(c->img[F184_SLOT_TAA_HISTORY].ptr, 0, c->img[F184_SLOT_TAA_HISTORY].desc.size_bytes);
    uint16_t* oc = image_ptr<uint16_t>(c, F184_SLOT_COLOR_OUT);
    uint16_t* ot = image_ptr<uint16_t>(c, F184_SLOT_TAA_OUT);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < (int64_t)I.H; y++)
        for (uint32_t x = 0; x < I.W; x++)
            f184_composite_pixel(I, x, (uint32_t)y, &oc[4 * ((size_t)y * I.W + x)], &ot[4 * ((size_t)y * I.W + x)]);
    return F184_OK;
}
