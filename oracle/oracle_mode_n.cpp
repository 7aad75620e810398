// oracle_mode_n.cpp — TEST INFRASTRUCTURE (see oracle_common.h).
// CPU definition of the north-star stages ("mode N"): conservative voxelization into sum+count
// accumulators, normalise, light injection, six-direction anisotropic mips, diffuse+specular cone tracing.
// The reference ships NONE of these: there is no reference code to restate, so this file
// follows the normative spec in DESIGN.md §"Mode N" (derived from BASELINE.json:north_star and SURVEY.md
// Appendix B) and reuses the reference's constants where one exists:
//   alpha discard 0.05, MaxLod 4 sampler      Pipelang/Internal/main.lua:199, MegaPipeline.cpp:39-43
//   first-bounce lighting / shadow test        Shader/Lighting/indirect.frag:157-169
//   sky term, tangent frame, temporal blend    indirect.frag:180, 71-86, 225-240
//   Fresnel                                    Shader/Lighting/aggregateLights.frag:135-141
// Parity status: unpinned by construction (nothing in the reference to pin against).
#include <algorithm>
#include <cstdio>

#include "oracle_common.h"

using namespace orc;

namespace {

// ---- texture sampling: mode R's sampler geometry (wrap addressing, LOD from the uv derivatives, MaxLod 4), but
// filtered in 8-BIT UNITS: texel bytes enter the lerps as 0..255 floats and the result stays on that scale (the
// caller quantises to 8 bits anyway), which keeps 32 divisions by 255 out of every fragment on the device.
inline int wrapi(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

V4 bilinear_level(const Texture& t, uint32_t level, float u, float v)
{
    uint32_t w = std::max(1u, t.w >> level), h = std::max(1u, t.h >> level);
    const uint8_t* px = t.levels[level].data();
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = wrapi(dm_f2i(x0f), (int)w), y0 = wrapi(dm_f2i(y0f), (int)h);
    int x1 = wrapi(x0 + 1, (int)w), y1 = wrapi(y0 + 1, (int)h);
    float r[4];
    for (int c = 0; c < 4; c++)
    {
        float a = (float)px[4 * ((size_t)y0 * w + x0) + c], b = (float)px[4 * ((size_t)y0 * w + x1) + c];
        float cc = (float)px[4 * ((size_t)y1 * w + x0) + c], d = (float)px[4 * ((size_t)y1 * w + x1) + c];
        float top = a * (1.0f - fx) + b * fx, bot = cc * (1.0f - fx) + d * fx;
        r[c] = top * (1.0f - fy) + bot * fy;
    }
    return {r[0], r[1], r[2], r[3]};
}

V4 sample_trilinear(const Texture& t, float u, float v, float dudx, float dvdx, float dudy, float dvdy)
{
    float ax = dudx * (float)t.w, ay = dvdx * (float)t.h, bx = dudy * (float)t.w, by = dvdy * (float)t.h;
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
    return {c0.x * (1.0f - f) + c1.x * f, c0.y * (1.0f - f) + c1.y * f, c0.z * (1.0f - f) + c1.z * f, c0.w * (1.0f - f) + c1.w * f};
}

inline int64_t floor_div256(int64_t a) { return a >= 0 ? a / 256 : -((-a + 255) / 256); }

// Exact triangle / closed-box overlap (Akenine-Moller SAT) on the 1/256-voxel integer lattice.
// Box = [256*i, 256*(i+1)] per axis; v = triangle vertices; n = (v1-v0) x (v2-v0).
bool tri_box_overlap(const int64_t v[3][3], const int64_t n[3], const int64_t box[3])
{
    int64_t p[3][3];                       // vertices relative to the box centre
    for (int k = 0; k < 3; k++)
        for (int a = 0; a < 3; a++) p[k][a] = v[k][a] - (256 * box[a] + 128);
    const int64_t hs = 128;
    for (int a = 0; a < 3; a++)            // box axes
    {
        int64_t mn = std::min({p[0][a], p[1][a], p[2][a]}), mx = std::max({p[0][a], p[1][a], p[2][a]});
        if (mn > hs || mx < -hs) return false;
    }
    {                                       // triangle plane
        int64_t d = n[0] * p[0][0] + n[1] * p[0][1] + n[2] * p[0][2];
        int64_t r = hs * (std::llabs(n[0]) + std::llabs(n[1]) + std::llabs(n[2]));
        if (d > r || d < -r) return false;
    }
    for (int e = 0; e < 3; e++)            // 9 edge x axis cross products
    {
        const int64_t ex = p[(e + 1) % 3][0] - p[e][0], ey = p[(e + 1) % 3][1] - p[e][1], ez = p[(e + 1) % 3][2] - p[e][2];
        const int64_t ax[3][3] = {{0, -ez, ey}, {ez, 0, -ex}, {-ey, ex, 0}};
        for (int a = 0; a < 3; a++)
        {
            int64_t q0 = ax[a][0] * p[0][0] + ax[a][1] * p[0][1] + ax[a][2] * p[0][2];
            int64_t q1 = ax[a][0] * p[1][0] + ax[a][1] * p[1][1] + ax[a][2] * p[1][2];
            int64_t q2 = ax[a][0] * p[2][0] + ax[a][1] * p[2][1] + ax[a][2] * p[2][2];
            int64_t r = hs * (std::llabs(ax[a][0]) + std::llabs(ax[a][1]) + std::llabs(ax[a][2]));
            int64_t mn = std::min({q0, q1, q2}), mx = std::max({q0, q1, q2});
            if (mn > r || mx < -r) return false;
        }
    }
    return true;
}

struct VtxN { V3 n; float u, v; };

// RGBA8 radiance stores radiance / exposure.  exposure = config value, or (0 = auto) the largest sun
// luminance component: the injected radiance is albedo^2.2 * luminance * |cos| * shade <= that.
float exposure_of(const f184o_ctx* c, const f184_sun* sun)
{
    if (c->cfg.radiance_exposure > 0.0f) return c->cfg.radiance_exposure;
    float m = std::max(sun->luminance[0], std::max(sun->luminance[1], sun->luminance[2]));
    return m > 0.0f ? m : 1.0f;
}

// double-precision 4x4 inverse (Gauss-Jordan, partial pivoting) of a float matrix in upload order
M4 invert(const M4& A)
{
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) { a[r][c] = A.m[c * 4 + r]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
    for (int col = 0; col < 4; col++)
    {
        int piv = col;
        for (int r = col + 1; r < 4; r++) if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
        if (piv != col) for (int c = 0; c < 8; c++) std::swap(a[piv][c], a[col][c]);
        double d = a[col][col];
        for (int c = 0; c < 8; c++) a[col][c] /= d;
        for (int r = 0; r < 4; r++)
            if (r != col)
            {
                double f = a[r][col];
                if (f != 0.0) for (int c = 0; c < 8; c++) a[r][c] -= f * a[col][c];
            }
    }
    M4 R;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) R.m[c * 4 + r] = (float)a[r][4 + c];
    return R;
}

}  // namespace

// =================================================================================================
// B.1 conservative voxelization + B.2 normalise
// =================================================================================================
extern "C" int orc_normalise_n(f184o_ctx* c);
extern "C" int orc_voxelize_accumulate_n(f184o_ctx* c, const f184_view_constants* cam);

extern "C" int orc_voxelize_n(f184o_ctx* c, const f184_view_constants* cam)
{
    int rc = orc_voxelize_accumulate_n(c, cam);
    return rc ? rc : orc_normalise_n(c);
}

// B.1: accumulate the fragments of triangles [tri_first, tri_first + tri_count) into ZEROED accumulators (a partial
// volume: the multi-GPU schedule sums these across ranks before normalising).
extern "C" int orc_voxelize_accumulate_n(f184o_ctx* c, const f184_view_constants* cam)
{
    double t0 = now_ms();
    for (int s : {F184_SLOT_ACCUM_COLOR, F184_SLOT_ACCUM_NORMAL, F184_SLOT_VOX_ALBEDO, F184_SLOT_VOX_NORMAL})
    { int rc = ensure_image(c, s); if (rc) return rc; }
    const uint32_t N = c->cfg.grid_n;
    const size_t nvox = (size_t)N * N * N;
    float* accC = image_ptr<float>(c, F184_SLOT_ACCUM_COLOR);
    float* accN = image_ptr<float>(c, F184_SLOT_ACCUM_NORMAL);
    memset(accC, 0, nvox * 16);
    memset(accN, 0, nvox * 16);
    memcpy(c->last_vox_cam, cam->ViewMat, 64);
    memcpy(c->last_vox_cam + 16, cam->ProjMat, 64);
    if (c->cache_held && memcmp(c->last_vox_cam, c->cache_cam, sizeof(c->cache_cam)) != 0)
    { c->err = "voxelize: the static cache was captured with another voxel camera"; return F184_ERR_INVALID_ARGUMENT; }

    const M4 View = load_m4(cam->ViewMat), Proj = load_m4(cam->ProjMat);
    std::vector<M4> VM(c->n_models), MM(c->n_models);
    for (uint32_t m = 0; m < c->n_models; m++) { MM[m] = load_m4(&c->model_mats[16 * m]); VM[m] = matmul(View, MM[m]); }
    const float Nf = (float)N;
    uint64_t frags = 0;
    const uint32_t first = c->tri_first, last = (uint32_t)std::min<uint64_t>((uint64_t)c->tri_first + c->tri_count, c->n_tris);

    // Triangles in parallel: every addend is an integer-valued float and the per-voxel sums stay below 2^24, so
    // the atomic adds are exact and the result does not depend on the order (same argument as on the device).
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : frags)
    for (uint32_t t = first; t < last; t++)
    {
        if (!c->chunk_mask.empty() && (t / F184_TRIANGLE_CHUNK >= c->chunk_mask.size() || !c->chunk_mask[t / F184_TRIANGLE_CHUNK])) continue;
        const uint32_t* id = &c->idx[3 * t];
        const M4& vm = VM[c->tri_model[t]];
        const M4& mm = MM[c->tri_model[t]];
        int64_t v[3][3];
        VtxN at[3];
        bool bad = false;
        for (int i = 0; i < 3; i++)
        {
            const float* p = &c->pos[3 * id[i]];
            V4 q = mul(vm, V4{p[0], p[1], p[2], 1.0f});
            V4 g = mul(Proj, V4{q.x, q.y, q.z, 1.0f});
            float gx = g.x / g.w, gy = g.y / g.w, gz = g.z / g.w;
            // voxel space: the tracer's mapping for every axis (indirect.frag:141-143)
            float vx = (gx * 0.5f + 0.5f) * Nf, vy = (gy * 0.5f + 0.5f) * Nf, vz = gz * Nf;
            // guard band: vertices outside [-N, 2N) drop the triangle (keeps the integer SAT inside int64)
            if (!(vx >= -Nf && vx < 2.0f * Nf) || !(vy >= -Nf && vy < 2.0f * Nf) || !(vz >= -Nf && vz < 2.0f * Nf)) { bad = true; break; }
            v[i][0] = (int64_t)rintf(vx * 256.0f); v[i][1] = (int64_t)rintf(vy * 256.0f); v[i][2] = (int64_t)rintf(vz * 256.0f);
            const float* nn = &c->nrm[3 * id[i]];
            at[i].n = normalize(mul3(mm, V3{nn[0], nn[1], nn[2]}));
            at[i].u = c->uv[2 * id[i]]; at[i].v = c->uv[2 * id[i] + 1];
        }
        if (bad) continue;
        int64_t e1[3], e2[3], n[3];
        for (int a = 0; a < 3; a++) { e1[a] = v[1][a] - v[0][a]; e2[a] = v[2][a] - v[0][a]; }
        n[0] = e1[1] * e2[2] - e1[2] * e2[1];
        n[1] = e1[2] * e2[0] - e1[0] * e2[2];
        n[2] = e1[0] * e2[1] - e1[1] * e2[0];
        if (n[0] == 0 && n[1] == 0 && n[2] == 0) continue;            // zero-area after snapping
        const int64_t anx = std::llabs(n[0]), any = std::llabs(n[1]), anz = std::llabs(n[2]);
        int d;
        if (anx > any) d = (anx > anz) ? 0 : 2;
        else d = (any > anz) ? 1 : 2;
        const int ua = (d + 1) % 3, va = (d + 2) % 3;
        // closed-box candidate range per axis
        int64_t lo[3], hi[3];
        bool empty = false;
        for (int a = 0; a < 3; a++)
        {
            int64_t mn = std::min({v[0][a], v[1][a], v[2][a]}), mx = std::max({v[0][a], v[1][a], v[2][a]});
            lo[a] = std::max<int64_t>(0, floor_div256(mn - 1));
            hi[a] = std::min<int64_t>(N - 1, floor_div256(mx));
            if (lo[a] > hi[a]) empty = true;
        }
        if (empty) continue;
        // 2D set-up in the dominant-axis projection: area = n[d] (cyclic axes), made positive
        int64_t area = n[d];
        const int64_t sgn = area < 0 ? -1 : 1;
        area *= sgn;
        int64_t eu[3], ev[3];                  // edge k: vertex (k+1) -> (k+2), sign-normalised
        for (int k = 0; k < 3; k++)
        {
            const int a = (k + 1) % 3, b = (k + 2) % 3;
            eu[k] = sgn * (v[b][ua] - v[a][ua]); ev[k] = sgn * (v[b][va] - v[a][va]);
        }
        const float areaf = (float)area;
        float dbdu[3], dbdv[3];
        for (int k = 0; k < 3; k++) { dbdu[k] = (float)(-ev[k] * 256) / areaf; dbdv[k] = (float)(eu[k] * 256) / areaf; }
        const float dudx = (at[0].u * dbdu[0] + at[1].u * dbdu[1]) + at[2].u * dbdu[2];
        const float dvdx = (at[0].v * dbdu[0] + at[1].v * dbdu[1]) + at[2].v * dbdu[2];
        const float dudy = (at[0].u * dbdv[0] + at[1].u * dbdv[1]) + at[2].u * dbdv[2];
        const float dvdy = (at[0].v * dbdv[0] + at[1].v * dbdv[1]) + at[2].v * dbdv[2];
        const Material& mat = c->materials[c->tri_mat[t]];
        const Texture* tex = (mat.use_textures && mat.tex >= 0) ? &c->textures[mat.tex] : nullptr;

        int64_t box[3];
        for (box[2] = lo[2]; box[2] <= hi[2]; box[2]++)
            for (box[1] = lo[1]; box[1] <= hi[1]; box[1]++)
                for (box[0] = lo[0]; box[0] <= hi[0]; box[0]++)
                {
                    if (!tri_box_overlap(v, n, box)) continue;
                    // barycentrics of the voxel centre in the projection, clamped into the triangle
                    const int64_t cu = 256 * box[ua] + 128, cv = 256 * box[va] + 128;
                    float b[3];
                    for (int k = 0; k < 3; k++)
                    {
                        const int a = (k + 1) % 3;
                        int64_t w = eu[k] * (cv - v[a][va]) - ev[k] * (cu - v[a][ua]);
                        b[k] = (float)w / areaf;
                        if (b[k] < 0.0f) b[k] = 0.0f;
                    }
                    const float s = (b[0] + b[1]) + b[2];
                    b[0] = b[0] / s; b[1] = b[1] / s; b[2] = b[2] / s;
                    const float u = (at[0].u * b[0] + at[1].u * b[1]) + at[2].u * b[2];
                    const float vv = (at[0].v * b[0] + at[1].v * b[1]) + at[2].v * b[2];
                    V3 nn = {(at[0].n.x * b[0] + at[1].n.x * b[1]) + at[2].n.x * b[2], (at[0].n.y * b[0] + at[1].n.y * b[1]) + at[2].n.y * b[2],
                             (at[0].n.z * b[0] + at[1].n.z * b[1]) + at[2].n.z * b[2]};
                    V4 base;
                    // base colour in 8-bit units (0..255)
                    if (!mat.use_textures) base = {mat.factor[0] * 255.0f, mat.factor[1] * 255.0f, mat.factor[2] * 255.0f, mat.factor[3] * 255.0f};
                    else
                    {
                        V4 sc = tex ? sample_trilinear(*tex, u, vv, dudx, dvdx, dudy, dvdy) : V4{0, 0, 0, 0};
                        base = {sc.x * mat.factor[0], sc.y * mat.factor[1], sc.z * mat.factor[2], sc.w * mat.factor[3]};
                        if (base.w < 12.75f) continue;                  // main.lua:199, alpha < 0.05
                    }
                    // two-sided lighting downstream (indirect.frag:60-62): fold the normal into one hemisphere so
                    // the two faces of a thin wall add up instead of cancelling
                    const float fx = fabsf(nn.x), fy = fabsf(nn.y), fz = fabsf(nn.z);
                    float lead = (fx >= fy && fx >= fz) ? nn.x : ((fy >= fz) ? nn.y : nn.z);
                    if (lead < 0.0f) nn = neg(nn);
                    // quantise so the fp32 sums are exact integers (order-independent, exact across GPUs)
                    const float r8 = floorf(dm_clamp(base.x, 0.0f, 255.0f) + 0.5f);
                    const float g8 = floorf(dm_clamp(base.y, 0.0f, 255.0f) + 0.5f);
                    const float b8 = floorf(dm_clamp(base.z, 0.0f, 255.0f) + 0.5f);
                    const float nx8 = rintf(dm_clamp(nn.x, -1.0f, 1.0f) * 127.0f), ny8 = rintf(dm_clamp(nn.y, -1.0f, 1.0f) * 127.0f),
                                nz8 = rintf(dm_clamp(nn.z, -1.0f, 1.0f) * 127.0f);
                    const size_t o = ((size_t)box[2] * N + box[1]) * N + box[0];
                    const float add[7] = {r8, g8, b8, 1.0f, nx8, ny8, nz8};
                    for (int q = 0; q < 4; q++)
                    {
#pragma omp atomic
                        accC[4 * o + q] += add[q];
                    }
                    for (int q = 0; q < 3; q++)
                    {
#pragma omp atomic
                        accN[4 * o + q] += add[4 + q];
                    }
                    frags++;
                }
    }
    c->counters[F184_COUNTER_FRAGMENTS] = frags;
    c->stage_ms[F184_STAGE_VOXELIZE] = (float)(now_ms() - t0);
    return F184_OK;
}

// B.2 normalise (reads the accumulators as they are: summed partial volumes included)
extern "C" int orc_normalise_n(f184o_ctx* c)
{
    for (int s : {F184_SLOT_ACCUM_COLOR, F184_SLOT_ACCUM_NORMAL, F184_SLOT_VOX_ALBEDO, F184_SLOT_VOX_NORMAL})
    { int rc = ensure_image(c, s); if (rc) return rc; }
    const uint32_t N = c->cfg.grid_n;
    const size_t nvox = (size_t)N * N * N;
    float* accC = image_ptr<float>(c, F184_SLOT_ACCUM_COLOR);
    float* accN = image_ptr<float>(c, F184_SLOT_ACCUM_NORMAL);
    // static cache (include/f184.h "static / dynamic split"): the sums normalise sees are the frame's plus the cached static ones —
    // integer-valued floats, exact in any order
    std::vector<float> sumC, sumN;
    if (c->cache_held)
    {
        sumC.resize(nvox * 4); sumN.resize(nvox * 4);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < nvox * 4; i++) { sumC[i] = accC[i] + c->cacheC[i]; sumN[i] = accN[i] + c->cacheN[i]; }
        accC = sumC.data(); accN = sumN.data();
    }
    double t1 = now_ms();
    uint8_t* alb = image_ptr<uint8_t>(c, F184_SLOT_VOX_ALBEDO);
    int8_t* nrm = image_ptr<int8_t>(c, F184_SLOT_VOX_NORMAL);
    uint64_t occ = 0;
    std::vector<uint8_t> brick((size_t)(N / 8) * (N / 8) * (N / 8), 0);
#pragma omp parallel for schedule(static) reduction(+ : occ)
    for (size_t o = 0; o < nvox; o++)
    {
        const float cnt = accC[4 * o + 3];
        if (cnt > 0.0f)
        {
            alb[4 * o] = (uint8_t)floorf(accC[4 * o] / cnt + 0.5f);
            alb[4 * o + 1] = (uint8_t)floorf(accC[4 * o + 1] / cnt + 0.5f);
            alb[4 * o + 2] = (uint8_t)floorf(accC[4 * o + 2] / cnt + 0.5f);
            alb[4 * o + 3] = 255;
            const float sx = accN[4 * o], sy = accN[4 * o + 1], sz = accN[4 * o + 2];
            const float len = sqrtf((sx * sx + sy * sy) + sz * sz);
            if (len > 0.0f)
            {
                nrm[4 * o] = (int8_t)rintf(sx / len * 127.0f); nrm[4 * o + 1] = (int8_t)rintf(sy / len * 127.0f); nrm[4 * o + 2] = (int8_t)rintf(sz / len * 127.0f);
            }
            else { nrm[4 * o] = nrm[4 * o + 1] = nrm[4 * o + 2] = 0; }
            nrm[4 * o + 3] = 0;
            occ++;
            size_t x = o % N, y = (o / N) % N, z = o / ((size_t)N * N);
            uint8_t& flag = brick[((z / 8) * (N / 8) + (y / 8)) * (N / 8) + (x / 8)];
#pragma omp atomic write
            flag = 1;
        }
        else { memset(&alb[4 * o], 0, 4); memset(&nrm[4 * o], 0, 4); }
    }
    uint64_t nb = 0;
    for (uint8_t b : brick) nb += b;
    c->counters[F184_COUNTER_OCCUPIED] = occ;
    c->counters[F184_COUNTER_BRICKS] = nb;
    c->stage_ms[F184_STAGE_NORMALISE] = (float)(now_ms() - t1);
    return F184_OK;
}

// static / dynamic split: keep the accumulators as the static cache (dense here; sparse on the device), leave them clear
extern "C" int orc_static_cache_capture_n(f184o_ctx* c)
{
    for (int s : {F184_SLOT_ACCUM_COLOR, F184_SLOT_ACCUM_NORMAL})
    { int rc = ensure_image(c, s); if (rc) return rc; }
    const size_t nvox = (size_t)c->cfg.grid_n * c->cfg.grid_n * c->cfg.grid_n;
    float* accC = image_ptr<float>(c, F184_SLOT_ACCUM_COLOR);
    float* accN = image_ptr<float>(c, F184_SLOT_ACCUM_NORMAL);
    c->cacheC.assign(accC, accC + nvox * 4);
    c->cacheN.assign(accN, accN + nvox * 4);
    memset(accC, 0, nvox * 16);
    memset(accN, 0, nvox * 16);
    c->cache_fragments = c->counters[F184_COUNTER_FRAGMENTS];
    memcpy(c->cache_cam, c->last_vox_cam, sizeof(c->cache_cam));
    c->cache_held = true;
    return F184_OK;
}

// =================================================================================================
// B.3 light injection
// =================================================================================================
extern "C" int orc_inject_n(f184o_ctx* c, const f184_sun* sun, const f184_extended_matrices* m)
{
    double t0 = now_ms();
    for (int s : {F184_SLOT_VOX_ALBEDO, F184_SLOT_VOX_NORMAL, F184_SLOT_RADIANCE, F184_SLOT_SHADOW})
    { int rc = ensure_image(c, s); if (rc) return rc; }
    const uint32_t N = c->cfg.grid_n, S = c->cfg.shadow_res;
    const uint8_t* alb = image_ptr<uint8_t>(c, F184_SLOT_VOX_ALBEDO);
    const int8_t* nrm = image_ptr<int8_t>(c, F184_SLOT_VOX_NORMAL);
    const float* shadow = image_ptr<float>(c, F184_SLOT_SHADOW);
    uint8_t* rad = image_ptr<uint8_t>(c, F184_SLOT_RADIANCE);
    const M4 w2v = matmul(load_m4(m->VoxelProj), load_m4(m->VoxelView));
    const M4 v2w = invert(w2v);
    const M4 ShadowView = load_m4(m->ShadowView), ShadowProj = load_m4(m->ShadowProj);
    const V3 sunPos = {sun->position[0], sun->position[1], sun->position[2]};
    const float Nf = (float)N, inv_exposure = 1.0f / exposure_of(c, sun);
#pragma omp parallel for schedule(static)
    for (int64_t z = 0; z < (int64_t)N; z++)
        for (uint32_t y = 0; y < N; y++)
            for (uint32_t x = 0; x < N; x++)
            {
                const size_t o = ((size_t)z * N + y) * N + x;
                if (alb[4 * o + 3] == 0) { memset(&rad[4 * o], 0, 4); continue; }
                // voxel centre -> world
                const float nx = ((float)x + 0.5f) / Nf * 2.0f - 1.0f, ny = ((float)y + 0.5f) / Nf * 2.0f - 1.0f, nz = ((float)z + 0.5f) / Nf;
                V4 p4 = mul(v2w, V4{nx, ny, nz, 1.0f});
                V3 p = {p4.x / p4.w, p4.y / p4.w, p4.z / p4.w};
                V3 n = {(float)nrm[4 * o] / 127.0f, (float)nrm[4 * o + 1] / 127.0f, (float)nrm[4 * o + 2] / 127.0f};
                const float nl = length(n);
                if (nl > 0.0f) n = n / nl;
                // indirect.frag:157-169
                V3 col = {O_POW((float)alb[4 * o] / 255.0f, 2.2f), O_POW((float)alb[4 * o + 1] / 255.0f, 2.2f), O_POW((float)alb[4 * o + 2] / 255.0f, 2.2f)};
                V3 sp = {p.x + n.x * 0.06f, p.y + n.y * 0.06f, p.z + n.z * 0.06f};
                V4 s4 = mul(ShadowProj, mul(ShadowView, V4{sp.x, sp.y, sp.z, 1.0f}));
                float sx = s4.x / s4.w, sy = s4.y / s4.w, sz = s4.z / s4.w;
                sx = sx * 0.5f + 0.5f; sy = sy * 0.5f + 0.5f;
                const int tx = dm_f2i(sx * (float)S), ty = dm_f2i(sy * (float)S);
                float shadowZ = 0.0f;
                if (tx >= 0 && ty >= 0 && tx < (int)S && ty < (int)S) shadowZ = shadow[(size_t)ty * S + tx];
                const float shade = dm_step(sz + 0.005f, shadowZ);
                const float l = fabsf(dot(neg(sunPos), n));
                const float rr[3] = {col.x * sun->luminance[0] * l * shade, col.y * sun->luminance[1] * l * shade, col.z * sun->luminance[2] * l * shade};
                for (int k = 0; k < 3; k++) rad[4 * o + k] = (uint8_t)floorf(dm_clamp(rr[k] * inv_exposure, 0.0f, 1.0f) * 255.0f + 0.5f);
                rad[4 * o + 3] = 255;
            }
    c->stage_ms[F184_STAGE_INJECT] = (float)(now_ms() - t0);
    return F184_OK;
}

// =================================================================================================
// B.4 six-direction mips (integer arithmetic: bit-exact by construction)
// =================================================================================================
// Directions: 0:+x 1:-x 2:+y 3:-y 4:+z 5:-z = the direction a ray TRAVELS.  For each of the 4 columns of a
// 2x2x2 block parallel to the axis, f = the child the ray meets first, b = the second:
//   col = f*255 + (255 - f.a) * b      (premultiplied "over", scaled by 255)
//   out = (sum of 4 cols + 510) / 1020 (mean, rounded half up, back to 8 bits)
extern "C" int orc_mips_n(f184o_ctx* c)
{
    double t0 = now_ms();
    int rc = ensure_image(c, F184_SLOT_RADIANCE); if (rc) return rc;
    rc = ensure_image(c, F184_SLOT_MIPS); if (rc) return rc;
    const uint32_t N = c->cfg.grid_n;
    uint8_t* mips = image_ptr<uint8_t>(c, F184_SLOT_MIPS);
    const uint8_t* l0 = image_ptr<uint8_t>(c, F184_SLOT_RADIANCE);
    uint64_t off = 0;                           // texel offset of level l, direction 0
    uint64_t prev_off = 0;
    for (uint32_t n = N / 2, level = 1; n >= 1; n /= 2, level++)
    {
        const uint32_t sn = n * 2;
        for (int d = 0; d < 6; d++)
        {
            const uint8_t* src = (level == 1) ? l0 : mips + 4 * (prev_off + (uint64_t)d * sn * sn * sn);
            uint8_t* dst = mips + 4 * (off + (uint64_t)d * n * n * n);
            const int axis = d / 2, neg_dir = d & 1;
#pragma omp parallel for schedule(static)
            for (int64_t z = 0; z < (int64_t)n; z++)
                for (uint32_t y = 0; y < n; y++)
                    for (uint32_t x = 0; x < n; x++)
                    {
                        uint32_t sum[4] = {0, 0, 0, 0};
                        for (int j = 0; j < 2; j++)
                            for (int i = 0; i < 2; i++)
                            {
                                uint32_t o[3];       // offsets inside the block for (axis, other1, other2)
                                auto child = [&](int along) {
                                    o[axis] = along; o[(axis + 1) % 3] = i; o[(axis + 2) % 3] = j;
                                    return src + 4 * ((((size_t)(2 * z + o[2])) * sn + (2 * y + o[1])) * sn + (2 * x + o[0]));
                                };
                                const uint8_t* f = child(neg_dir ? 1 : 0);   // travelling +axis meets the low child first
                                const uint8_t* b = child(neg_dir ? 0 : 1);
                                for (int ch = 0; ch < 4; ch++) sum[ch] += (uint32_t)f[ch] * 255u + (255u - f[3]) * b[ch];
                            }
                        uint8_t* o8 = dst + 4 * (((size_t)z * n + y) * n + x);
                        for (int ch = 0; ch < 4; ch++) o8[ch] = (uint8_t)((sum[ch] + 510u) / 1020u);
                    }
        }
        prev_off = off;
        off += 6ull * n * n * n;
    }
    c->stage_ms[F184_STAGE_MIPS] = (float)(now_ms() - t0);
    return F184_OK;
}

// =================================================================================================
// B.5 cone tracing
// =================================================================================================
namespace {

// Emulation of a hardware trilinear fetch from an RGBA8 3D texture: normalised coordinates, border
// addressing (texels outside read 0), filter weights with 8 fractional bits.
struct Vol { const uint8_t* px; uint32_t n; };

inline float q8(float f) { return floorf(f * 256.0f + 0.5f) * (1.0f / 256.0f); }

V4 fetch_trilinear(const Vol& v, float qx, float qy, float qz)
{
    const float n = (float)v.n;
    float x = qx * n - 0.5f, y = qy * n - 0.5f, z = qz * n - 0.5f;
    float x0 = floorf(x), y0 = floorf(y), z0 = floorf(z);
    float fx = q8(x - x0), fy = q8(y - y0), fz = q8(z - z0);
    int ix = (int)x0, iy = (int)y0, iz = (int)z0;
    float acc[4] = {0, 0, 0, 0};
    for (int dz = 0; dz < 2; dz++)
        for (int dy = 0; dy < 2; dy++)
            for (int dx = 0; dx < 2; dx++)
            {
                int xi = ix + dx, yi = iy + dy, zi = iz + dz;
                if (xi < 0 || yi < 0 || zi < 0 || xi >= (int)v.n || yi >= (int)v.n || zi >= (int)v.n) continue;
                float w = (dx ? fx : 1.0f - fx) * (dy ? fy : 1.0f - fy) * (dz ? fz : 1.0f - fz);
                const uint8_t* t = v.px + 4 * (((size_t)zi * v.n + yi) * v.n + xi);
                for (int ch = 0; ch < 4; ch++) acc[ch] += w * ((float)t[ch] * (1.0f / 255.0f));
            }
    return {acc[0], acc[1], acc[2], acc[3]};
}

struct ConeCtx
{
    Vol level0;
    std::vector<Vol> dir[6];     // dir[d][l-1] = level l
    M4 w2v;
    float h, max_dist, exposure;
    uint64_t samples;
    bool spec_b;                 // F184_FLAG_SPEC_APPENDIX_B: SURVEY.md Appendix B.5 as written (mip-linear, half-diameter steps)
};

// direction-weighted fetch from the six-direction chain at mip index `l` (0 = level 1), nearest level: one
// trilinear fetch per axis.  (The first version blended two levels per fetch, "quadrilinear"; on the device
// that made the tracer texture-pipe bound at 88 % of the TEX wavefront peak, so the spec now takes the level
// whose voxel size is nearest the cone diameter — DESIGN.md B.5.)
V4 fetch_dir(const ConeCtx& C, const float w[3], const int face[3], float qx, float qy, float qz, int l)
{
    const int maxl = (int)C.dir[0].size() - 1;
    if (l > maxl) l = maxl;
    V4 r = {0, 0, 0, 0};
    for (int a = 0; a < 3; a++)
    {
        if (w[a] == 0.0f) continue;
        V4 s0 = fetch_trilinear(C.dir[face[a]][l], qx, qy, qz);
        r = {r.x + w[a] * s0.x, r.y + w[a] * s0.y, r.z + w[a] * s0.z, r.w + w[a] * s0.w};
    }
    return r;
}

// Appendix-B sampling above lod 1: the hardware's linear mip filter between level floor(lod) and the next one of the
// six-direction chain (level fraction with 8 fractional bits, like the texel weights); lodp = lod - 1 indexes the chain
V4 fetch_dir_linear(const ConeCtx& C, const float w[3], const int face[3], float qx, float qy, float qz, float lodp)
{
    const int maxl = (int)C.dir[0].size() - 1;
    if (lodp >= (float)maxl) return fetch_dir(C, w, face, qx, qy, qz, maxl);
    const float lf = floorf(lodp);
    const int l = (int)lf;
    const float f = q8(lodp - lf);
    V4 a = fetch_dir(C, w, face, qx, qy, qz, l);
    if (f == 0.0f) return a;
    V4 b = fetch_dir(C, w, face, qx, qy, qz, l + 1);
    return {a.x + f * (b.x - a.x), a.y + f * (b.y - a.y), a.z + f * (b.z - a.z), a.w + f * (b.w - a.w)};
}

// one cone; returns radiance (world units) including the sky term for the unoccluded remainder
V3 trace_cone(ConeCtx& C, V3 origin, V3 dir, float tan_half)
{
    const float h = C.h;
    float w[3] = {dir.x * dir.x, dir.y * dir.y, dir.z * dir.z};
    // the volumes are indexed in VOXEL axes; the voxel camera maps world axes to voxel axes by a signed
    // permutation, so weight/face selection uses the direction expressed in voxel space
    V4 dv4 = mul(C.w2v, V4{dir.x, dir.y, dir.z, 0.0f});
    V3 dv = {dv4.x * 0.5f, dv4.y * 0.5f, dv4.z};          // per-unit-length change of normalised voxel coords
    const float dl = length(dv);
    V3 du = {dv.x / dl, dv.y / dl, dv.z / dl};
    w[0] = du.x * du.x; w[1] = du.y * du.y; w[2] = du.z * du.z;
    const int face[3] = {du.x < 0.0f ? 1 : 0, du.y < 0.0f ? 3 : 2, du.z < 0.0f ? 5 : 4};
    float t = 2.0f * h, A = 0.0f;
    V3 acc = {0, 0, 0};
    while (A < 0.95f && t < C.max_dist)
    {
        const float diam = std::max(h, 2.0f * t * tan_half);
        const float lod = O_LOG2(diam / h);
        V3 p = {origin.x + dir.x * t, origin.y + dir.y * t, origin.z + dir.z * t};
        V4 q4 = mul(C.w2v, V4{p.x, p.y, p.z, 1.0f});
        const float qx = q4.x * 0.5f + 0.5f, qy = q4.y * 0.5f + 0.5f, qz = q4.z;
        if (!(qx >= 0.0f && qx <= 1.0f && qy >= 0.0f && qy <= 1.0f && qz >= 0.0f && qz <= 1.0f)) break;
        C.samples++;
        V4 s;
        if (C.spec_b)
        {   // SURVEY.md Appendix B.5 as written: mip-linear — below lod 1 between the isotropic level 0 and the directional level 1
            if (lod < 1.0f)
            {
                V4 a = fetch_trilinear(C.level0, qx, qy, qz), b = fetch_dir(C, w, face, qx, qy, qz, 0);
                s = {a.x + lod * (b.x - a.x), a.y + lod * (b.y - a.y), a.z + lod * (b.z - a.z), a.w + lod * (b.w - a.w)};
            }
            else s = fetch_dir_linear(C, w, face, qx, qy, qz, lod - 1.0f);
        }
        else
        {   // nearest level: 0 = the isotropic radiance volume, L >= 1 = the six-direction chain
            const int L = (int)floorf(lod + 0.5f);
            s = (L <= 0) ? fetch_trilinear(C.level0, qx, qy, qz) : fetch_dir(C, w, face, qx, qy, qz, L - 1);
        }
        const float k = 1.0f - A;
        acc = {acc.x + k * s.x, acc.y + k * s.y, acc.z + k * s.z};
        A += k * s.w;
        t += C.spec_b ? 0.5f * diam : diam;      // amended spec: one sample per voxel of the level along the axis
    }
    const float rem = std::max(0.0f, 1.0f - A);
    return {acc.x * C.exposure + 0.7f * 0.4f * rem, acc.y * C.exposure + 0.8f * 0.4f * rem, acc.z * C.exposure + 1.0f * 0.4f * rem};
}

inline float unorm16(uint16_t v) { return (float)v / 65535.0f; }

// tangent-space diffuse cone set: one along the normal (w 1/4), five at 60 deg polar, 72 deg apart (w 3/20)
const float kDiffuseDirs[6][3] = {
    {0.0f, 0.0f, 1.0f},
    {0.8660254f, 0.0f, 0.5f},
    {0.26761657f, 0.82363910f, 0.5f},
    {-0.70062927f, 0.50903696f, 0.5f},
    {-0.70062927f, -0.50903696f, 0.5f},
    {0.26761657f, -0.82363910f, 0.5f}};
const float kDiffuseW[6] = {0.25f, 0.15f, 0.15f, 0.15f, 0.15f, 0.15f};
const float kTanHalfDiffuse = 0.57735027f;    // tan(30 deg)

}  // namespace

extern "C" int orc_trace_n(f184o_ctx* c, const f184_trace_constants* k)
{
    double t0 = now_ms();
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_MATERIAL, F184_SLOT_RADIANCE, F184_SLOT_MIPS, F184_SLOT_INDIRECT_OUT, F184_SLOT_INDIRECT_HISTORY})
    { int rc = ensure_image(c, s); if (rc) return rc; }
    const uint32_t W = c->cfg.width, H = c->cfg.height, N = c->cfg.grid_n;
    ConeCtx C0;
    C0.level0 = {image_ptr<uint8_t>(c, F184_SLOT_RADIANCE), N};
    {
        const uint8_t* mips = image_ptr<uint8_t>(c, F184_SLOT_MIPS);
        uint64_t off = 0;
        for (uint32_t n = N / 2; n >= 1; n /= 2)
        {
            for (int d = 0; d < 6; d++) C0.dir[d].push_back(Vol{mips + 4 * (off + (uint64_t)d * n * n * n), n});
            off += 6ull * n * n * n;
        }
    }
    C0.w2v = matmul(load_m4(k->ext.VoxelProj), load_m4(k->ext.VoxelView));
    {
        const M4 v2w = invert(C0.w2v);
        // world size of one voxel along voxel-x
        V4 a = mul(v2w, V4{2.0f / (float)N, 0.0f, 0.0f, 0.0f});
        C0.h = length(V3{a.x, a.y, a.z});
    }
    C0.max_dist = c->cfg.cone_max_distance;
    C0.exposure = exposure_of(c, &k->sun);
    C0.samples = 0;
    C0.spec_b = (c->cfg.flags & F184_FLAG_SPEC_APPENDIX_B) != 0;
    const M4 InvProj = load_m4(k->view.InvProj), InvModelView = load_m4(k->ext.InvModelView);
    const M4 prevModelView = load_m4(k->prev.PrevModelView), prevProjection = load_m4(k->prev.PrevProjection);
    const float* depthp = image_ptr<float>(c, F184_SLOT_DEPTH);
    const uint16_t* normals = image_ptr<uint16_t>(c, F184_SLOT_NORMALS);
    const uint8_t* material = image_ptr<uint8_t>(c, F184_SLOT_MATERIAL);
    uint16_t* hist = image_ptr<uint16_t>(c, F184_SLOT_INDIRECT_HISTORY);
    uint16_t* out = image_ptr<uint16_t>(c, F184_SLOT_INDIRECT_OUT);
    // view window (f184o_trace_views): the view's rows are an image of their own — uv, history addressing and history reset
    const uint32_t vy0 = c->view_h ? c->view_y0 : 0, vh = c->view_h ? c->view_h : H;
    if (k->reset_history) memset(hist + (size_t)4 * W * vy0, 0, (size_t)W * vh * 8);
    const uint32_t y0 = c->view_h ? vy0 : c->row0, y1 = c->view_h ? vy0 + vh : std::min(c->row1, H);
    uint64_t total = 0;
    // camera position in world space = InvModelView * (0,0,0,1)
    const V3 cam = {InvModelView.m[12], InvModelView.m[13], InvModelView.m[14]};

#pragma omp parallel for schedule(dynamic, 2) reduction(+ : total)
    for (int64_t y = y0; y < (int64_t)y1; y++)
    {
        if (!c->view_h && (uint32_t)(y >> 3) % c->tile_stride != c->tile_first) continue;      // tile rows interleaved over the ranks
        ConeCtx C = C0;
        C.samples = 0;
        for (uint32_t x = 0; x < W; x++)
        {
            const float uvx = ((float)x + 0.5f) / (float)W, uvy = ((float)(y - vy0) + 0.5f) / (float)vh;
            const float depth = depthp[(size_t)y * W + x];
            V4 cp = mul(InvProj, V4{uvx * 2.0f - 1.0f, uvy * 2.0f - 1.0f, depth, 1.0f});
            V3 cspos = {cp.x / cp.w, cp.y / cp.w, cp.z / cp.w};
            uint16_t* o = &out[4 * ((size_t)y * W + x)];
            if (depth >= 1.0f)      // sky: nothing to shade
            {
                o[0] = o[1] = o[2] = 0; o[3] = dm_f32_to_f16(-cspos.z);
                continue;
            }
            V4 wp4 = mul(InvModelView, V4{cspos.x, cspos.y, cspos.z, 1.0f});
            V3 wpos = {wp4.x, wp4.y, wp4.z};
            const uint16_t* np = &normals[4 * ((size_t)y * W + x)];
            V3 csnorm = normalize(V3{fmaf(unorm16(np[0]), 2.0f, -1.0f), fmaf(unorm16(np[1]), 2.0f, -1.0f), fmaf(unorm16(np[2]), 2.0f, -1.0f)});
            V3 wnorm = mul3(InvModelView, csnorm);
            // tangent frame, indirect.frag:71-86
            V3 z = wnorm, hh = wnorm;
            if (fabsf(hh.x) <= fabsf(hh.y) && fabsf(hh.x) <= fabsf(hh.z)) hh.x = 1.0f;
            else if (fabsf(hh.y) <= fabsf(hh.x) && fabsf(hh.y) <= fabsf(hh.z)) hh.y = 1.0f;
            else hh.z = 1.0f;
            z = normalize(z);
            V3 ty = normalize(cross(hh, z));
            V3 tx = normalize(cross(z, ty));
            V3 origin = {wpos.x + z.x * C.h, wpos.y + z.y * C.h, wpos.z + z.z * C.h};
            V3 ind = {0, 0, 0};
            for (int i = 0; i < 6; i++)
            {
                const float* d = kDiffuseDirs[i];
                V3 dir = {(tx.x * d[0] + ty.x * d[1]) + z.x * d[2], (tx.y * d[0] + ty.y * d[1]) + z.y * d[2], (tx.z * d[0] + ty.z * d[1]) + z.z * d[2]};
                V3 r = trace_cone(C, origin, dir, kTanHalfDiffuse);
                ind = {ind.x + kDiffuseW[i] * r.x, ind.y + kDiffuseW[i] * r.y, ind.z + kDiffuseW[i] * r.z};
            }
            {   // specular cone about the mirror direction; Schlick Fresnel of aggregateLights.frag:135-141
                const float rough = (float)material[4 * ((size_t)y * W + x) + 1] / 255.0f;
                const float tan_half = dm_clamp(rough * rough, 0.02f, 0.6f);
                V3 I = normalize(wpos - cam);
                const float ndi = dot(z, I);
                V3 R = {I.x - 2.0f * ndi * z.x, I.y - 2.0f * ndi * z.y, I.z - 2.0f * ndi * z.z};
                if (dot(R, z) > 0.0f)
                {
                    V3 r = trace_cone(C, origin, R, tan_half);
                    const float ct = std::max(-ndi, 0.0f);
                    const float om = 1.0f - ct;
                    const float F = 0.04f + 0.96f * (om * om * om * om * om);
                    ind = {ind.x + F * r.x, ind.y + F * r.y, ind.z + F * r.z};
                }
            }
            // temporal reprojection, indirect.frag:225-240 (same as mode R)
            V4 pc = mul(prevModelView, V4{wpos.x, wpos.y, wpos.z, 1.0f});
            V4 pp = mul(prevProjection, pc);
            float ru = pp.x / pp.w, rv = pp.y / pp.w;
            ru = ru * 0.5f + 0.5f; rv = rv * 0.5f + 0.5f;
            if (dm_clamp(ru, 0.0f, 1.0f) == ru && dm_clamp(rv, 0.0f, 1.0f) == rv)
            {
                float fx = ru * (float)W - 0.5f, fy = rv * (float)vh - 0.5f;
                float x0f = floorf(fx), y0f = floorf(fy);
                float wx = fx - x0f, wy = fy - y0f;
                int xi0 = dm_f2i(x0f), yi0 = dm_f2i(y0f);
                int xa = wrapi(xi0, (int)W), xb = wrapi(xi0 + 1, (int)W), ya = (int)vy0 + wrapi(yi0, (int)vh), yb = (int)vy0 + wrapi(yi0 + 1, (int)vh);
                float prev[4];
                for (int ch = 0; ch < 4; ch++)
                {
                    float a = dm_f16_to_f32(hist[4 * ((size_t)ya * W + xa) + ch]), b = dm_f16_to_f32(hist[4 * ((size_t)ya * W + xb) + ch]);
                    float cc = dm_f16_to_f32(hist[4 * ((size_t)yb * W + xa) + ch]), d = dm_f16_to_f32(hist[4 * ((size_t)yb * W + xb) + ch]);
                    prev[ch] = (a * (1.0f - wx) + b * wx) * (1.0f - wy) + (cc * (1.0f - wx) + d * wx) * wy;
                }
                float bw = 0.95f * dm_smoothstep(0.0f, 1.0f, 1.0f - fabsf(prev[3] + cspos.z));
                ind = {dm_clamp(dm_mix(ind.x, prev[0], bw), 0.0f, 16.0f), dm_clamp(dm_mix(ind.y, prev[1], bw), 0.0f, 16.0f),
                       dm_clamp(dm_mix(ind.z, prev[2], bw), 0.0f, 16.0f)};
            }
            o[0] = dm_f32_to_f16(ind.x); o[1] = dm_f32_to_f16(ind.y); o[2] = dm_f32_to_f16(ind.z); o[3] = dm_f32_to_f16(-cspos.z);
        }
        total += C.samples;
    }
    c->counters[F184_COUNTER_MARCH_STEPS] = (c->keep_samples ? c->counters[F184_COUNTER_MARCH_STEPS] : 0) + total;
    c->stage_ms[F184_STAGE_TRACE] = (c->keep_samples ? c->stage_ms[F184_STAGE_TRACE] : 0.0f) + (float)(now_ms() - t0);
    return F184_OK;
}

// f184_trace_views (include/f184.h): view v = rows [v*view_h, (v+1)*view_h), traced with ks[v]
extern "C" int orc_trace_views_n(f184o_ctx* c, const f184_trace_constants* ks, uint32_t view_h, uint32_t first, uint32_t count)
{
    const uint32_t H = c->cfg.height;
    if (view_h == 0 || H % view_h != 0 || (uint64_t)(first + (uint64_t)count) * view_h > H) return F184_ERR_INVALID_ARGUMENT;
    c->counters[F184_COUNTER_MARCH_STEPS] = 0;
    c->stage_ms[F184_STAGE_TRACE] = 0.0f;
    int rc = F184_OK;
    for (uint32_t v = first; v < first + count && rc == F184_OK; v++)
    {
        c->view_y0 = v * view_h; c->view_h = view_h; c->keep_samples = true;
        rc = orc_trace_n(c, &ks[v]);
    }
    c->view_y0 = 0; c->view_h = 0; c->keep_samples = false;
    return rc;
}
