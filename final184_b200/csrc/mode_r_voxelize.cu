// mode_r_voxelize.cu — reference-faithful voxelization (F184_MODE_REFERENCE).
//
// Replaces: ClearImage(VoxelImage) (Foreground/Renderer/MegaPipeline.cpp:196) + the voxelization render
// pass (:218-223 → CVoxelizeRenderer::RenderList, VoxelizeRenderer.cpp:19-30, 81-127) whose GPU programs
// are Pipelang/Internal/main.lua:60-75 (VS), 83-144 (VoxelGS), 179-206 (BasicMaterial), 242-275 (VoxelPS)
// behind a 1-sample, cull-none, depth-off fixed-function rasteriser (VoxelizeRenderer.cpp:50-56).
//
// B200 design.  The reference issues 103 draws through VS → GS → raster → PS; 96 % of Sponza's triangles
// cover no pixel centre of the 128^2 viewport, so the work is triangle-setup bound, not fill bound.  Here
// ONE launch walks all triangles, 32 per warp:
//   phase 1  every lane sets its own triangle up (GS maths, snap to 1/256 px, integer edge functions)
//            and parks the survivors in shared memory;
//   phase 2  the warp drains its survivors one at a time, lanes striding over the pixels of the
//            bounding box, so a floor quad covering thousands of pixels costs the same per pixel as a
//            one-pixel triangle (no long-pole thread);
//   store    the reference's plain imageStore race (main.lua:273) is made deterministic: each texel is
//            written with atomicMax on a 64-bit key (draw order + 1) << 32 | RG16UI texel, which yields
//            exactly "last writer in draw order wins" for any scheduling.  A resolve pass copies the low
//            32 bits into the RG16UI volume and zeroes the key, so it doubles as next frame's clear.
// Texture filtering is done in software with fp32 weights (8 texel reads per tap) instead of the texture
// unit, whose 8-bit weights would break bit-exact parity; fragments are few (0.18 M at 128^3, 2.9 M at 512^3).
#include <algorithm>

#include "f184_device.cuh"

namespace {

struct TriShared
{
    int x[3], y[3];          // snapped vertices, 1/256 px, ordered so that area > 0
    float cz[3];             // raster depth per vertex (same order)
    float u[3], v[3];
    f3 n[3];
    float dudx, dvdx, dudy, dvdy;
    long long area;
    int px0, py0, bw, bh;
    uint32_t tri;            // draw-order index
    uint16_t mat;
    uint16_t orient;
};

constexpr int WARPS_PER_BLOCK = 4;

__device__ __forceinline__ f4 bilinear_level(const TexDev& t, uint32_t level, float u, float v)
{
    uint32_t w = max(1u, t.w >> level), h = max(1u, t.h >> level);
    const uchar4* px = reinterpret_cast<const uchar4*>(t.base + t.off[level]);
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = wrap_pow2(dm_f2i(x0f), (int)w), y0 = wrap_pow2(dm_f2i(y0f), (int)h);
    int x1 = wrap_pow2(x0 + 1, (int)w), y1 = wrap_pow2(y0 + 1, (int)h);
    uchar4 t00 = __ldg(px + (size_t)y0 * w + x0), t10 = __ldg(px + (size_t)y0 * w + x1);
    uchar4 t01 = __ldg(px + (size_t)y1 * w + x0), t11 = __ldg(px + (size_t)y1 * w + x1);
    auto lerp2 = [&](unsigned char a, unsigned char b, unsigned char c, unsigned char d) {
        float fa = (float)a / 255.0f, fb = (float)b / 255.0f, fc = (float)c / 255.0f, fd = (float)d / 255.0f;
        float top = fa * (1.0f - fx) + fb * fx;
        float bot = fc * (1.0f - fx) + fd * fx;
        return top * (1.0f - fy) + bot * fy;
    };
    return {lerp2(t00.x, t10.x, t01.x, t11.x), lerp2(t00.y, t10.y, t01.y, t11.y), lerp2(t00.z, t10.z, t01.z, t11.z),
            lerp2(t00.w, t10.w, t01.w, t11.w)};
}

// texture(sampler2D(BaseColorTex, GlobalLinearSampler), uv): linear/linear/linear-mip, wrap, MaxLod 4
// (MegaPipeline.cpp:39-43), LOD from the (constant, affine) screen-space derivatives.
__device__ __forceinline__ f4 sample_trilinear(const TexDev& t, float u, float v, float dudx, float dvdx, float dudy, float dvdy)
{
    float ax = dudx * (float)t.w, ay = dvdx * (float)t.h;
    float bx = dudy * (float)t.w, by = dvdy * (float)t.h;
    float mx = __fsqrt_rn(ax * ax + ay * ay), my = __fsqrt_rn(bx * bx + by * by);
    float rho = mx > my ? mx : my;
    float maxlod = (float)min(4u, t.nlevels - 1u);
    float lod = 0.0f;
    if (rho > 1.0f) lod = dm_log2(rho);
    if (!(lod < maxlod)) lod = maxlod;
    float lf = floorf(lod);
    uint32_t l0 = (uint32_t)lf;
    float f = lod - lf;
    f4 c0 = bilinear_level(t, l0, u, v);
    if (f == 0.0f) return c0;
    f4 c1 = bilinear_level(t, l0 + 1, u, v);
    return {c0.x * (1.0f - f) + c1.x * f, c0.y * (1.0f - f) + c1.y * f, c0.z * (1.0f - f) + c1.z * f, c0.w * (1.0f - f) + c1.w * f};
}

__device__ __forceinline__ long long ceil_div256(long long a) { return (a + 255) >> 8; }     // arithmetic shift = floor
__device__ __forceinline__ long long floor_div256(long long a) { return a >> 8; }

// One candidate pixel of one set-up triangle: coverage (integer edge functions, top-left rule), depth clip, material, pack,
// ordered store.  Returns 1 if a texel was stored.
struct EdgeSetup { long long ex[3], ey[3]; int bias[3]; float areaf; };
__device__ __forceinline__ EdgeSetup edge_setup(const TriShared& s)
{
    EdgeSetup E;
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        E.ex[k] = (long long)s.x[b] - s.x[a]; E.ey[k] = (long long)s.y[b] - s.y[a];
        const bool top_left = (E.ey[k] < 0) || (E.ey[k] == 0 && E.ex[k] > 0);
        E.bias[k] = top_left ? 0 : -1;
    }
    E.areaf = (float)s.area;
    return E;
}
__device__ __forceinline__ unsigned int shade_pixel(const TriShared& s, const EdgeSetup& E, const MatDev& mat, const TexDev* __restrict__ texs,
                                                    int px, int py, uint32_t N, float maxDepth, unsigned long long* __restrict__ keys)
{
    const long long cxp = (long long)px * 256 + 128, cyp = (long long)py * 256 + 128;
    long long w[3];
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        const int a = (k + 1) % 3;
        w[k] = E.ex[k] * (cyp - s.y[a]) - E.ey[k] * (cxp - s.x[a]);
        if (w[k] + E.bias[k] < 0) inside = false;
    }
    if (!inside) return 0;
    const float areaf = E.areaf;
    const float b0 = (float)w[0] / areaf, b1 = (float)w[1] / areaf, b2 = (float)w[2] / areaf;
    const float z = (s.cz[0] * b0 + s.cz[1] * b1) + s.cz[2] * b2;
    if (!(z >= 0.0f && z <= 1.0f)) return 0;                    // depth clip, no clamp
    const float u = (s.u[0] * b0 + s.u[1] * b1) + s.u[2] * b2;
    const float v = (s.v[0] * b0 + s.v[1] * b1) + s.v[2] * b2;
    const f3 n = {(s.n[0].x * b0 + s.n[1].x * b1) + s.n[2].x * b2, (s.n[0].y * b0 + s.n[1].y * b1) + s.n[2].y * b2,
                  (s.n[0].z * b0 + s.n[1].z * b1) + s.n[2].z * b2};
    // ---- BasicMaterial, main.lua:188-205
    f4 base;
    if (!mat.use_textures) base = {mat.factor[0], mat.factor[1], mat.factor[2], mat.factor[3]};
    else
    {
        f4 sc = {0.f, 0.f, 0.f, 0.f};
        if (mat.tex >= 0) sc = sample_trilinear(texs[mat.tex], u, v, s.dudx, s.dvdx, s.dudy, s.dvdy);
        base = {sc.x * mat.factor[0], sc.y * mat.factor[1], sc.z * mat.factor[2], sc.w * mat.factor[3]};
        if (base.w < 0.05f) return 0;                            // discard
    }
    // ---- VoxelPS, main.lua:249-273
    const float fxc = (float)px + 0.5f, fyc = (float)py + 0.5f;
    float vx, vy, vz;
    if (s.orient == 0) { vx = fxc; vy = fyc; vz = z * maxDepth; }
    else if (s.orient == 1) { vx = (1.0f - z) * maxDepth; vy = fyc; vz = fxc; }
    else { vx = fxc; vy = z * maxDepth; vz = maxDepth - fyc; }
    const int ix = dm_f2i(vx), iy = dm_f2i(vy), iz = dm_f2i(vz);
    if (ix < 0 || iy < 0 || iz < 0 || ix >= (int)N || iy >= (int)N || iz >= (int)N) return 0;
    const uint32_t pc = (dm_f2uint(base.x * 31.0f) << 11) | (dm_f2uint(base.y * 63.0f) << 5) | dm_f2uint(base.z * 31.0f);
    const uint32_t pn = (dm_f2uint(n.x * 16.0f + 15.0f) << 11) | (dm_f2uint(n.y * 32.0f + 31.0f) << 5) | dm_f2uint(n.z * 16.0f + 15.0f);
    const unsigned long long key = ((unsigned long long)(s.tri + 1u) << 32) | ((pn & 0xffffu) << 16) | (pc & 0xffffu);
    atomicMax(keys + (((size_t)iz * N + iy) * N + ix), key);
    return 1;
}

// A bounding box of more than this many pixels goes to the second launch (one CTA per triangle); below it the owning warp
// drains it on the spot (<= 4 strides of 32 lanes).
constexpr int BIG_BOX = 128;

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_voxelize_r(const float* __restrict__ pos, const float* __restrict__ nrm, const float* __restrict__ uv,
             const uint32_t* __restrict__ idx, const uint16_t* __restrict__ tri_mat, const uint16_t* __restrict__ tri_model,
             const M4* __restrict__ model_mats, const M4* __restrict__ vm_mats, M4 Proj, const TexDev* __restrict__ texs,
             const MatDev* __restrict__ mats, uint32_t tri_first, uint32_t tri_end, uint32_t N,
             unsigned long long* __restrict__ keys, unsigned long long* __restrict__ frag_counter, TriShared* __restrict__ big_queue,
             unsigned int* __restrict__ big_count)
{
    __shared__ TriShared sh[WARPS_PER_BLOCK][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warp_global = blockIdx.x * WARPS_PER_BLOCK + warp;
    const uint32_t t = tri_first + warp_global * 32 + lane;
    const float Nf = (float)N, halfN = Nf * 0.5f, maxDepth = (float)(N - 1);
    bool active = false;

    if (t < tri_end)
    {
        const uint32_t i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
        const uint32_t id[3] = {i0, i1, i2};
        const uint32_t model = tri_model[t];
        const M4& vm = vm_mats[model];
        // ---- VoxelGS, main.lua:94-115
        f3 vp[3];
#pragma unroll
        for (int i = 0; i < 3; i++)
            vp[i] = mul43(vm, f3{pos[3 * id[i]], pos[3 * id[i] + 1], pos[3 * id[i] + 2]}, 1.0f);
        f3 fn = abs3(cross3(vp[1] - vp[0], vp[2] - vp[0]));
        int orient;
        if (fn.x > fn.y) orient = (fn.x > fn.z) ? 1 : 0;
        else orient = (fn.y > fn.z) ? 2 : 0;
        // ---- main.lua:116-134, then viewport transform + snap
        float cz[3];
        int X[3], Y[3];
        bool bad = false;
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            f4 g = mul44(Proj, f4{vp[i].x, vp[i].y, vp[i].z, 1.0f});
            float gx = g.x / g.w, gy = g.y / g.w, gz = g.z / g.w;
            if (orient == 1) { float nx = gz * 2.0f - 1.0f; float nz = -gx / 2.0f + 0.5f; gx = nx; gz = nz; }
            else if (orient == 2) { float ny = 1.0f - gz * 2.0f; float nz = gy / 2.0f + 0.5f; gy = ny; gz = nz; }
            cz[i] = gz;
            float xf = gx * halfN + halfN, yf = gy * halfN + halfN;
            if (!(fabsf(xf) < 1048576.0f) || !(fabsf(yf) < 1048576.0f)) bad = true;
            X[i] = (int)rintf(xf * 256.0f);
            Y[i] = (int)rintf(yf * 256.0f);
        }
        if (!bad)
        {
            long long area = (long long)(X[1] - X[0]) * (Y[2] - Y[0]) - (long long)(X[2] - X[0]) * (Y[1] - Y[0]);
            if (area != 0)
            {
                int o1 = 1, o2 = 2;
                if (area < 0) { o1 = 2; o2 = 1; area = -area; }
                const int ord[3] = {0, o1, o2};
                int minx = min(X[0], min(X[1], X[2])), maxx = max(X[0], max(X[1], X[2]));
                int miny = min(Y[0], min(Y[1], Y[2])), maxy = max(Y[0], max(Y[1], Y[2]));
                long long px0 = max(0ll, ceil_div256((long long)minx - 128)), px1 = min((long long)N - 1, floor_div256((long long)maxx - 128));
                long long py0 = max(0ll, ceil_div256((long long)miny - 128)), py1 = min((long long)N - 1, floor_div256((long long)maxy - 128));
                if (px0 <= px1 && py0 <= py1)
                {
                    active = true;
                    TriShared& s = sh[warp][lane];
                    const M4& mm = model_mats[model];
#pragma unroll
                    for (int k = 0; k < 3; k++)
                    {
                        const int i = ord[k];
                        s.x[k] = X[i]; s.y[k] = Y[i]; s.cz[k] = cz[i];
                        s.u[k] = uv[2 * id[i]]; s.v[k] = uv[2 * id[i] + 1];
                        s.n[k] = normalize3(mul33(mm, f3{nrm[3 * id[i]], nrm[3 * id[i] + 1], nrm[3 * id[i] + 2]}));   // main.lua:137
                    }
                    s.area = area;
                    s.px0 = (int)px0; s.py0 = (int)py0; s.bw = (int)(px1 - px0 + 1); s.bh = (int)(py1 - py0 + 1);
                    s.tri = t; s.mat = tri_mat[t]; s.orient = (uint16_t)orient;
                    // attribute gradients for the implicit-derivative LOD (affine: ortho camera)
                    const float areaf = (float)area;
                    float dbdx[3], dbdy[3];
#pragma unroll
                    for (int k = 0; k < 3; k++)
                    {
                        const int a = (k + 1) % 3, b = (k + 2) % 3;
                        long long ex = (long long)s.x[b] - s.x[a], ey = (long long)s.y[b] - s.y[a];
                        dbdx[k] = (float)(-ey * 256) / areaf;
                        dbdy[k] = (float)(ex * 256) / areaf;
                    }
                    s.dudx = (s.u[0] * dbdx[0] + s.u[1] * dbdx[1]) + s.u[2] * dbdx[2];
                    s.dvdx = (s.v[0] * dbdx[0] + s.v[1] * dbdx[1]) + s.v[2] * dbdx[2];
                    s.dudy = (s.u[0] * dbdy[0] + s.u[1] * dbdy[1]) + s.u[2] * dbdy[2];
                    s.dvdy = (s.v[0] * dbdy[0] + s.v[1] * dbdy[1]) + s.v[2] * dbdy[2];
                }
            }
        }
    }
    // large bounding boxes leave the warp: a floor quad covering the whole 128^2 viewport is 512 strides of this warp, and 32 of
    // them in one warp were the kernel's long pole (6.3 ms at 128^3); they are rasterised by k_voxelize_r_big, one CTA each
    if (active && sh[warp][lane].bw * sh[warp][lane].bh > BIG_BOX)
    {
        big_queue[atomicAdd(big_count, 1u)] = sh[warp][lane];
        active = false;
    }
    unsigned int pending = __ballot_sync(0xffffffffu, active);
    __syncwarp();
    unsigned int frags = 0;
    while (pending)
    {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const TriShared& s = sh[warp][src];
        const EdgeSetup E = edge_setup(s);
        const MatDev mat = mats[s.mat];
        const int npix = s.bw * s.bh;
        for (int p = lane; p < npix; p += 32)
            frags += shade_pixel(s, E, mat, texs, s.px0 + p % s.bw, s.py0 + p / s.bw, N, maxDepth, keys);
    }
    warp_count_add(frag_counter, frags);
}

// Second launch: the queued large triangles, one CTA per triangle at a time (grid-stride over the queue), threads striding over
// the bounding box row by row.  Scheduling is free: the 64-bit ordered store makes the result independent of it.
constexpr int BIG_THREADS = 256;
__global__ void __launch_bounds__(BIG_THREADS)
k_voxelize_r_big(const TriShared* __restrict__ big_queue, const unsigned int* __restrict__ big_count, const TexDev* __restrict__ texs,
                 const MatDev* __restrict__ mats, uint32_t N, unsigned long long* __restrict__ keys, unsigned long long* __restrict__ frag_counter)
{
    __shared__ TriShared s;
    const unsigned int n = *big_count;
    const float maxDepth = (float)(N - 1);
    unsigned int frags = 0;
    for (unsigned int q = blockIdx.x; q < n; q += gridDim.x)
    {
        __syncthreads();
        // the record is 30 words: the first 30 threads copy it
        if (threadIdx.x < sizeof(TriShared) / 4) reinterpret_cast<uint32_t*>(&s)[threadIdx.x] = reinterpret_cast<const uint32_t*>(big_queue + q)[threadIdx.x];
        __syncthreads();
        const EdgeSetup E = edge_setup(s);
        const MatDev mat = mats[s.mat];
        // a thread keeps its column offset and walks down rows in steps that cover BIG_THREADS pixels
        const int npix = s.bw * s.bh;
        for (int p = threadIdx.x; p < npix; p += BIG_THREADS)
            frags += shade_pixel(s, E, mat, texs, s.px0 + p % s.bw, s.py0 + p / s.bw, N, maxDepth, keys);
    }
    warp_count_add(frag_counter, frags);
}

// vm[m] = View * Model[m], each element ((a0*b0 + a1*b1) + a2*b2) + a3*b3 — the GLSL association (main.lua:98)
__global__ void k_view_model(M4 View, const M4* __restrict__ model, M4* __restrict__ vm, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 16) return;
    const uint32_t m = i / 16, j = (i % 16) / 4, r = i % 4;
    const M4& B = model[m];
    vm[m].m[4 * j + r] = ((View.m[r] * B.m[4 * j] + View.m[4 + r] * B.m[4 * j + 1]) + View.m[8 + r] * B.m[4 * j + 2]) + View.m[12 + r] * B.m[4 * j + 3];
}

// keys -> RG16UI volume; zero the key behind us (this is ClearImage for the next frame).
__global__ void k_resolve_r(unsigned long long* __restrict__ keys, uint32_t* __restrict__ vox, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
    {
        const unsigned long long k = keys[i];
        vox[i] = (uint32_t)k;
        if (k) keys[i] = 0ull;
    }
}

}  // namespace

int f184_voxelize_r(f184_ctx* c, const f184_view_constants* cam)
{
    int rc = f184_ensure_image(c, F184_SLOT_VOXELS);
    if (rc) return rc;
    const uint32_t N = c->cfg.grid_n;
    const size_t nvox = (size_t)N * N * N;
    if (!c->vox_keys)
    {
        CK(c, cudaMalloc(&c->vox_keys, nvox * 8));
        CK(c, cudaMemsetAsync(c->vox_keys, 0, nvox * 8, c->stream));
    }
    // ViewMat * ModelMat per model, associated as the GLSL does (main.lua:98)
    M4 View, Proj;
    memcpy(View.m, cam->ViewMat, 64);
    memcpy(Proj.m, cam->ProjMat, 64);
    M4* vm_dev = nullptr;
    CK(c, cudaMallocAsync(&vm_dev, sizeof(M4) * c->n_models, c->stream));
    k_view_model<<<(c->n_models * 16 + 127) / 128, 128, 0, c->stream>>>(View, c->model_mats, vm_dev, c->n_models);
    CK_LAUNCH(c);

    const uint32_t first = c->tri_first < c->n_tris ? c->tri_first : c->n_tris;
    const uint64_t end64 = (uint64_t)first + c->tri_count;
    const uint32_t end = end64 < c->n_tris ? (uint32_t)end64 : c->n_tris;

    rc = f184_stage_begin(c, F184_STAGE_VOXELIZE);
    if (rc) return rc;
    CK(c, cudaMemsetAsync(c->counters_dev + F184_COUNTER_FRAGMENTS, 0, 8, c->stream));
    if (end > first)
    {
        static_assert(sizeof(TriShared) % 4 == 0 && sizeof(TriShared) / 4 <= BIG_THREADS, "queue record is copied one word per thread");
        if (c->r_queue_cap < c->n_tris)
        {   // worst case every triangle is large
            if (c->r_queue) cudaFree(c->r_queue);
            CK(c, cudaMalloc(&c->r_queue, sizeof(TriShared) * (size_t)c->n_tris + 16));
            c->r_queue_cap = c->n_tris;
        }
        TriShared* queue = reinterpret_cast<TriShared*>(reinterpret_cast<char*>(c->r_queue) + 16);
        unsigned int* qcount = reinterpret_cast<unsigned int*>(c->r_queue);
        CK(c, cudaMemsetAsync(qcount, 0, 4, c->stream));
        const uint32_t tris = end - first;
        const uint32_t blocks = (tris + WARPS_PER_BLOCK * 32 - 1) / (WARPS_PER_BLOCK * 32);
        k_voxelize_r<<<blocks, WARPS_PER_BLOCK * 32, 0, c->stream>>>(c->pos, c->nrm, c->uv, c->idx, c->tri_mat, c->tri_model,
                                                                     c->model_mats, vm_dev, Proj, c->tex_dev, c->mat_dev, first, end, N,
                                                                     c->vox_keys, c->counters_dev + F184_COUNTER_FRAGMENTS, queue, qcount);
        CK_LAUNCH(c);
        k_voxelize_r_big<<<148 * 8, BIG_THREADS, 0, c->stream>>>(queue, qcount, c->tex_dev, c->mat_dev, N, c->vox_keys,
                                                                 c->counters_dev + F184_COUNTER_FRAGMENTS);
        CK_LAUNCH(c);
    }
    {
        const int threads = 256;
        const int blocks = (int)std::min<size_t>((nvox + threads - 1) / threads, 148 * 16);
        k_resolve_r<<<blocks, threads, 0, c->stream>>>(c->vox_keys, img_ptr<uint32_t>(c, F184_SLOT_VOXELS), nvox);
        CK_LAUNCH(c);
    }
    rc = f184_stage_end(c, F184_STAGE_VOXELIZE);
    if (rc) return rc;
    CK(c, cudaFreeAsync(vm_dev, c->stream));
    return F184_OK;
}
