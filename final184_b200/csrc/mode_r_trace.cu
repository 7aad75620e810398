// mode_r_trace.cu — reference-faithful indirect pass (F184_MODE_REFERENCE).
//
// Replaces the lighting_indirect full-screen pass (Foreground/Renderer/MegaPipeline.cpp:252-268), i.e.
// Shader/Lighting/indirect.frag: per pixel 4 cosine-weighted hemisphere rays, each with an optional second
// bounce, each a 60-step fixed-step march through the RG16UI voxel volume with first-hit termination, a
// shadow-map lookup at the hit, then the temporal reprojection blend.  Noise is Shader/math.inc:83-87,
// 107-110, 189-194, 214-219 evaluated with f184_detmath.h (see that header for why).
//
// B200 design: one thread per pixel, warps own 8x4 pixel tiles so the rays of a warp start from neighbouring surface points
// and walk through neighbouring voxels (the 8 MiB..512 MiB volume is served by L2/L1, 4-byte nearest fetches only when the
// integer voxel changes, as in the shader).  The march is ALU: the affine world->voxel transform is applied per step exactly
// as the shader does (the arithmetic order is part of the parity contract).  What a straight transcription loses is lanes:
// first-hit termination leaves 11 of 32 active while the warp waits for its longest ray, 8 times per pixel.  The kernel
// therefore regroups a pixel's work — convergent set-up, one flattened loop per bounce in which every lane walks its own
// rays back to back, convergent shading — see the block comment above ray_rands.  Results are bit-identical to the shader
// text (tests/golden/refshader_*.npz); the cone-sample counter still counts every step the shader would take.
#include "f184_device.cuh"

namespace {

struct TraceParams
{
    M4 InvProj, InvModelView, ShadowView, ShadowProj, w2voxel, prevModelView, prevProjection;
    const float* depth;
    const uint16_t* normals;
    const float* shadow;
    const uint32_t* vox;
    const uint16_t* hist;
    uint16_t* out;
    const float* rands;
    uint32_t W, H, S, N, steps, y0, y1, tile0, tile_stride;
    float step_size, iiTime, resx, resy;
    f3 sunLum, sunPos;
};

// math.inc:83-87
__device__ __forceinline__ float glsl_hash(float px, float py)
{
    float p3x = dm_fract(px * 0.2031f), p3y = dm_fract(py * 0.2031f), p3z = dm_fract(px * 0.2031f);
    float d = (p3x * (p3y + 19.19f) + p3y * (p3z + 19.19f)) + p3z * (p3x + 19.19f);
    p3x += d; p3y += d; p3z += d;
    return dm_fract((p3x + p3y) * p3z);
}
// math.inc:107-110
__device__ __forceinline__ float nrand(float nx, float ny) { return dm_fract(dm_sin(nx * 12.9898f + ny * 78.233f) * 43758.5453f); }
// math.inc:189-194
__device__ __forceinline__ float n4rand_ss(float nx, float ny, float t0, float t1)
{
    float nrnd0 = nrand(nx + t0, ny + t0);
    float nrnd1 = nrand(nx + t1, ny + t1);
    return 0.23f * __fsqrt_rn(-dm_log(nrnd0 + 0.00001f)) * dm_cos(2.0f * 3.141592f * nrnd1) + 0.5f;
}
// math.inc:214-219
__device__ __forceinline__ float blugausnoise2(float cx, float cy, float t0, float t1)
{
    float nrand1 = n4rand_ss(cx, cy, t0, t1);
    float nrand0 = n4rand_ss(cx - 1.0f, cy, t0, t1);
    float nrand2 = n4rand_ss(cx + 1.0f, cy, t0, t1);
    return 2.0f * nrand1 - 0.5f * (nrand0 + nrand2);
}

struct Hit { f3 wpos, wnorm, brdf; bool hit; };

__device__ __forceinline__ f3 voxel_pos(const M4& w2v, f3 p, float Nf)
{
    f3 v = mul43(w2v, p, 1.0f);
    return {(v.x * 0.5f + 0.5f) * Nf, (v.y * 0.5f + 0.5f) * Nf, v.z * Nf};
}

// ---- one march (indirect.frag:104-184) in three convergent pieces and one flattened loop ------------------------------
//
// The shader runs, per pixel, 4 x (primary march [+ secondary march if it hit]), each a loop of up to 60 dependent steps that
// ends at the first occupied voxel.  Executed as written, a warp pays max-over-lanes(steps) for every one of the 8 marches:
// ncu on the straight transcription showed 60 iterations per march at 11 of 32 lanes active (some lane always runs to the
// end), 2.27 G warp instructions per 720p frame.  Here the work of a pixel is regrouped without touching any ray's arithmetic:
//   rands + set-up of the 4 primary rays        convergent (every lane does the same 4 set-ups)
//   march_pool                                  ONE loop per bounce over the block's pool of rays: a lane that finishes a
//                                               ray takes the next unmarched one in the same iteration instead of idling,
//                                               so a warp pays about the mean ray length, not the sum of per-march maxima
//   shading of the hits + set-up of secondaries convergent again (lanes that hit shade together)
//   march_pool over the secondary rays, shading, and the sum in the shader's order: ((a0 + brdf0*b0) + a1) + brdf1*b1 ...
// Rays live in shared memory between the pieces (16 words per ray, [ray][word][thread] so a warp's accesses are conflict-free).
constexpr int RAY_WORDS = 16;
constexpr int TRACE_THREADS = 128;
// words: 0-2 dir, 3-5 march position, 6 NdotD, 7-9 previous voxel; after the march: 0 hit texel, 1 "ran out of steps", 3-5 final
// position; 10-12 radiance of the primary ray, 13-15 its brdf (kept across the secondary march)
#define RS(k, f) S[((k) * RAY_WORDS + (f)) * TRACE_THREADS + threadIdx.x]

// rands.x / rands.y, indirect.frag:111-116
__device__ __forceinline__ void ray_rands(const TraceParams& P, float seed, float uvx, float uvy, const float* ext, float& rx, float& ry)
{
    if (ext) { rx = ext[0]; ry = ext[1]; return; }
    const float t0 = 0.07f * dm_fract(P.iiTime), t1 = 0.11f * dm_fract(P.iiTime + 0.573953f);
    const float ra = glsl_hash(seed, seed);
    const float ax = (uvx + ra) * P.resx, ay = (uvy + ra) * P.resy;
    rx = blugausnoise2(-ax, -ay, t0, t1);
    ry = blugausnoise2(ax, ay, t0, t1);
}

// indirect.frag:117-135: direction in the tangent frame of wnorm, first march position, starting voxel -> ray k of this thread
__device__ void ray_setup(const TraceParams& P, float* S, int k, f3 wpos, f3 wnorm, float rx, float ry)
{
    // hemisphereSample_cos, math.inc:76-81
    const float phi = ry * 2.0f * 3.1415926f;
    const float cosTheta = __fsqrt_rn(1.0f - rx);
    const float sinTheta = __fsqrt_rn(1.0f - cosTheta * cosTheta);
    const f3 d = {dm_cos(phi) * sinTheta, dm_sin(phi) * sinTheta, cosTheta};
    // make_coord_space, indirect.frag:71-86
    f3 z = wnorm, h = wnorm;
    if (fabsf(h.x) <= fabsf(h.y) && fabsf(h.x) <= fabsf(h.z)) h.x = 1.0f;
    else if (fabsf(h.y) <= fabsf(h.x) && fabsf(h.y) <= fabsf(h.z)) h.y = 1.0f;
    else h.z = 1.0f;
    z = normalize3(z);
    const f3 y = normalize3(cross3(h, z));
    const f3 x = normalize3(cross3(z, y));
    f3 dir = {(x.x * d.x + y.x * d.y) + z.x * d.z, (x.y * d.x + y.y * d.y) + z.y * d.z, (x.z * d.x + y.z * d.y) + z.z * d.z};
    if (dot3(dir, wnorm) < 0.0f) dir = neg3(dir);
    const float NdotD = dot3(dir, wnorm);
    const float s1 = 1.0f + ry, step_size = P.step_size;
    const f3 march_pos = {wpos.x + dir.x * s1 * step_size / NdotD, wpos.y + dir.y * s1 * step_size / NdotD, wpos.z + dir.z * s1 * step_size / NdotD};
    const f3 sv = voxel_pos(P.w2voxel, wpos, (float)P.N);
    RS(k, 0) = dir.x; RS(k, 1) = dir.y; RS(k, 2) = dir.z;
    RS(k, 3) = march_pos.x; RS(k, 4) = march_pos.y; RS(k, 5) = march_pos.z;
    RS(k, 6) = NdotD;
    RS(k, 7) = __int_as_float(dm_f2i(sv.x)); RS(k, 8) = __int_as_float(dm_f2i(sv.y)); RS(k, 9) = __int_as_float(dm_f2i(sv.z));
}

// indirect.frag:137-152, 176.  The block's 4 x TRACE_THREADS rays form a pool: a lane takes the next unmarched ray (a shared
// cursor), walks it to its end, writes the outcome into that ray's slots and takes another — whichever pixel it belongs to.
// Walking only its own pixel's rays left a warp waiting for its unluckiest lane (118 iterations against a mean of 44, ncu);
// any lane can march any ray because a ray's arithmetic does not depend on who executes it.
#define RSO(owner, k, f) S[((k) * RAY_WORDS + (f)) * TRACE_THREADS + (owner)]
__device__ void march_pool(const TraceParams& P, float* S, const unsigned int* need, unsigned int* cursor, unsigned int& steps_taken)
{
    const float Nf = (float)P.N, hi = (float)(P.N - 1), step_size = P.step_size;
    const int N = (int)P.N;
    int k = 0, owner = 0, pvx = 0, pvy = 0, pvz = 0;
    uint32_t i = 0;
    f3 dir = {0.f, 0.f, 0.f}, pos = {0.f, 0.f, 0.f};
    bool have = false;
    auto next_ray = [&]() {
        have = false;
        for (;;)
        {
            const unsigned int j = atomicAdd(cursor, 1u);
            if (j >= 4u * TRACE_THREADS) break;
            owner = (int)(j % TRACE_THREADS); k = (int)(j / TRACE_THREADS);
            if (!((need[owner] >> k) & 1u)) continue;
            dir = {RSO(owner, k, 0), RSO(owner, k, 1), RSO(owner, k, 2)};
            pos = {RSO(owner, k, 3), RSO(owner, k, 4), RSO(owner, k, 5)};
            pvx = __float_as_int(RSO(owner, k, 7)); pvy = __float_as_int(RSO(owner, k, 8)); pvz = __float_as_int(RSO(owner, k, 9));
            i = 0;
            bool idle = (P.steps == 0);
            // ~15 % of the reference's rays have a NaN direction (blugausnoise2 leaves [0, 1]).  Under the pinned
            // NaN rules such a ray never leaves the loop — every comparison is false, ivec3(NaN) = (0,0,0): it reads voxel (0,0,0)
            // at most once, spins for all `steps` iterations, and its sky term is x * smoothstep(NaN) = 0.  Same outcome without
            // the iterations: no hit, ran out, `steps` counted.  (If voxel (0,0,0) is occupied the loop below handles the ray.)
            if (!idle && dm_isnan(dir.x))
            {
                uint32_t texel0 = 0;
                if ((pvx | pvy | pvz) != 0) texel0 = __ldg(P.vox);
                if ((texel0 & 0xffffu) == 0) { steps_taken += P.steps; idle = true; }
            }
            if (idle) { RSO(owner, k, 0) = __uint_as_float(0u); RSO(owner, k, 1) = __uint_as_float(1u); continue; }
            have = true;
            break;
        }
    };
    next_ray();
    while (have)
    {
        steps_taken++;
        pos = {pos.x + dir.x * step_size, pos.y + dir.y * step_size, pos.z + dir.z * step_size};
        const f3 vp = voxel_pos(P.w2voxel, pos, Nf);
        bool finished = false, ranout = false;
        uint32_t hit_texel = 0;
        if (vp.x < 0.f || vp.y < 0.f || vp.z < 0.f || vp.x > hi || vp.y > hi || vp.z > hi) finished = true;
        else
        {
            const int ix = dm_f2i(vp.x), iy = dm_f2i(vp.y), iz = dm_f2i(vp.z);       // NaN -> 0
            if (pvx != ix || pvy != iy || pvz != iz)
            {
                uint32_t texel = 0;
                if ((ix | iy | iz) >= 0 && ix < N && iy < N && iz < N) texel = __ldg(P.vox + ((uint32_t)(iz * N + iy) * (uint32_t)N + (uint32_t)ix));
                pvx = ix; pvy = iy; pvz = iz;
                if ((texel & 0xffffu) != 0) { hit_texel = texel; finished = true; }
            }
        }
        if (!finished && ++i == P.steps) { finished = true; ranout = true; }
        if (finished)
        {
            RSO(owner, k, 0) = __uint_as_float(hit_texel); RSO(owner, k, 1) = __uint_as_float(ranout ? 1u : 0u);
            RSO(owner, k, 3) = pos.x; RSO(owner, k, 4) = pos.y; RSO(owner, k, 5) = pos.z;
            next_ray();
        }
    }
}

// indirect.frag:154-181 for ray k after its march: first-bounce lighting at the hit, or the sky term
__device__ f3 ray_shade(const TraceParams& P, const float* S, int k, Hit& hit)
{
    hit.hit = false;
    f3 Lo = {0.f, 0.f, 0.f};
    const uint32_t hit_texel = __float_as_uint(RS(k, 0));
    const bool ranout = __float_as_uint(RS(k, 1)) != 0u;
    const float NdotD = RS(k, 6);
    if (hit_texel != 0)
    {
        const f3 march_pos = {RS(k, 3), RS(k, 4), RS(k, 5)};
        const uint32_t r = hit_texel & 0xffffu, g = hit_texel >> 16;
        f3 col = {dm_pow((float)((r & 0xF800u) >> 11) / 31.0f, 2.2f), dm_pow((float)((r & 0x7E0u) >> 5) / 63.0f, 2.2f),
                  dm_pow((float)(r & 0x1Fu) / 31.0f, 2.2f)};
        f3 vn = normalize3(f3{(float)(g & 0x1Fu) / 16.0f - 1.0f, (float)((g & 0x7E0u) >> 5) / 32.0f - 1.0f,
                              (float)((g & 0xF800u) >> 11) / 16.0f - 1.0f});
        f3 sp = {march_pos.x + vn.x * 0.06f, march_pos.y + vn.y * 0.06f, march_pos.z + vn.z * 0.06f};
        f4 sv4 = mul44(P.ShadowProj, mul44(P.ShadowView, f4{sp.x, sp.y, sp.z, 1.0f}));
        float spx = sv4.x / sv4.w, spy = sv4.y / sv4.w, spz = sv4.z / sv4.w;
        spx = spx * 0.5f + 0.5f; spy = spy * 0.5f + 0.5f;
        const int tx = dm_f2i(spx * (float)P.S), ty = dm_f2i(spy * (float)P.S);
        float shadowZ = 0.0f;
        if (tx >= 0 && ty >= 0 && tx < (int)P.S && ty < (int)P.S) shadowZ = __ldg(P.shadow + (size_t)ty * P.S + tx);
        const float shade = dm_step(spz + 0.005f, shadowZ);
        const float l = fabsf(dot3(neg3(P.sunPos), vn));
        const float den = dm_max(0.01f, NdotD);
        f3 rr = {l * col.x / den, l * col.y / den, l * col.z / den};
        hit.brdf = rr;
        Lo = {Lo.x + P.sunLum.x * shade * rr.x, Lo.y + P.sunLum.y * shade * rr.y, Lo.z + P.sunLum.z * shade * rr.z};
        hit.wpos = march_pos; hit.wnorm = vn; hit.hit = true;
    }
    if (!hit.hit && ranout)
    {
        const float den = dm_max(0.01f, NdotD);
        const float sm = dm_smoothstep(0.0f, 0.01f, NdotD);
        Lo = {Lo.x + 0.7f * 0.4f / den * sm, Lo.y + 0.8f * 0.4f / den * sm, Lo.z + 1.0f * 0.4f / den * sm};
    }
    return Lo;
}

__device__ __forceinline__ float unorm16(uint16_t v) { return (float)v / 65535.0f; }
__device__ __forceinline__ int wrapn(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

// 128 threads = 2x2 warps of 8x4 pixels -> a 16x8 pixel tile per block
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_r(const TraceParams P, unsigned long long* __restrict__ step_counter)
{
    extern __shared__ float S[];          // 4 rays x RAY_WORDS x TRACE_THREADS
    __shared__ unsigned int need[TRACE_THREADS];
    __shared__ unsigned int cursor[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t y = P.y0 + (P.tile0 + blockIdx.y * P.tile_stride) * 8 + (warp >> 1) * 4 + (lane >> 3);
    const uint32_t W = P.W, H = P.H;
    const bool inside = x < P.W && y < P.y1;      // threads off the image own no rays but still march the block's pool
    unsigned int steps_taken = 0;
    float uvx = 0.f, uvy = 0.f;
    f3 cspos = {0.f, 0.f, 0.f}, wpos = {0.f, 0.f, 0.f};
    const float* er = nullptr;
    if (threadIdx.x == 0) { cursor[0] = 0u; cursor[1] = 0u; }
    if (inside)
    {
        uvx = ((float)x + 0.5f) / (float)W; uvy = ((float)y + 0.5f) / (float)H;
        // getCSpos, indirect.frag:44-53
        const int dx = dm_f2i(uvx * (float)W), dy = dm_f2i(uvy * (float)H);
        const float depth = (dx >= 0 && dy >= 0 && dx < (int)W && dy < (int)H) ? __ldg(P.depth + (size_t)dy * W + dx) : 0.0f;
        f4 cp = mul44(P.InvProj, f4{uvx * 2.0f - 1.0f, uvy * 2.0f - 1.0f, depth, 1.0f});
        cspos = {cp.x / cp.w, cp.y / cp.w, cp.z / cp.w};
        wpos = mul43(P.InvModelView, cspos, 1.0f);
        // getNormal, indirect.frag:55-58 (texel-centre fetch)
        const ushort4 nq = __ldg(reinterpret_cast<const ushort4*>(P.normals) + (size_t)y * W + x);
        const f3 raw = {fmaf(unorm16(nq.x), 2.0f, -1.0f), fmaf(unorm16(nq.y), 2.0f, -1.0f), fmaf(unorm16(nq.z), 2.0f, -1.0f)};
        const f3 csnorm = normalize3(raw);
        const f3 wnorm = mul33(P.InvModelView, csnorm);
        er = P.rands ? P.rands + 16 * ((size_t)y * W + x) : nullptr;
        // the 4 primary rays: same origin and normal, seeds 0, 2, 4, 6 (indirect.frag:201-221)
#pragma unroll 1
        for (int k = 0; k < 4; k++)
        {
            float rx, ry;
            ray_rands(P, (float)(2 * k), uvx, uvy, er ? er + 4 * k : nullptr, rx, ry);
            ray_setup(P, S, k, wpos, wnorm, rx, ry);
        }
    }
    need[threadIdx.x] = inside ? 0xfu : 0u;
    __syncthreads();
    march_pool(P, S, need, &cursor[0], steps_taken);
    __syncthreads();
    // shade them; every hit spawns the secondary ray of its pair (seeds 1, 3, 5, 7) from the hit point along the voxel normal
    unsigned int second = 0;
    if (inside)
    {
#pragma unroll 1
        for (int k = 0; k < 4; k++)
        {
            Hit st;
            const f3 a = ray_shade(P, S, k, st);
            RS(k, 10) = a.x; RS(k, 11) = a.y; RS(k, 12) = a.z;
            if (st.hit)
            {
                RS(k, 13) = st.brdf.x; RS(k, 14) = st.brdf.y; RS(k, 15) = st.brdf.z;     // read before the second call (indirect.frag:204)
                float rx, ry;
                ray_rands(P, (float)(2 * k + 1), uvx, uvy, er ? er + 4 * k + 2 : nullptr, rx, ry);
                ray_setup(P, S, k, st.wpos, st.wnorm, rx, ry);
                second |= 1u << k;
            }
        }
    }
    need[threadIdx.x] = second;
    __syncthreads();
    march_pool(P, S, need, &cursor[1], steps_taken);
    __syncthreads();
    if (inside)
    {
        f3 ind = {0.f, 0.f, 0.f};
#pragma unroll 1
        for (int k = 0; k < 4; k++)
        {
            ind = ind + f3{RS(k, 10), RS(k, 11), RS(k, 12)};
            if (second & (1u << k))
            {
                Hit st;
                const f3 b = ray_shade(P, S, k, st);
                ind = ind + f3{RS(k, 13), RS(k, 14), RS(k, 15)} * b;
            }
        }
        ind = ind * 0.25f;
        // temporal reprojection, indirect.frag:225-240
        f4 pc = mul44(P.prevModelView, f4{wpos.x, wpos.y, wpos.z, 1.0f});
        f4 pp = mul44(P.prevProjection, pc);
        float ru = pp.x / pp.w, rv = pp.y / pp.w;
        ru = ru * 0.5f + 0.5f; rv = rv * 0.5f + 0.5f;
        if (dm_clamp(ru, 0.0f, 1.0f) == ru && dm_clamp(rv, 0.0f, 1.0f) == rv)
        {
            const float fx = ru * (float)W - 0.5f, fy = rv * (float)H - 0.5f;
            const float x0f = floorf(fx), y0f = floorf(fy);
            const float wx = fx - x0f, wy = fy - y0f;
            const int xi0 = dm_f2i(x0f), yi0 = dm_f2i(y0f);
            const int xa = wrapn(xi0, (int)W), xb = wrapn(xi0 + 1, (int)W), ya = wrapn(yi0, (int)H), yb = wrapn(yi0 + 1, (int)H);
            const ushort4 h00 = __ldg(reinterpret_cast<const ushort4*>(P.hist) + (size_t)ya * W + xa);
            const ushort4 h10 = __ldg(reinterpret_cast<const ushort4*>(P.hist) + (size_t)ya * W + xb);
            const ushort4 h01 = __ldg(reinterpret_cast<const ushort4*>(P.hist) + (size_t)yb * W + xa);
            const ushort4 h11 = __ldg(reinterpret_cast<const ushort4*>(P.hist) + (size_t)yb * W + xb);
            auto bil = [&](uint16_t a, uint16_t b, uint16_t c, uint16_t d) {
                const float fa = dm_f16_to_f32(a), fb = dm_f16_to_f32(b), fc = dm_f16_to_f32(c), fd = dm_f16_to_f32(d);
                return (fa * (1.0f - wx) + fb * wx) * (1.0f - wy) + (fc * (1.0f - wx) + fd * wx) * wy;
            };
            const float p0 = bil(h00.x, h10.x, h01.x, h11.x), p1 = bil(h00.y, h10.y, h01.y, h11.y);
            const float p2 = bil(h00.z, h10.z, h01.z, h11.z), p3 = bil(h00.w, h10.w, h01.w, h11.w);
            const float bw = 0.95f * dm_smoothstep(0.0f, 1.0f, 1.0f - fabsf(p3 + cspos.z));
            ind = {dm_clamp(dm_mix(ind.x, p0, bw), 0.0f, 16.0f), dm_clamp(dm_mix(ind.y, p1, bw), 0.0f, 16.0f),
                   dm_clamp(dm_mix(ind.z, p2, bw), 0.0f, 16.0f)};
        }
        ushort4 o = make_ushort4(dm_f32_to_f16(ind.x), dm_f32_to_f16(ind.y), dm_f32_to_f16(ind.z), dm_f32_to_f16(-cspos.z));
        reinterpret_cast<ushort4*>(P.out)[(size_t)y * W + x] = o;
    }
    warp_count_add(step_counter, steps_taken);
}

}  // namespace

int f184_trace_r(f184_ctx* c, const f184_trace_constants* k)
{
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_SHADOW, F184_SLOT_VOXELS, F184_SLOT_INDIRECT_OUT, F184_SLOT_INDIRECT_HISTORY})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    TraceParams P{};
    memcpy(P.InvProj.m, k->view.InvProj, 64);
    memcpy(P.InvModelView.m, k->ext.InvModelView, 64);
    memcpy(P.ShadowView.m, k->ext.ShadowView, 64);
    memcpy(P.ShadowProj.m, k->ext.ShadowProj, 64);
    M4 vp, vv;
    memcpy(vp.m, k->ext.VoxelProj, 64);
    memcpy(vv.m, k->ext.VoxelView, 64);
    P.w2voxel = host_matmul(vp, vv);                                  // indirect.frag:127
    memcpy(P.prevModelView.m, k->prev.PrevModelView, 64);
    memcpy(P.prevProjection.m, k->prev.PrevProjection, 64);
    P.depth = img_ptr<float>(c, F184_SLOT_DEPTH);
    P.normals = img_ptr<uint16_t>(c, F184_SLOT_NORMALS);
    P.shadow = img_ptr<float>(c, F184_SLOT_SHADOW);
    P.vox = img_ptr<uint32_t>(c, F184_SLOT_VOXELS);
    P.hist = img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_HISTORY);
    P.out = img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_OUT);
    P.rands = (c->cfg.flags & F184_FLAG_EXTERNAL_RANDS) ? c->rands : nullptr;
    P.W = c->cfg.width; P.H = c->cfg.height; P.S = c->cfg.shadow_res; P.N = c->cfg.grid_n; P.steps = c->cfg.march_steps;
    const uint32_t grid_y = f184_trace_tiles(c, P.H, &P.y0, &P.y1, &P.tile0, &P.tile_stride);
    if (P.tile_stride > 1 && (P.y0 & 7)) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "trace: row range must start on a multiple of 8 when tiles are interleaved");
    P.step_size = c->cfg.step_size;
    P.iiTime = (float)k->miscs.frameCount * 0.03125f;                 // indirect.frag:111
    P.resx = k->miscs.resolution[0]; P.resy = k->miscs.resolution[1];
    P.sunLum = {k->sun.luminance[0], k->sun.luminance[1], k->sun.luminance[2]};
    P.sunPos = {k->sun.position[0], k->sun.position[1], k->sun.position[2]};

    int rc = f184_stage_begin(c, F184_STAGE_TRACE);
    if (rc) return rc;
    if (k->reset_history)     // first frame: history cleared (MegaPipeline.cpp:197-204)
        CK(c, cudaMemsetAsync(c->img[F184_SLOT_INDIRECT_HISTORY].ptr, 0, c->img[F184_SLOT_INDIRECT_HISTORY].desc.size_bytes, c->stream));
    CK(c, cudaMemsetAsync(c->counters_dev + F184_COUNTER_MARCH_STEPS, 0, 8, c->stream));
    if (grid_y)
    {
        dim3 grid((P.W + 15) / 16, grid_y);
        const size_t smem = 4 * RAY_WORDS * TRACE_THREADS * sizeof(float);       // 32 KB
        k_trace_r<<<grid, TRACE_THREADS, smem, c->stream>>>(P, c->counters_dev + F184_COUNTER_MARCH_STEPS);
        CK_LAUNCH(c);
    }
    return f184_stage_end(c, F184_STAGE_TRACE);
}
