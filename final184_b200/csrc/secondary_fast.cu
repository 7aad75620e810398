// secondary_fast.cu — the secondary passes of the north-star contract: GTAO (+ its 4x4 blur) and the separable bilateral blur of
// the indirect image, tolerance-checked instead of bit-pinned.
//
// Same passes, same definitions as gtao.cu / blur.cu (Shader/GTAO/gtao.frag:48-120, GTAO/blur.frag:12-27,
// Shader/Lighting/bilateralBlur.inc; Foreground/Renderer/MegaPipeline.cpp:225-239, 270-284).  Those two files hold the
// reference-faithful contract: IEEE divisions and square roots, the pinned dm_sin / dm_cos / dm_exp2 polynomials, -fmad=false —
// every pixel equal to the shader text bit for bit, and ALU-bound on exactly that arithmetic (1.84 ms + 0.63 ms at 3840 x 2160 for
// 20 B/pixel of compulsory traffic).  The north-star contract asks for 1e-2 relative L2 per image (BASELINE.json), which leaves
// room for the hardware's own units: MUFU reciprocal / rsqrt / sin / cos / ex2 and fused multiply-adds.  This translation unit is
// compiled with fmad on and selected for F184_MODE_NORTHSTAR unless F184_FLAG_EXACT_SECONDARY asks for the pinned kernels.
#include "f184_device.cuh"

namespace {

constexpr float PI_ = 3.1415926f;
constexpr float half_PI_ = 3.1415926f / 2.0f;

__device__ __forceinline__ float rcp_fast(float x) { return __fdividef(1.0f, x); }
__device__ __forceinline__ float ex2_fast(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fastSqrt(float x) { return __int_as_float(0x1FBD1DF5 + (__float_as_int(x) >> 1)); }   // math.inc:14-20
__device__ __forceinline__ float fastAcos(float x)                                                                      // math.inc:22-33
{
    float res = fmaf(-0.156583f, fabsf(x), half_PI_);
    res *= fastSqrt(1.0f - fabsf(x));
    return x >= 0.0f ? res : PI_ - res;
}
__device__ __forceinline__ f3 norm_fast(f3 a)
{
    const float r = rsqrtf(fmaf(a.x, a.x, fmaf(a.y, a.y, a.z * a.z)));
    return {a.x * r, a.y * r, a.z * r};
}
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

struct GtaoFastParams
{
    M4 InvProj;
    const float* depth;
    const uint16_t* normals;
    uint16_t* raw;
    uint16_t* out;
    int W, H;
    float invW, invH;
};

__device__ __forceinline__ f3 cs_pos(const GtaoFastParams& G, float u, float v)
{
    const int dx = (int)(u * (float)G.W), dy = (int)(v * (float)G.H);
    const float depth = (u >= 0.0f && v >= 0.0f && dx < G.W && dy < G.H) ? __ldg(G.depth + (size_t)dy * G.W + dx) : 0.0f;
    const float nx = fmaf(u, 2.0f, -1.0f), ny = fmaf(v, 2.0f, -1.0f);
    const M4& M = G.InvProj;
    const float px = fmaf(M.m[0], nx, fmaf(M.m[4], ny, fmaf(M.m[8], depth, M.m[12])));
    const float py = fmaf(M.m[1], nx, fmaf(M.m[5], ny, fmaf(M.m[9], depth, M.m[13])));
    const float pz = fmaf(M.m[2], nx, fmaf(M.m[6], ny, fmaf(M.m[10], depth, M.m[14])));
    const float pw = fmaf(M.m[3], nx, fmaf(M.m[7], ny, fmaf(M.m[11], depth, M.m[15])));
    const float rw = rcp_fast(pw);
    return {px * rw, py * rw, pz * rw};
}

// gtao.frag:48-120: 4 slices x 2 steps x 2 sides
__global__ void __launch_bounds__(128) k_gtao_fast(const GtaoFastParams G)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= G.W || y >= G.H) return;
    const float u = ((float)x + 0.5f) * G.invW, v = ((float)y + 0.5f) * G.invH;
    const f3 cur = cs_pos(G, u, v);
    float vis = 1.0f;
    if (-cur.z <= 32.0f)
    {
        const f3 Vv = neg3(norm_fast(cur));
        const ushort4 nq = __ldg(reinterpret_cast<const ushort4*>(G.normals) + (size_t)y * G.W + x);
        const float s16 = 2.0f / 65535.0f;
        const f3 cn = norm_fast(f3{fmaf((float)nq.x, s16, -1.0f), fmaf((float)nq.y, s16, -1.0f), fmaf((float)nq.z, s16, -1.0f)});
        const float radius = (float)G.H * 0.5f * rcp_fast(-cur.z);
        const float rStep = radius * 0.5f;
        float phi = -(1.0f / 16.0f) * (float)((((x + y) & 0x3) << 2) + (x & 0x3)) * PI_;
        const float r0 = rStep * (0.25f * (float)((y - x) & 0x3));
        // On the diagonal of the 4x4 interleave r0 = 0: the first tap of BOTH sides is the pixel itself, normalize(0) is NaN, the NaN
        // survives the horizon blend (mix(..)*0 + ..), and the reference's max(NaN, -pi/2) / min(NaN, pi/2) then open the horizon
        // completely (min/max return the non-NaN operand).  Said outright here instead of left to NaN propagation
        // through approximate units: those pixels need no depth taps at all.
        const bool open = ((y - x) & 0x3) == 0;
        float integral = 0.0f;
#pragma unroll 1
        for (int samp = 0; samp < 4; samp++)
        {
            float sph, cph;
            __sincosf(phi, &sph, &cph);
            phi += PI_ / 4.0f;
            float hx = -1.0f, hy = -1.0f;
            float r = r0;
#pragma unroll
            for (int j = 0; j < 2 && !open; j++)
            {
                const float ox = r * cph * G.invW, oy = -r * sph * G.invH;
                r += rStep;
                const float u1 = u - ox, v1 = v - oy, u2 = u + ox, v2 = v + oy;
                const f3 ds = cs_pos(G, u1, v1) - cur, dt = cs_pos(G, u2, v2) - cur;
                float hsx = dot3(Vv, norm_fast(ds)), hsy = dot3(Vv, norm_fast(dt));
                if (u1 < 0.0f || u1 > 1.0f || v1 < 0.0f || v1 > 1.0f) hsx = -1.0f;
                if (u2 < 0.0f || u2 > 1.0f || v2 < 0.0f || v2 > 1.0f) hsy = -1.0f;
                hx = hx >= hsx ? hx : 0.5f * (hx + hsx);          // mix(mix(h, hs, .5), max(h, hs), step(hs, h))   gtao.frag:90-93
                hy = hy >= hsy ? hy : 0.5f * (hy + hsy);
            }
            const f3 sliceDir = {cph, sph, 0.0f};
            const f3 sliceNormal = norm_fast(cross3(Vv, sliceDir));
            const f3 sliceBitangent = norm_fast(cross3(sliceNormal, Vv));
            const float dn = dot3(cn, sliceNormal);
            f3 projNorm = {cn.x - sliceNormal.x * dn, cn.y - sliceNormal.y * dn, cn.z - sliceNormal.z * dn};
            const float weight = sqrtf(dot3(projNorm, projNorm)) + 1e-6f;
            const float rwgt = rcp_fast(weight);
            projNorm = {projNorm.x * rwgt, projNorm.y * rwgt, projNorm.z * rwgt};
            const float cosn = dot3(projNorm, Vv), sinn = dot3(projNorm, sliceBitangent);
            const float n = fastAcos(cosn) * (sinn > 0.0f ? 1.0f : (sinn < 0.0f ? -1.0f : 0.0f));
            hx = open ? n - half_PI_ : n + fmaxf(-fastAcos(hx) - n, -half_PI_);
            hy = open ? n + half_PI_ : n + fminf(fastAcos(hy) - n, half_PI_);
            float sn, cn_;
            __sincosf(n, &sn, &cn_);
            const float ax = -__cosf(2.0f * hx - n) + cn_ + 2.0f * hx * sn;
            const float ay = -__cosf(2.0f * hy - n) + cn_ + 2.0f * hy * sn;
            integral = fmaf(0.25f * (ax + ay), weight, integral);
        }
        vis = integral * 0.25f;
    }
    reinterpret_cast<ushort4*>(G.raw)[(size_t)y * G.W + x] = make_ushort4(f2h(vis), 0, 0, f2h(1.0f));
}

// GTAO/blur.frag:12-27: mean of the 4x4 neighbourhood [x-1, x+2] x [y-1, y+2] (four gathers), repeat addressing
__global__ void __launch_bounds__(128) k_gtao_blur_fast(const GtaoFastParams G)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= G.W || y >= G.H) return;
    float sum = 0.0f;
#pragma unroll
    for (int j = -1; j <= 2; j++)
    {
        int yy = y + j; yy = yy < 0 ? yy + G.H : (yy >= G.H ? yy - G.H : yy);
#pragma unroll
        for (int i = -1; i <= 2; i++)
        {
            int xx = x + i; xx = xx < 0 ? xx + G.W : (xx >= G.W ? xx - G.W : xx);
            sum += h2f(__ldg(G.raw + 4 * ((size_t)yy * G.W + xx)));
        }
    }
    const uint16_t hv = f2h(sum * (1.0f / 16.0f));
    reinterpret_cast<ushort4*>(G.out)[(size_t)y * G.W + x] = make_ushort4(hv, hv, hv, f2h(1.0f));
}

struct BlurFastParams { const uint16_t* src; const float* depth; uint16_t* dst; int W, H, dirx, diry; };

// bilateralBlur.inc: centre + 12 taps at pixel offsets +-2, 4, 6, 8, 11, 15 with r = 1, 2, 3, 4, 5, 7
__global__ void __launch_bounds__(128) k_blur_fast(const BlurFastParams B)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= B.W || y >= B.H) return;
    const ushort4 p0 = __ldg(reinterpret_cast<const ushort4*>(B.src) + (size_t)y * B.W + x);
    float t0 = h2f(p0.x), t1 = h2f(p0.y), t2 = h2f(p0.z), tw = 1.0f;
    const float cz = __ldg(B.depth + (size_t)y * B.W + x);
    const int offs[6] = {2, 4, 6, 8, 11, 15};
    const float rr[6] = {1.0f, 2.0f, 3.0f, 4.0f, 5.0f, 7.0f};
#pragma unroll
    for (int k = 0; k < 6; k++)
#pragma unroll
        for (int sgn = -1; sgn <= 1; sgn += 2)
        {
            const int xi = min(max(x + sgn * offs[k] * B.dirx, 0), B.W - 1), yi = min(max(y + sgn * offs[k] * B.diry, 0), B.H - 1);
            const ushort4 p = __ldg(reinterpret_cast<const ushort4*>(B.src) + (size_t)yi * B.W + xi);
            const float dz = (cz - __ldg(B.depth + (size_t)yi * B.W + xi)) * 512.0f;
            const float w = ex2_fast(fmaf(-dz, dz, -rr[k] * rr[k] * (1.0f / 32.0f)));
            t0 = fmaf(h2f(p.x), w, t0); t1 = fmaf(h2f(p.y), w, t1); t2 = fmaf(h2f(p.z), w, t2);
            tw += w;
        }
    const float r = rcp_fast(tw);
    reinterpret_cast<ushort4*>(B.dst)[(size_t)y * B.W + x] = make_ushort4(f2h(t0 * r), f2h(t1 * r), f2h(t2 * r), 0);
}

}  // namespace

int f184_gtao_fast_impl(f184_ctx* c, const f184_view_constants* view)
{
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_AO_RAW, F184_SLOT_AO_OUT})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    GtaoFastParams G{};
    memcpy(G.InvProj.m, view->InvProj, 64);
    G.depth = img_ptr<float>(c, F184_SLOT_DEPTH);
    G.normals = img_ptr<uint16_t>(c, F184_SLOT_NORMALS);
    G.raw = img_ptr<uint16_t>(c, F184_SLOT_AO_RAW);
    G.out = img_ptr<uint16_t>(c, F184_SLOT_AO_OUT);
    G.W = (int)c->cfg.width; G.H = (int)c->cfg.height;
    G.invW = 1.0f / (float)G.W; G.invH = 1.0f / (float)G.H;
    int rc = f184_stage_begin(c, F184_STAGE_GTAO);
    if (rc) return rc;
    dim3 grid((G.W + 15) / 16, (G.H + 7) / 8);
    k_gtao_fast<<<grid, 128, 0, c->stream>>>(G);
    CK_LAUNCH(c);
    k_gtao_blur_fast<<<grid, 128, 0, c->stream>>>(G);
    CK_LAUNCH(c);
    return f184_stage_end(c, F184_STAGE_GTAO);
}

int f184_blur_fast_impl(f184_ctx* c, const f184_engine_miscs*)
{
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_INDIRECT_OUT, F184_SLOT_INDIRECT_BLUR_X, F184_SLOT_INDIRECT_FINAL})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    const int W = (int)c->cfg.width, H = (int)c->cfg.height;
    int rc = f184_stage_begin(c, F184_STAGE_BLUR);
    if (rc) return rc;
    dim3 grid((W + 15) / 16, (H + 7) / 8);
    // the pass the reference names indirect_blurX steps along y, indirect_blurY along x (blurX.frag:5, blurY.frag:5)
    BlurFastParams bx{img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_OUT), img_ptr<float>(c, F184_SLOT_DEPTH), img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_BLUR_X), W, H, 0, 1};
    k_blur_fast<<<grid, 128, 0, c->stream>>>(bx);
    CK_LAUNCH(c);
    BlurFastParams by{img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_BLUR_X), img_ptr<float>(c, F184_SLOT_DEPTH), img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_FINAL), W, H, 1, 0};
    k_blur_fast<<<grid, 128, 0, c->stream>>>(by);
    CK_LAUNCH(c);
    return f184_stage_end(c, F184_STAGE_BLUR);
}
