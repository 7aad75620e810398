// gtao.cu — ground-truth ambient occlusion, the path's secondary pass.
//
// Replaces gtao_visibility + gtao_blur (Foreground/Renderer/MegaPipeline.cpp:225-239):
//   Shader/GTAO/gtao.frag:48-120  per pixel 4 slice directions x 2 steps x 2 sides horizon search on the
//                                 depth buffer, cosine-weighted arc integral, fastAcos/fastSqrt of
//                                 Shader/math.inc:14-33
//   Shader/GTAO/blur.frag:12-27   mean of the 4x4 neighbourhood (four textureGatherOffset footprints)
// B200 design: 17 depth taps per pixel land within `radius` pixels of the centre, so a warp's 8x4 tile
// re-reads the same lines from L1; the two passes stay separate kernels because the blur reads the
// visibility AFTER its rounding to the RGBA16F target (parity), and both are far below 0.1 ms.
#include "f184_device.cuh"

namespace {

constexpr float PI_ = 3.1415926f;
constexpr float half_PI_ = 3.1415926f / 2.0f;

__device__ __forceinline__ float fastSqrt(float x) { return __int_as_float(0x1FBD1DF5 + (__float_as_int(x) >> 1)); }
__device__ __forceinline__ float fastAcos1(float x)
{
    float res = -0.156583f * fabsf(x) + half_PI_;
    res *= fastSqrt(1.0f - fabsf(x));
    return x >= 0 ? res : PI_ - res;
}
__device__ __forceinline__ float fastAcos2(float x)
{
    float res = -0.156583f * fabsf(x) + half_PI_;
    res *= fastSqrt(1.0f - fabsf(x));
    float flag = dm_step(x, 0.0f);
    return res * fmaf(-flag, 2.0f, 1.0f) + flag * PI_;
}
__device__ __forceinline__ float unorm16(uint16_t v) { return (float)v / 65535.0f; }
__device__ __forceinline__ float sgn(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

struct GtaoParams { M4 InvProj; const float* depth; const uint16_t* normals; uint16_t* raw; uint16_t* out; uint32_t W, H; const float2* phi_table; };

// The slice angle takes 16 start values (4x4 interleave, gtao.frag:64) x 4 slices: phi = -(k/16) pi, then += pi/4 three times.
// (cos, sin) of those 64 angles, evaluated once on the device with the same float additions and the same dm_cos / dm_sin the
// per-pixel code used, so the table is bit-identical to recomputing them (8 of the ~24 range reductions per pixel).
__global__ void k_gtao_phi_table(float2* __restrict__ table)
{
    const int k = threadIdx.x;               // 0..15
    float phi = -(1.0f / 16.0f) * (float)k * PI_;
    for (int samp = 0; samp < 4; samp++)
    {
        table[k * 4 + samp] = make_float2(dm_cos(phi), dm_sin(phi));
        phi += PI_ / 4.0f;
    }
}

__device__ __forceinline__ f3 cs_pos(const GtaoParams& G, float u, float v)
{
    const int dx = dm_f2i(u * (float)G.W), dy = dm_f2i(v * (float)G.H);
    const float depth = (dx >= 0 && dy >= 0 && dx < (int)G.W && dy < (int)G.H) ? __ldg(G.depth + (size_t)dy * G.W + dx) : 0.0f;
    f4 p = mul44(G.InvProj, f4{u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f});
    return {p.x / p.w, p.y / p.w, p.z / p.w};
}

__global__ void __launch_bounds__(128) k_gtao(const GtaoParams G)
{
    __shared__ float2 phi_cs[64];
    if (threadIdx.x < 64) phi_cs[threadIdx.x] = G.phi_table[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= G.W || y >= G.H) return;
    const uint32_t W = G.W, H = G.H;
    const float u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;
    const f3 cur = cs_pos(G, u, v);
    const f3 Vv = neg3(normalize3(cur));
    float vis;
    if (-cur.z > 32.0f) vis = 1.0f;
    else
    {
        const ushort4 nq = __ldg(reinterpret_cast<const ushort4*>(G.normals) + (size_t)y * W + x);
        const f3 cn = normalize3(f3{fmaf(unorm16(nq.x), 2.0f, -1.0f), fmaf(unorm16(nq.y), 2.0f, -1.0f), fmaf(unorm16(nq.z), 2.0f, -1.0f)});
        float integral = 0.0f;
        const float radius = (float)H * 0.5f / -cur.z;
        const int cx = (int)x, cy = (int)y;
        const int phi_k = (((cx + cy) & 0x3) << 2) + (cx & 0x3);      // phi = -(1/16) * phi_k * pi, + pi/4 per slice: tabulated
        const float rStep = radius / 2.0f;
#pragma unroll 1
        for (int samp = 0; samp < 4; samp++)
        {
            float hx = -1.0f, hy = -1.0f;
            const float cph = phi_cs[phi_k * 4 + samp].x, sph = phi_cs[phi_k * 4 + samp].y;
            const f3 sliceDir = {cph, sph, 0.0f};
            const float sdx = cph, sdy = -sph;
            float r = rStep * (0.25f * (float)((cy - cx) & 0x3));
#pragma unroll
            for (int j = 0; j < 2; j++)
            {
                const float ox = r * sdx / (float)W, oy = r * sdy / (float)H;
                r += rStep;
                const float u1 = u - ox, v1 = v - oy, u2 = u + ox, v2 = v + oy;
                const f3 ds = cs_pos(G, u1, v1) - cur, dt = cs_pos(G, u2, v2) - cur;
                float hsx = dot3(Vv, normalize3(ds)), hsy = dot3(Vv, normalize3(dt));
                if (dm_clamp(u1, 0.0f, 1.0f) != u1 || dm_clamp(v1, 0.0f, 1.0f) != v1) hsx = -1.0f;
                if (dm_clamp(u2, 0.0f, 1.0f) != u2 || dm_clamp(v2, 0.0f, 1.0f) != v2) hsy = -1.0f;
                const float fx = dm_step(hsx, hx), fy = dm_step(hsy, hy);
                hx = dm_mix(dm_mix(hx, hsx, 0.5f), dm_max(hx, hsx), fx);
                hy = dm_mix(dm_mix(hy, hsy, 0.5f), dm_max(hy, hsy), fy);
            }
            hx = fastAcos2(hx); hy = fastAcos2(hy);
            const f3 sliceNormal = normalize3(cross3(Vv, sliceDir));
            const f3 sliceBitangent = normalize3(cross3(sliceNormal, Vv));
            f3 projNorm = cn - sliceNormal * dot3(cn, sliceNormal);
            const float weight = length3(projNorm) + 1e-6f;
            projNorm = projNorm / weight;
            const float cosn = dot3(projNorm, Vv), sinn = dot3(projNorm, sliceBitangent);
            const float n = fastAcos1(cosn) * sgn(sinn);
            hx = n + dm_max(-hx - n, -half_PI_);
            hy = n + dm_min(hy - n, half_PI_);
            const float ax = -dm_cos(2.0f * hx - n) + dm_cos(n) + 2.0f * hx * dm_sin(n);
            const float ay = -dm_cos(2.0f * hy - n) + dm_cos(n) + 2.0f * hy * dm_sin(n);
            const float a = 0.25f * (ax * 1.0f + ay * 1.0f);
            integral += a * weight;
        }
        vis = integral / 4.0f;
    }
    reinterpret_cast<ushort4*>(G.raw)[(size_t)y * W + x] = make_ushort4(dm_f32_to_f16(vis), 0, 0, dm_f32_to_f16(1.0f));
}

__device__ __forceinline__ int wrapn(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

__global__ void __launch_bounds__(128) k_gtao_blur(const GtaoParams G)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= G.W || y >= G.H) return;
    const int W = (int)G.W, H = (int)G.H;
    auto R = [&](int xi, int yi) { return dm_f16_to_f32(__ldg(G.raw + 4 * ((size_t)wrapn(yi, H) * W + wrapn(xi, W)))); };
    const int offs[4][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
    float sum4[4];
#pragma unroll
    for (int g = 0; g < 4; g++)
    {
        const int bx = (int)x + offs[g][0], by = (int)y + offs[g][1];
        const float gx = R(bx, by + 1), gy = R(bx + 1, by + 1), gz = R(bx + 1, by), gw = R(bx, by);
        sum4[g] = ((gx * 1.0f + gy * 1.0f) + gz * 1.0f) + gw * 1.0f;
    }
    const float avg = (((sum4[0] * 1.0f + sum4[1] * 1.0f) + sum4[2] * 1.0f) + sum4[3] * 1.0f) / 16.0f;
    const uint16_t hv = dm_f32_to_f16(avg);
    reinterpret_cast<ushort4*>(G.out)[(size_t)y * W + x] = make_ushort4(hv, hv, hv, dm_f32_to_f16(1.0f));
}

}  // namespace

int f184_gtao_impl(f184_ctx* c, const f184_view_constants* view)
{
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_AO_RAW, F184_SLOT_AO_OUT})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    GtaoParams G{};
    memcpy(G.InvProj.m, view->InvProj, 64);
    G.depth = img_ptr<float>(c, F184_SLOT_DEPTH);
    G.normals = img_ptr<uint16_t>(c, F184_SLOT_NORMALS);
    G.raw = img_ptr<uint16_t>(c, F184_SLOT_AO_RAW);
    G.out = img_ptr<uint16_t>(c, F184_SLOT_AO_OUT);
    G.W = c->cfg.width; G.H = c->cfg.height;
    if (!c->gtao_phi_table)
    {
        CK(c, cudaMalloc(&c->gtao_phi_table, 64 * sizeof(float2)));
        k_gtao_phi_table<<<1, 16, 0, c->stream>>>(reinterpret_cast<float2*>(c->gtao_phi_table));
        CK_LAUNCH(c);
    }
    G.phi_table = reinterpret_cast<const float2*>(c->gtao_phi_table);
    int rc = f184_stage_begin(c, F184_STAGE_GTAO);
    if (rc) return rc;
    dim3 grid((G.W + 15) / 16, (G.H + 7) / 8);
    k_gtao<<<grid, 128, 0, c->stream>>>(G);
    CK_LAUNCH(c);
    k_gtao_blur<<<grid, 128, 0, c->stream>>>(G);
    CK_LAUNCH(c);
    return f184_stage_end(c, F184_STAGE_GTAO);
}
