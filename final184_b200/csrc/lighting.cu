// lighting.cu — deferred direct lighting, the consumer next to the voxel-GI section.
//
// Replaces the lighting_deferred pass (Foreground/Renderer/MegaPipeline.cpp:286-300; Shader/Lighting/aggregateLights.frag):
// per pixel, every point light and every directional light through a Cook-Torrance BRDF (GGX NDF :143-154, Schlick-GGX
// geometry :156-173, Schlick Fresnel :134-141), the directional lights shadowed by 12 Poisson taps of a bicubic-weighted
// 2x2 comparison (shadowTexSmooth :83-117 with the B-spline weights of Shader/math.inc:35-66).  Output = lightingBuffer,
// RGBA16F, alpha 1.
//
// Parity: the arithmetic follows the shader text operation by operation (no contraction: the library is built with
// -fmad=false) and is held bit for bit to that text compiled by g++ (tests/golden/refshader_*.npz, tests/test_refshader_pin.py).
// B200 notes: 48 shadow texels per pixel per directional light, gathered around one projected point — neighbouring pixels
// of a warp's 8x4 tile project to neighbouring shadow texels, so the taps are served by L1/L2 (the 2048^2 map is 16 MB);
// the kernel is ALU-bound by the 12 x (6 IEEE divisions + 8 cubic weights), not by memory.
#include "f184_device.cuh"

namespace {

constexpr float PI_ = 3.1415926f;

struct LightDev { float lum[3]; float pos[3]; };
struct LightParams
{
    M4 InvProj, ViewMat, InvModelView, ShadowView, ShadowProj;
    const float* depth; const uint16_t* normals; const uchar4* material; const float* shadow; uint16_t* out;
    const LightDev* point; const LightDev* directional;
    int n_point, n_directional;
    uint32_t W, H, S;
};

__constant__ float kPoisson12[12][2] = {
    {-0.326212f, -0.40581f}, {-0.840144f, -0.07358f}, {-0.695914f, 0.457137f}, {-0.203345f, 0.620716f},
    {0.96234f, -0.194983f},  {0.473434f, -0.480026f}, {0.519456f, 0.767022f},  {0.185461f, -0.893124f},
    {0.507431f, 0.064425f},  {0.89642f, 0.412458f},   {-0.32194f, -0.932615f}, {-0.791559f, -0.59771f}};

__device__ __forceinline__ float unorm16(uint16_t v) { return (float)v / 65535.0f; }
__device__ __forceinline__ float fastSqrt(float x) { return __int_as_float(0x1FBD1DF5 + (__float_as_int(x) >> 1)); }

// cubic B-spline weights, Shader/math.inc:35-66
__device__ __forceinline__ float w0(float a) { return (1.0f / 6.0f) * (a * (a * (-a + 3.0f) - 3.0f) + 1.0f); }
__device__ __forceinline__ float w1(float a) { return (1.0f / 6.0f) * (a * a * (3.0f * a - 6.0f) + 4.0f); }
__device__ __forceinline__ float w2(float a) { return (1.0f / 6.0f) * (a * (a * (-3.0f * a + 3.0f) + 3.0f) + 1.0f); }
__device__ __forceinline__ float w3(float a) { return (1.0f / 6.0f) * (a * a * a); }
__device__ __forceinline__ float g0(float a) { return w0(a) + w1(a); }
__device__ __forceinline__ float g1(float a) { return w2(a) + w3(a); }
__device__ __forceinline__ float h0(float a) { return -1.0f + w1(a) / (w0(a) + w1(a)); }
__device__ __forceinline__ float h1(float a) { return 1.0f + w3(a) / (w2(a) + w3(a)); }

__device__ __forceinline__ float shadow_fetch(const LightParams& P, float fx, float fy)
{
    const int x = dm_f2i(fx + 0.5f), y = dm_f2i(fy + 0.5f);
    return (x >= 0 && y >= 0 && x < (int)P.S && y < (int)P.S) ? __ldg(P.shadow + (size_t)y * P.S + x) : 0.0f;
}

// shadowTexSmooth, aggregateLights.frag:83-117
__device__ float shadow_smooth(const LightParams& P, float sx, float sy, float sz, float bias)
{
    const float res = (float)P.S;
    const float ux = sx * res - 1.0f, uy = sy * res - 1.0f;
    const float ix = floorf(ux), iy = floorf(uy);
    const float fx = ux - ix, fy = uy - iy;
    const float g0x = g0(fx), g1x = g1(fx);
    const float h0x = h0(fx) * 0.75f, h1x = h1(fx) * 0.75f, h0y = h0(fy) * 0.75f, h1y = h1(fy) * 0.75f;
    const float r0 = dm_step(sz, shadow_fetch(P, ix + h0x, iy + h0y) + bias);
    const float r1 = dm_step(sz, shadow_fetch(P, ix + h1x, iy + h0y) + bias);
    const float r2 = dm_step(sz, shadow_fetch(P, ix + h0x, iy + h1y) + bias);
    const float r3 = dm_step(sz, shadow_fetch(P, ix + h1x, iy + h1y) + bias);
    return g0(fy) * (g0x * r0 + g1x * r1) + g1(fy) * (g0x * r2 + g1x * r3);
}

__device__ __forceinline__ float ggx_schlick(float NdotV, float roughness)      // :156-165
{
    const float r = roughness + 1.0f;
    const float k = r * r / 8.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
}

// illumination, aggregateLights.frag:176-205
__device__ f3 illumination(f3 lightVector, const float* lum, f3 cspos, f3 csnorm, float metallicity, float roughness)
{
    const f3 wi = normalize3(lightVector);
    const f3 wo = normalize3(neg3(cspos));
    const f3 halfvec = normalize3(wi + wo);
    const float distSq = dot3(lightVector, lightVector);
    const float dist = fastSqrt(distSq);
    const f3 radiance = {lum[0] / distSq, lum[1] / distSq, lum[2] / distSq};
    // NDF :143-154
    float r4 = roughness; r4 *= r4; r4 *= r4;
    float cTheta = dm_max(dot3(csnorm, halfvec), 0.0f);
    cTheta *= cTheta;
    float nd = cTheta * (r4 - 1.0f) + 1.0f;
    nd *= nd;
    const float normalDist = r4 / (PI_ * nd);
    // G :167-174
    const float NdotV = dm_max(dot3(csnorm, wo), 0.0f), NdotL0 = dm_max(dot3(csnorm, wi), 0.0f);
    const float g = ggx_schlick(NdotL0, roughness) * ggx_schlick(NdotV, roughness);
    // metallicFresnel :134-141 (the albedo term it builds is not used by the shipped shader)
    const float cosTheta = dm_max(dot3(wo, halfvec), 0.0f);
    const float fresnel = 0.04f + 0.96f * dm_pow(1.0f - cosTheta, 5.0f);
    const float num = normalDist * g * fresnel;
    const float denom = 4.0f * dm_max(dot3(csnorm, wo), 0.0f) * dm_max(dot3(csnorm, wi), 0.0f);
    const float specular = num / dm_max(denom, 0.001f);
    float diffuse = 1.0f - fresnel;
    diffuse *= 1.0f - metallicity;
    const f3 wid = {wi.x / dist, wi.y / dist, wi.z / dist};
    const float NdotL = dm_max(dot3(csnorm, wid), 0.0f);
    const float ds = diffuse + specular;
    return {ds * radiance.x * NdotL, ds * radiance.y * NdotL, ds * radiance.z * NdotL};
}

__global__ void __launch_bounds__(128) k_lighting_deferred(const LightParams P)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= P.W || y >= P.H) return;
    const uint32_t W = P.W, H = P.H;
    const float u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;
    const int dx = dm_f2i(u * (float)W), dy = dm_f2i(v * (float)H);
    const float depth = (dx >= 0 && dy >= 0 && dx < (int)W && dy < (int)H) ? __ldg(P.depth + (size_t)dy * W + dx) : 0.0f;
    const f4 cp = mul44(P.InvProj, f4{u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f});
    const f3 cspos = {cp.x / cp.w, cp.y / cp.w, cp.z / cp.w};
    const ushort4 nq = __ldg(reinterpret_cast<const ushort4*>(P.normals) + (size_t)y * W + x);
    const f3 csnorm = normalize3(f3{fmaf(unorm16(nq.x), 2.0f, -1.0f), fmaf(unorm16(nq.y), 2.0f, -1.0f), fmaf(unorm16(nq.z), 2.0f, -1.0f)});
    const uchar4 mq = __ldg(P.material + (size_t)y * W + x);
    const float roughness = (float)mq.y / 255.0f, metallicity = (float)mq.z / 255.0f;     // getMaterial = .yz (:60-70)
    f3 result = {0.0f, 0.0f, 0.0f};
    for (int i = 0; i < P.n_point; i++)
    {
        const LightDev L = P.point[i];
        const f3 lp = mul43(P.ViewMat, f3{L.pos[0], L.pos[1], L.pos[2]}, 1.0f);
        result = result + illumination(lp - cspos, L.lum, cspos, csnorm, metallicity, roughness);
    }
    for (int i = 0; i < P.n_directional; i++)
    {
        const LightDev L = P.directional[i];
        const f3 lightVector = neg3(normalize3(mul33(P.ViewMat, f3{L.pos[0], L.pos[1], L.pos[2]})));
        const f3 wpos = mul43(P.InvModelView, f3{cspos.x + csnorm.x * 0.01f, cspos.y + csnorm.y * 0.01f, cspos.z + csnorm.z * 0.01f}, 1.0f);
        f4 spos = mul44(P.ShadowProj, mul44(P.ShadowView, f4{wpos.x, wpos.y, wpos.z, 1.0f}));       // no divide by w (:227-229)
        spos.x = spos.x * 0.5f + 0.5f; spos.y = spos.y * 0.5f + 0.5f;
        const float pix = 1.0f / (float)P.S;
        float shade = 0.0f;
#pragma unroll 1
        for (int j = 0; j < 12; j++)
            shade += shadow_smooth(P, spos.x + kPoisson12[j][0] * pix, spos.y + kPoisson12[j][1] * pix, spos.z + 0.0f, 0.002f);
        shade /= 12.0f;
        const f3 il = illumination(lightVector, L.lum, cspos, csnorm, metallicity, roughness);
        result = {result.x + il.x * shade, result.y + il.y * shade, result.z + il.z * shade};
    }
    reinterpret_cast<ushort4*>(P.out)[(size_t)y * W + x] =
        make_ushort4(dm_f32_to_f16(result.x), dm_f32_to_f16(result.y), dm_f32_to_f16(result.z), dm_f32_to_f16(1.0f));
}

}  // namespace

int f184_lighting_impl(f184_ctx* c, const f184_view_constants* view, const f184_extended_matrices* m, const f184_light_list* point,
                       const f184_light_list* directional)
{
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_MATERIAL, F184_SLOT_SHADOW, F184_SLOT_LIGHTING})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    const int np = point ? point->numLights : 0, nd = directional ? directional->numLights : 0;
    if (np < 0 || np > 100 || nd < 0 || nd > 100) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "lighting: light counts must be 0..100 (MAX_LIGHT_COUNT)");
    if (!c->lights_dev) CK(c, cudaMalloc(&c->lights_dev, 200 * sizeof(LightDev)));
    LightDev host[200];
    for (int i = 0; i < np; i++) { memcpy(host[i].lum, point->lights[i].luminance, 12); memcpy(host[i].pos, point->lights[i].position, 12); }
    for (int i = 0; i < nd; i++) { memcpy(host[100 + i].lum, directional->lights[i].luminance, 12); memcpy(host[100 + i].pos, directional->lights[i].position, 12); }
    // pageable source: the copy is staged before the call returns, `host` may go out of scope
    if (np) CK(c, cudaMemcpyAsync(c->lights_dev, host, np * sizeof(LightDev), cudaMemcpyHostToDevice, c->stream));
    if (nd) CK(c, cudaMemcpyAsync(reinterpret_cast<LightDev*>(c->lights_dev) + 100, host + 100, nd * sizeof(LightDev), cudaMemcpyHostToDevice, c->stream));
    LightParams P{};
    memcpy(P.InvProj.m, view->InvProj, 64); memcpy(P.ViewMat.m, view->ViewMat, 64);
    memcpy(P.InvModelView.m, m->InvModelView, 64); memcpy(P.ShadowView.m, m->ShadowView, 64); memcpy(P.ShadowProj.m, m->ShadowProj, 64);
    P.depth = img_ptr<float>(c, F184_SLOT_DEPTH); P.normals = img_ptr<uint16_t>(c, F184_SLOT_NORMALS);
    P.material = img_ptr<uchar4>(c, F184_SLOT_MATERIAL); P.shadow = img_ptr<float>(c, F184_SLOT_SHADOW);
    P.out = img_ptr<uint16_t>(c, F184_SLOT_LIGHTING);
    P.point = reinterpret_cast<const LightDev*>(c->lights_dev); P.directional = P.point + 100;
    P.n_point = np; P.n_directional = nd;
    P.W = c->cfg.width; P.H = c->cfg.height; P.S = c->cfg.shadow_res;
    int rc = f184_stage_begin(c, F184_STAGE_LIGHTING);
    if (rc) return rc;
    k_lighting_deferred<<<dim3((P.W + 15) / 16, (P.H + 7) / 8), 128, 0, c->stream>>>(P);
    CK_LAUNCH(c);
    return f184_stage_end(c, F184_STAGE_LIGHTING);
}
