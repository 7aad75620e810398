// mode_n_trace.cu — north-star indirect pass: per-pixel diffuse + specular voxel cone tracing over the
// G-buffer (DESIGN.md "Mode N" B.5).
//
// Takes the place of the reference's lighting_indirect pass (Foreground/Renderer/MegaPipeline.cpp:252-268,
// Shader/Lighting/indirect.frag), which marches 8 stochastic rays of 60 fixed steps through a single-level
// volume.  Here each pixel traces 6 diffuse cones (60 deg aperture) and 1 specular cone through the
// injected radiance volume and its six-direction mip chain, front-to-back, and ends with the reference's
// temporal reprojection blend (indirect.frag:225-240) so the output image is a drop-in for indirectImage.
//
// B200 design
//   * tile-per-warp: a warp owns an 8x4 pixel tile, so its 32 cones of the same index leave neighbouring
//     surface points in nearly the same direction and their texel footprints overlap in L1/tex cache.
//   * all volume reads are hardware-filtered: level 0 is a 3D array (trilinear), levels >= 1 are one
//     mipmapped 3D array (six directions as z-slabs) sampled with tex3DLod at the NEAREST level (trilinear, point mip filter), so a cone
//     sample is 3 texture instructions of 8 texels each (direction-weighted x/y/z faces), or 1 while the cone is
//     thinner than ~1.4 voxels.  ncu on the first version (mip-linear, 16 texels per fetch) showed the kernel at
//     87.5 % of the TEX data-pipe wavefront peak with a 99.8 % L1 hit rate: the texture pipe, not memory, is
//     the bound, so the fix is fewer texels per sample, not better locality.
//   * early termination: a cone stops at alpha >= 0.95, on leaving the volume or at max distance; the loop
//     exit reconverges per warp, i.e. the warp leaves as soon as its last lane is done (the vote).
//   * cone-samples are counted (one warp-aggregated atomic per warp) because Gcone-samples/s is a metric.
#include <cstdlib>

#include "f184_device.cuh"
#include "f184_cone.cuh"

namespace {

struct ConeParams
{
    M4 InvProj, InvModelView, w2v, prevModelView, prevProjection;
    cudaTextureObject_t level0;
    cudaTextureObject_t atlas;             // levels >= 1, six directions as z-slabs of one mipmapped 3D array (nearest mip level)
    cudaTextureObject_t atlas_lin;         // the same array, linear between mip levels (Appendix-B spec)
    // one NVLink box: [0] = level 0 of the sampled set holds every rank's bricks; [1] = sticky error word.  nullptr on one GPU.
    const uint32_t* l0_full;
    uint32_t* dev_error;
    const float* depth;
    const uint16_t* normals;
    const uchar4* material;
    const uint16_t* hist;
    uint16_t* out;
    float h, max_dist, exposure, max_lod;
    f3 cam;
    uint32_t W, H, y0, y1, tile0, tile_stride;
    uint32_t vy0, vh;                      // view window: rows [vy0, vy0 + vh) are one view (the whole image unless f184_trace_views)
};

__constant__ float kDiffuseDirs[6][3] = {
    {0.0f, 0.0f, 1.0f},
    {0.8660254f, 0.0f, 0.5f},
    {0.26761657f, 0.82363910f, 0.5f},
    {-0.70062927f, 0.50903696f, 0.5f},
    {-0.70062927f, -0.50903696f, 0.5f},
    {0.26761657f, -0.82363910f, 0.5f}};
__constant__ float kDiffuseW[6] = {0.25f, 0.15f, 0.15f, 0.15f, 0.15f, 0.15f};

// Direction-weighted fetch from the six-direction atlas: direction d of every level occupies the normalised z range
// [d/6, d/6 + 1/12) (f184_internal.h), so a face is a per-cone z offset and all three fetches use ONE warp-uniform texture
// handle: three independent TEX instructions back to back.  (With one texture object per direction the handle differed
// per lane and ptxas wrapped every TEX in an R2UR/BRA.U.ANY waterfall loop whose result fed the next one.)
__device__ __forceinline__ float4 tex_dir(cudaTextureObject_t atlas, const float w[3], const float zoff[3], float qx, float qy, float qz, float lod)
{
    const float z12 = qz * (1.0f / 12.0f);
    float4 s[3];
#pragma unroll
    for (int a = 0; a < 3; a++)
        s[a] = (w[a] != 0.0f) ? tex3DLod<float4>(atlas, qx, qy, z12 + zoff[a], lod) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 r;
    r.x = (w[0] * s[0].x + w[1] * s[1].x) + w[2] * s[2].x;
    r.y = (w[0] * s[0].y + w[1] * s[1].y) + w[2] * s[2].y;
    r.z = (w[0] * s[0].z + w[1] * s[1].z) + w[2] * s[2].z;
    r.w = (w[0] * s[0].w + w[1] * s[1].w) + w[2] * s[2].w;
    return r;
}

// SPEC_B = false: DESIGN.md B.5 as amended (nearest mip level, one sample per voxel of the sampled level: t += diam).
// SPEC_B = true:  SURVEY.md Appendix B.5 as written (F184_FLAG_SPEC_APPENDIX_B): mip-linear sampling — below lod 1 a blend of the
//                 isotropic level 0 and the directional level 1, above it the hardware's linear mip filter — and t += diam / 2.
template <bool SPEC_B>
__device__ f3 trace_cone(const ConeParams& P, f3 origin, f3 dir, float tan_half, unsigned int& samples)
{
    const float h = P.h;
    // direction through the voxel lattice (the voxel camera is a signed axis permutation + scale)
    const f3 dv = {(P.w2v.m[0] * dir.x + P.w2v.m[4] * dir.y + P.w2v.m[8] * dir.z) * 0.5f,
                   (P.w2v.m[1] * dir.x + P.w2v.m[5] * dir.y + P.w2v.m[9] * dir.z) * 0.5f,
                   (P.w2v.m[2] * dir.x + P.w2v.m[6] * dir.y + P.w2v.m[10] * dir.z)};
    const float dl = length3(dv);
    const f3 du = {dv.x / dl, dv.y / dl, dv.z / dl};
    const float w[3] = {du.x * du.x, du.y * du.y, du.z * du.z};
    const float zoff[3] = {du.x < 0.0f ? 1.0f / 6.0f : 0.0f, du.y < 0.0f ? 3.0f / 6.0f : 2.0f / 6.0f, du.z < 0.0f ? 5.0f / 6.0f : 4.0f / 6.0f};
    // normalised volume coordinate of the origin; q(t) = q0 + dv * t (affine)
    const f3 o3 = mul43(P.w2v, origin, 1.0f);
    const f3 q0 = {o3.x * 0.5f + 0.5f, o3.y * 0.5f + 0.5f, o3.z};
    float t = 2.0f * h, A = 0.0f;
    f3 acc = {0.f, 0.f, 0.f};
    const float inv_h = 1.0f / h;
    while (A < 0.95f && t < P.max_dist)
    {
        float diam;
        const float lod = cone_lod(t, tan_half, h, inv_h, &diam);
        const float qx = q0.x + dv.x * t, qy = q0.y + dv.y * t, qz = q0.z + dv.z * t;
        if (!(qx >= 0.0f && qx <= 1.0f && qy >= 0.0f && qy <= 1.0f && qz >= 0.0f && qz <= 1.0f)) break;
        samples++;
        float4 s;
        if (SPEC_B)
        {
            if (lod < 1.0f)
            {   // between the isotropic level 0 and the directional level 1
                if (P.l0_full && !*P.l0_full) atomicOr(P.dev_error, F184_DEVERR_LEVEL0_MISSING);
                const float4 a = tex3D<float4>(P.level0, qx, qy, qz);
                const float4 b = tex_dir(P.atlas, w, zoff, qx, qy, qz, 0.0f);
                s = make_float4(a.x + lod * (b.x - a.x), a.y + lod * (b.y - a.y), a.z + lod * (b.z - a.z), a.w + lod * (b.w - a.w));
            }
            else s = tex_dir(P.atlas_lin, w, zoff, qx, qy, qz, fminf(lod - 1.0f, P.max_lod));
        }
        else
        {
            // nearest level (point mip filter): 0 = the isotropic radiance volume, L >= 1 = the six-direction chain
            const int L = (int)floorf(lod + 0.5f);
            if (L <= 0)
            {   // one NVLink box: the gather fetched level 0 only if k_need_level0 saw a cone like this one; anything else is an API misuse
                if (P.l0_full && !*P.l0_full) atomicOr(P.dev_error, F184_DEVERR_LEVEL0_MISSING);
                s = tex3D<float4>(P.level0, qx, qy, qz);
            }
            else s = tex_dir(P.atlas, w, zoff, qx, qy, qz, fminf((float)(L - 1), P.max_lod));
        }
        const float k = 1.0f - A;
        acc = {acc.x + k * s.x, acc.y + k * s.y, acc.z + k * s.z};
        A += k * s.w;
        t += SPEC_B ? 0.5f * diam : diam;      // amended spec: one sample per voxel of the sampled level along the axis
    }
    const float rem = fmaxf(0.0f, 1.0f - A);
    return {acc.x * P.exposure + 0.7f * 0.4f * rem, acc.y * P.exposure + 0.8f * 0.4f * rem, acc.z * P.exposure + 1.0f * 0.4f * rem};
}

__device__ __forceinline__ float unorm16(uint16_t v) { return (float)v / 65535.0f; }
__device__ __forceinline__ int wrapn(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

// CTA = 256 threads = a 32 x 8 pixel strip of eight 8 x 4 warp tiles, 64 registers per thread: FOUR CTAs fill an SM's register
// file, so every CTA that retires leaves a hole of 16 K registers — and every kernel of the build stream (frame pipeline) is
// shaped to fit that hole (<= 16 K registers per CTA).  With 128-thread CTAs the holes were 8 K registers, no build kernel fitted
// one, the block scheduler back-filled them with the next trace CTA, and the "concurrent" build kernels waited for the trace grid
// to drain (the gather took 4.0 ms beside the trace, 0.6 ms alone).
constexpr int TRACE_THREADS = 256;
template <int MIN_CTAS, bool SPEC_B>
__global__ void __launch_bounds__(TRACE_THREADS, MIN_CTAS) k_trace_n(const ConeParams P, unsigned long long* __restrict__ sample_counter)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const uint32_t y = P.y0 + (P.tile0 + blockIdx.y * P.tile_stride) * 8 + (warp >> 2) * 4 + (lane >> 3);
    unsigned int samples = 0;
    if (x < P.W && y < P.y1)
    {
        const uint32_t W = P.W, H = P.vh;
        const float uvx = ((float)x + 0.5f) / (float)W, uvy = ((float)(y - P.vy0) + 0.5f) / (float)P.vh;
        const float depth = __ldg(P.depth + (size_t)y * W + x);
        const f4 cp = mul44(P.InvProj, f4{uvx * 2.0f - 1.0f, uvy * 2.0f - 1.0f, depth, 1.0f});
        const f3 cspos = {cp.x / cp.w, cp.y / cp.w, cp.z / cp.w};
        ushort4 o;
        if (depth >= 1.0f) o = make_ushort4(0, 0, 0, dm_f32_to_f16(-cspos.z));
        else
        {
            const f3 wpos = mul43(P.InvModelView, cspos, 1.0f);
            const ushort4 nq = __ldg(reinterpret_cast<const ushort4*>(P.normals) + (size_t)y * W + x);
            const f3 csnorm = normalize3(f3{fmaf(unorm16(nq.x), 2.0f, -1.0f), fmaf(unorm16(nq.y), 2.0f, -1.0f), fmaf(unorm16(nq.z), 2.0f, -1.0f)});
            const f3 wnorm = mul33(P.InvModelView, csnorm);
            f3 z = wnorm, hh = wnorm;
            if (fabsf(hh.x) <= fabsf(hh.y) && fabsf(hh.x) <= fabsf(hh.z)) hh.x = 1.0f;
            else if (fabsf(hh.y) <= fabsf(hh.x) && fabsf(hh.y) <= fabsf(hh.z)) hh.y = 1.0f;
            else hh.z = 1.0f;
            z = normalize3(z);
            const f3 ty = normalize3(cross3(hh, z));
            const f3 tx = normalize3(cross3(z, ty));
            const f3 origin = {wpos.x + z.x * P.h, wpos.y + z.y * P.h, wpos.z + z.z * P.h};
            f3 ind = {0.f, 0.f, 0.f};
#pragma unroll 1
            for (int i = 0; i < 6; i++)
            {
                const float d0 = kDiffuseDirs[i][0], d1 = kDiffuseDirs[i][1], d2 = kDiffuseDirs[i][2];
                const f3 dir = {(tx.x * d0 + ty.x * d1) + z.x * d2, (tx.y * d0 + ty.y * d1) + z.y * d2, (tx.z * d0 + ty.z * d1) + z.z * d2};
                const f3 r = trace_cone<SPEC_B>(P, origin, dir, kTanHalfDiffuse, samples);
                const float wgt = kDiffuseW[i];
                ind = {ind.x + wgt * r.x, ind.y + wgt * r.y, ind.z + wgt * r.z};
            }
            {
                const float rough = (float)__ldg(P.material + (size_t)y * W + x).y / 255.0f;
                const float tan_half = cone_specular_tan(rough);
                const f3 I = normalize3(wpos - P.cam);
                const float ndi = dot3(z, I);
                const f3 R = {I.x - 2.0f * ndi * z.x, I.y - 2.0f * ndi * z.y, I.z - 2.0f * ndi * z.z};
                if (dot3(R, z) > 0.0f)
                {
                    const f3 r = trace_cone<SPEC_B>(P, origin, R, tan_half, samples);
                    const float om = 1.0f - fmaxf(-ndi, 0.0f);
                    const float F = 0.04f + 0.96f * (om * om * om * om * om);
                    ind = {ind.x + F * r.x, ind.y + F * r.y, ind.z + F * r.z};
                }
            }
            // temporal reprojection, indirect.frag:225-240
            const f4 pc = mul44(P.prevModelView, f4{wpos.x, wpos.y, wpos.z, 1.0f});
            const f4 pp = mul44(P.prevProjection, pc);
            float ru = pp.x / pp.w, rv = pp.y / pp.w;
            ru = ru * 0.5f + 0.5f; rv = rv * 0.5f + 0.5f;
            if (dm_clamp(ru, 0.0f, 1.0f) == ru && dm_clamp(rv, 0.0f, 1.0f) == rv)
            {
                const float fx = ru * (float)W - 0.5f, fy = rv * (float)H - 0.5f;
                const float x0f = floorf(fx), y0f = floorf(fy);
                const float wx = fx - x0f, wy = fy - y0f;
                const int xi0 = dm_f2i(x0f), yi0 = dm_f2i(y0f);
                const int xa = wrapn(xi0, (int)W), xb = wrapn(xi0 + 1, (int)W), ya = (int)P.vy0 + wrapn(yi0, (int)H), yb = (int)P.vy0 + wrapn(yi0 + 1, (int)H);
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
                ind = {dm_clamp(dm_mix(ind.x, p0, bw), 0.0f, 16.0f), dm_clamp(dm_mix(ind.y, p1, bw), 0.0f, 16.0f), dm_clamp(dm_mix(ind.z, p2, bw), 0.0f, 16.0f)};
            }
            o = make_ushort4(dm_f32_to_f16(ind.x), dm_f32_to_f16(ind.y), dm_f32_to_f16(ind.z), dm_f32_to_f16(-cspos.z));
        }
        reinterpret_cast<ushort4*>(P.out)[(size_t)y * W + x] = o;
    }
    warp_count_add(sample_counter, samples);
}

}  // namespace

// parameter block of one view; window = rows [vy0, vy0 + vh)
static int trace_params(f184_ctx* c, const VolumeSet& vs, const f184_trace_constants* k, uint32_t vy0, uint32_t vh, ConeParams& P)
{
    memcpy(P.InvProj.m, k->view.InvProj, 64);
    memcpy(P.InvModelView.m, k->ext.InvModelView, 64);
    M4 vp, vv;
    memcpy(vp.m, k->ext.VoxelProj, 64);
    memcpy(vv.m, k->ext.VoxelView, 64);
    P.w2v = host_matmul(vp, vv);
    memcpy(P.prevModelView.m, k->prev.PrevModelView, 64);
    memcpy(P.prevProjection.m, k->prev.PrevProjection, 64);
    P.h = f184_voxel_h(k->ext.VoxelProj, k->ext.VoxelView, c->cfg.grid_n);
    P.level0 = vs.rad_tex;
    P.atlas = vs.dir_tex;
    P.atlas_lin = vs.dir_tex_lin;
    if (c->cfg.nranks > 1)
    {
        P.l0_full = c->dev_state + F184_DEV_L0_FULL + (&vs - c->vs);
        P.dev_error = c->dev_state + F184_DEV_ERROR;
    }
    P.depth = img_ptr<float>(c, F184_SLOT_DEPTH);
    P.normals = img_ptr<uint16_t>(c, F184_SLOT_NORMALS);
    P.material = img_ptr<uchar4>(c, F184_SLOT_MATERIAL);
    P.hist = img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_HISTORY);
    P.out = img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_OUT);
    P.max_dist = c->cfg.cone_max_distance;
    P.exposure = f184_exposure(c, &k->sun);
    P.max_lod = (float)(c->n_mip_levels - 1);
    P.cam = {P.InvModelView.m[12], P.InvModelView.m[13], P.InvModelView.m[14]};
    P.W = c->cfg.width; P.H = c->cfg.height;
    P.vy0 = vy0; P.vh = vh;
    return F184_OK;
}

int f184_trace_init_n(f184_ctx* c)
{
    // The tracer uses no shared memory, so by default its launch configures the SMs with the smallest shared-memory carve-out (all of
    // the unified array as L1 / texture cache) — and a kernel that NEEDS shared memory (every kernel of the build stream: brick rings,
    // the gather's 64 KB ring) cannot become resident on an SM until the tracer's CTAs have drained from it and the SM is reconfigured:
    // the frame pipeline's streams would take turns instead of running side by side.  Ask for a carve-out that leaves room for them
    // (F184_TRACE_CARVEOUT, per cent of the maximum; the tracer's texel working set is a few KB per warp and keeps its hit rate).
    static const int carveout = [] { const char* e = getenv("F184_TRACE_CARVEOUT"); return e ? atoi(e) : 44; }();
    static bool carveout_set = false;
    if (!carveout_set && carveout >= 0)
    {
        CK(c, cudaFuncSetAttribute(k_trace_n<3, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
        CK(c, cudaFuncSetAttribute(k_trace_n<3, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
        CK(c, cudaFuncSetAttribute(k_trace_n<4, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
        carveout_set = true;
    }
    return F184_OK;
}

static int trace_launch(f184_ctx* c, const ConeParams& P, uint32_t grid_y, cudaStream_t stream)
{
    dim3 grid((P.W + 31) / 32, grid_y);
    // two register budgets of the same kernel: 4 CTAs/SM (64 registers) or 3 (80, no spill); F184_TRACE_CTAS=3 selects the latter (A/B knob)
    static const int min_ctas = [] { const char* e = getenv("F184_TRACE_CTAS"); return e ? atoi(e) : 4; }();
    if (c->cfg.flags & F184_FLAG_SPEC_APPENDIX_B) k_trace_n<3, true><<<grid, TRACE_THREADS, 0, stream>>>(P, c->counters_dev + F184_COUNTER_MARCH_STEPS);
    else if (min_ctas == 3) k_trace_n<3, false><<<grid, TRACE_THREADS, 0, stream>>>(P, c->counters_dev + F184_COUNTER_MARCH_STEPS);
    else k_trace_n<4, false><<<grid, TRACE_THREADS, 0, stream>>>(P, c->counters_dev + F184_COUNTER_MARCH_STEPS);
    CK_LAUNCH(c);
    return F184_OK;
}

static int trace_check(f184_ctx* c)
{
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_MATERIAL, F184_SLOT_INDIRECT_OUT, F184_SLOT_INDIRECT_HISTORY})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    if (!c->n_sets) return f184_fail(c, F184_ERR_NOT_READY, "trace: call f184_inject and f184_build_mips first");
    return F184_OK;
}

int f184_trace_n(f184_ctx* c, const f184_trace_constants* k)
{
    int rc = trace_check(c);
    if (rc) return rc;
    VolumeSet* vs = nullptr;
    rc = f184_volume_acquire(c, &vs);      // the newest complete set; the pass stream waits for its build (inject, mips, gather) only
    if (rc) return rc;
    ConeParams P{};
    trace_params(c, *vs, k, 0, c->cfg.height, P);
    const uint32_t grid_y = f184_trace_tiles(c, P.H, &P.y0, &P.y1, &P.tile0, &P.tile_stride);
    if (P.tile_stride > 1 && (P.y0 & 7)) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "trace: row range must start on a multiple of 8 when tiles are interleaved");
    rc = f184_stage_begin(c, F184_STAGE_TRACE);
    if (rc) return rc;
    if (k->reset_history && (rc = f184_fill_async(c, c->img[F184_SLOT_INDIRECT_HISTORY].ptr, 0u, c->img[F184_SLOT_INDIRECT_HISTORY].desc.size_bytes, c->stream))) return rc;
    if ((rc = f184_zero_counters(c, 1u << F184_COUNTER_MARCH_STEPS))) return rc;
    if (grid_y && (rc = trace_launch(c, P, grid_y, c->stream))) return rc;
    rc = f184_stage_end(c, F184_STAGE_TRACE);
    if (rc) return rc;
    return f184_volume_release(c, vs);
}

// Probe batch (BASELINE configs[4]; f184_trace_views in f184.h): the images hold H / view_h views stacked top to bottom;
// view v = rows [v * view_h, (v + 1) * view_h), traced with ks[v].  Views are independent (disjoint rows of every image,
// one shared read-only volume), so they go round-robin onto a few internal streams: the ragged tail of one view's grid
// (a 512 x 512 view is 1.7 waves of CTAs) is filled by the head of the next instead of idling the SMs.
int f184_trace_views_n(f184_ctx* c, const f184_trace_constants* ks, uint32_t view_h, uint32_t first, uint32_t count)
{
    int rc = trace_check(c);
    if (rc) return rc;
    const uint32_t H = c->cfg.height;
    if (view_h == 0 || H % view_h != 0 || (uint64_t)(first + (uint64_t)count) * view_h > H)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "trace_views: %u views of %u rows from view %u do not fit %u rows", count, view_h, first, H);
    if (!c->view_streams[0])
        for (int i = 0; i < F184_VIEW_STREAMS; i++)
        {
            CK(c, cudaStreamCreateWithFlags(&c->view_streams[i], cudaStreamNonBlocking));
            CK(c, cudaEventCreateWithFlags(&c->ev_view_done[i], cudaEventDisableTiming));
        }
    if (!c->ev_view_fork) CK(c, cudaEventCreateWithFlags(&c->ev_view_fork, cudaEventDisableTiming));
    VolumeSet* vs = nullptr;
    rc = f184_volume_acquire(c, &vs);
    if (rc) return rc;
    rc = f184_stage_begin(c, F184_STAGE_TRACE);
    if (rc) return rc;
    if ((rc = f184_zero_counters(c, 1u << F184_COUNTER_MARCH_STEPS))) return rc;
    CK(c, cudaEventRecord(c->ev_view_fork, c->stream));
    const int ns = count < (uint32_t)F184_VIEW_STREAMS ? (int)count : F184_VIEW_STREAMS;
    for (int i = 0; i < ns; i++) CK(c, cudaStreamWaitEvent(c->view_streams[i], c->ev_view_fork, 0));
    const size_t row_bytes = (size_t)c->cfg.width * 8;
    for (uint32_t v = first; v < first + count; v++)
    {
        cudaStream_t st = c->view_streams[(v - first) % F184_VIEW_STREAMS];
        ConeParams P{};
        trace_params(c, *vs, &ks[v], v * view_h, view_h, P);
        P.y0 = v * view_h; P.y1 = P.y0 + view_h; P.tile0 = 0; P.tile_stride = 1;
        if (ks[v].reset_history && (rc = f184_fill_async(c, img_ptr<uint8_t>(c, F184_SLOT_INDIRECT_HISTORY) + (size_t)P.y0 * row_bytes, 0u, (size_t)view_h * row_bytes, st))) return rc;
        if ((rc = trace_launch(c, P, (view_h + 7) / 8, st))) return rc;
    }
    for (int i = 0; i < ns; i++)
    {
        CK(c, cudaEventRecord(c->ev_view_done[i], c->view_streams[i]));
        CK(c, cudaStreamWaitEvent(c->stream, c->ev_view_done[i], 0));
    }
    rc = f184_stage_end(c, F184_STAGE_TRACE);
    if (rc) return rc;
    return f184_volume_release(c, vs);
}
