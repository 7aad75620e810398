// mode_n_inject.cu — north-star light injection (DESIGN.md "Mode N" B.3) and the texture-side storage the
// cone tracer samples.
//
// The reference has no injection pass: it evaluates the first bounce lazily at every ray hit
// (Shader/Lighting/indirect.frag:157-169: albedo^2.2, shadow-map test with a 0.06 normal offset and 0.005
// bias, two-sided Lambert against the sun).  Here that arithmetic is hoisted into the volume once per
// frame: one evaluation per OCCUPIED VOXEL instead of one per ray hit.
//
// B200 design: only the bricks on this frame's brick list are visited (one warp per 8^3 brick, 16 passes of
// 32 voxels); each voxel reads 8 B (albedo + normal), does one shadow-map gather and writes the RGBA8
// radiance twice: to the linear level-0 volume the TMA mip builder reads, and through a surface to the
// 3D CUDA array the tracer's texture unit filters.  Radiance is stored / exposure (default: the largest sun
// luminance component) so 8 bits cover the injected range exactly.
#include <algorithm>
#include <cmath>

#include "f184_device.cuh"

namespace {

struct InjectParams
{
    M4 v2w, ShadowView, ShadowProj;
    f3 sunPos, sunLum;
    float inv_exposure;
    const uchar4* alb;
    const char4* nrm;
    const float* shadow;
    uchar4* rad;
    cudaSurfaceObject_t rad_surf;
    const uint32_t* brick_list;
    const unsigned long long* brick_count;
    int N, S;
};

__global__ void __launch_bounds__(256) k_inject_n(const InjectParams P)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t count = (uint32_t)*P.brick_count;
    const int N = P.N, NB = N >> 3;
    const float Nf = (float)N;
    for (uint32_t i = warp_global; i < count; i += n_warps)
    {
        const uint32_t b = P.brick_list[i] & 0x7fffffffu;      // bit 31 = touched this frame (normalise)
        const int bx = (b % NB) << 3, by = ((b / NB) % NB) << 3, bz = (b / (NB * NB)) << 3;
#pragma unroll 2
        for (int pass = 0; pass < 16; pass++)
        {
            const int local = pass * 32 + lane;
            const int x = bx + (local & 7), y = by + ((local >> 3) & 7), z = bz + (local >> 6);
            const size_t o = ((size_t)z * N + y) * N + x;
            const uchar4 a = __ldg(P.alb + o);
            uchar4 out = make_uchar4(0, 0, 0, 0);
            if (a.w != 0)
            {
                const char4 nq = __ldg(P.nrm + o);
                const float nx = ((float)x + 0.5f) / Nf * 2.0f - 1.0f, ny = ((float)y + 0.5f) / Nf * 2.0f - 1.0f, nz = ((float)z + 0.5f) / Nf;
                const f4 p4 = mul44(P.v2w, f4{nx, ny, nz, 1.0f});
                const f3 p = {p4.x / p4.w, p4.y / p4.w, p4.z / p4.w};
                f3 n = {(float)nq.x / 127.0f, (float)nq.y / 127.0f, (float)nq.z / 127.0f};
                const float nl = length3(n);
                if (nl > 0.0f) n = n / nl;
                const f3 col = {dm_pow((float)a.x / 255.0f, 2.2f), dm_pow((float)a.y / 255.0f, 2.2f), dm_pow((float)a.z / 255.0f, 2.2f)};
                const f3 sp = {p.x + n.x * 0.06f, p.y + n.y * 0.06f, p.z + n.z * 0.06f};
                const f4 s4 = mul44(P.ShadowProj, mul44(P.ShadowView, f4{sp.x, sp.y, sp.z, 1.0f}));
                float sx = s4.x / s4.w, sy = s4.y / s4.w;
                const float sz = s4.z / s4.w;
                sx = sx * 0.5f + 0.5f; sy = sy * 0.5f + 0.5f;
                const int tx = dm_f2i(sx * (float)P.S), ty = dm_f2i(sy * (float)P.S);
                float shadowZ = 0.0f;
                if (tx >= 0 && ty >= 0 && tx < P.S && ty < P.S) shadowZ = __ldg(P.shadow + (size_t)ty * P.S + tx);
                const float shade = dm_step(sz + 0.005f, shadowZ);
                const float l = fabsf(dot3(neg3(P.sunPos), n));
                const float r0 = col.x * P.sunLum.x * l * shade, r1 = col.y * P.sunLum.y * l * shade, r2 = col.z * P.sunLum.z * l * shade;
                out = make_uchar4((unsigned char)floorf(dm_clamp(r0 * P.inv_exposure, 0.0f, 1.0f) * 255.0f + 0.5f),
                                  (unsigned char)floorf(dm_clamp(r1 * P.inv_exposure, 0.0f, 1.0f) * 255.0f + 0.5f),
                                  (unsigned char)floorf(dm_clamp(r2 * P.inv_exposure, 0.0f, 1.0f) * 255.0f + 0.5f), 255);
            }
            P.rad[o] = out;
            surf3Dwrite(out, P.rad_surf, x * 4, y, z);
        }
    }
}

__global__ void k_clear_array(cudaSurfaceObject_t s, int n)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x < n) surf3Dwrite(make_uchar4(0, 0, 0, 0), s, x * 4, y, z);
}

}  // namespace

float f184_exposure(const f184_ctx* c, const f184_sun* sun)
{
    if (c->cfg.radiance_exposure > 0.0f) return c->cfg.radiance_exposure;
    const float m = std::max(sun->luminance[0], std::max(sun->luminance[1], sun->luminance[2]));
    return m > 0.0f ? m : 1.0f;
}

// double-precision Gauss-Jordan inverse of a float 4x4 in upload order
M4 f184_invert_m4(const M4& A)
{
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int col = 0; col < 4; col++) { a[r][col] = A.m[col * 4 + r]; a[r][4 + col] = (r == col) ? 1.0 : 0.0; }
    for (int col = 0; col < 4; col++)
    {
        int piv = col;
        for (int r = col + 1; r < 4; r++)
            if (std::fabs(a[r][col]) > std::fabs(a[piv][col])) piv = r;
        if (piv != col)
            for (int k = 0; k < 8; k++) std::swap(a[piv][k], a[col][k]);
        const double d = a[col][col];
        for (int k = 0; k < 8; k++) a[col][k] /= d;
        for (int r = 0; r < 4; r++)
            if (r != col)
            {
                const double f = a[r][col];
                if (f != 0.0)
                    for (int k = 0; k < 8; k++) a[r][k] -= f * a[col][k];
            }
    }
    M4 R;
    for (int r = 0; r < 4; r++)
        for (int col = 0; col < 4; col++) R.m[col * 4 + r] = (float)a[r][4 + col];
    return R;
}

// Texture-side storage: level 0 as a 3D array, levels >= 1 as six mipmapped 3D arrays (one per direction),
// all surface-writable, sampled with normalised coordinates, trilinear + mip-linear, border addressing.
int f184_mode_n_alloc(f184_ctx* c)
{
    if (c->rad_array) return F184_OK;
    const int N = (int)c->cfg.grid_n;
    cudaChannelFormatDesc ch = cudaCreateChannelDesc<uchar4>();
    CK(c, cudaMalloc3DArray(&c->rad_array, &ch, make_cudaExtent(N, N, N), cudaArraySurfaceLoadStore));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->rad_array;
    CK(c, cudaCreateSurfaceObject(&c->rad_surf, &rd));
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    CK(c, cudaCreateTextureObject(&c->rad_tex, &rd, &td, nullptr));
    {
        dim3 g((N + 127) / 128, N, N);
        k_clear_array<<<g, 128, 0, c->stream>>>(c->rad_surf, N);
        CK_LAUNCH(c);
    }
    uint32_t levels = 0;
    for (int n = N / 2; n >= 1; n /= 2) levels++;
    c->n_mip_levels = levels;
    c->mip_levels.clear();
    uint64_t off = 0;
    for (int n = N / 2; n >= 1; n /= 2)
    {
        c->mip_levels.push_back(MipLevelInfo{(uint32_t)n, off});
        off += 6ull * n * n * n;
    }
    for (int d = 0; d < 6; d++)
    {
        CK(c, cudaMallocMipmappedArray(&c->dir_arrays[d], &ch, make_cudaExtent(N / 2, N / 2, N / 2), levels, cudaArraySurfaceLoadStore));
        for (uint32_t l = 0; l < levels; l++)
        {
            cudaArray_t la;
            CK(c, cudaGetMipmappedArrayLevel(&la, c->dir_arrays[d], l));
            cudaResourceDesc lrd{};
            lrd.resType = cudaResourceTypeArray;
            lrd.res.array.array = la;
            CK(c, cudaCreateSurfaceObject(&c->dir_surf[d][l], &lrd));
        }
        cudaResourceDesc mrd{};
        mrd.resType = cudaResourceTypeMipmappedArray;
        mrd.res.mipmap.mipmap = c->dir_arrays[d];
        cudaTextureDesc mtd = td;
        mtd.mipmapFilterMode = cudaFilterModeLinear;
        mtd.minMipmapLevelClamp = 0.0f;
        mtd.maxMipmapLevelClamp = (float)(levels - 1);
        CK(c, cudaCreateTextureObject(&c->dir_tex[d], &mrd, &mtd, nullptr));
    }
    return F184_OK;
}

int f184_mode_n_release(f184_ctx* c)
{
    if (c->rad_tex) cudaDestroyTextureObject(c->rad_tex);
    if (c->rad_surf) cudaDestroySurfaceObject(c->rad_surf);
    if (c->rad_array) cudaFreeArray(c->rad_array);
    for (int d = 0; d < 6; d++)
    {
        if (c->dir_tex[d]) cudaDestroyTextureObject(c->dir_tex[d]);
        for (int l = 0; l < 12; l++)
            if (c->dir_surf[d][l]) cudaDestroySurfaceObject(c->dir_surf[d][l]);
        if (c->dir_arrays[d]) cudaFreeMipmappedArray(c->dir_arrays[d]);
    }
    c->rad_tex = 0; c->rad_surf = 0; c->rad_array = nullptr;
    return F184_OK;
}

int f184_inject_n(f184_ctx* c, const f184_sun* sun, const f184_extended_matrices* m)
{
    for (int s : {F184_SLOT_VOX_ALBEDO, F184_SLOT_VOX_NORMAL, F184_SLOT_RADIANCE, F184_SLOT_SHADOW})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    if (!c->brick_list) return f184_fail(c, F184_ERR_NOT_READY, "inject: call f184_voxelize first");
    int rc = f184_mode_n_alloc(c);
    if (rc) return rc;
    InjectParams P{};
    M4 vp, vv;
    memcpy(vp.m, m->VoxelProj, 64);
    memcpy(vv.m, m->VoxelView, 64);
    P.v2w = f184_invert_m4(host_matmul(vp, vv));
    memcpy(P.ShadowView.m, m->ShadowView, 64);
    memcpy(P.ShadowProj.m, m->ShadowProj, 64);
    P.sunPos = {sun->position[0], sun->position[1], sun->position[2]};
    P.sunLum = {sun->luminance[0], sun->luminance[1], sun->luminance[2]};
    P.inv_exposure = 1.0f / f184_exposure(c, sun);
    P.alb = img_ptr<uchar4>(c, F184_SLOT_VOX_ALBEDO);
    P.nrm = img_ptr<char4>(c, F184_SLOT_VOX_NORMAL);
    P.shadow = img_ptr<float>(c, F184_SLOT_SHADOW);
    P.rad = img_ptr<uchar4>(c, F184_SLOT_RADIANCE);
    P.rad_surf = c->rad_surf;
    P.brick_list = c->brick_list;
    P.brick_count = c->counters_dev + F184_COUNTER_COUNT;
    P.N = (int)c->cfg.grid_n; P.S = (int)c->cfg.shadow_res;
    rc = f184_stage_begin(c, F184_STAGE_INJECT);
    if (rc) return rc;
    k_inject_n<<<148 * 4, 256, 0, c->stream>>>(P);
    CK_LAUNCH(c);
    return f184_stage_end(c, F184_STAGE_INJECT);
}
