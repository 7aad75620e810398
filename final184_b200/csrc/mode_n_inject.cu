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
    uint32_t* brick_prev;                  // per-brick history bits (f184_internal.h); set_bit = this texture set's bit
    uint32_t set_bit;
    int N, S;
};

constexpr int INJECT_WARPS = 8;

// gamma table: pow(a / 255, 2.2) for the 256 possible albedo bytes, filled on the device with the same dm_pow the
// oracle evaluates per voxel (bit-identical by construction), so the hot loop does a shared-memory lookup
// instead of three log2/exp2 chains per voxel.
__global__ void k_gamma_table(float* __restrict__ table)
{
    table[threadIdx.x] = dm_pow((float)threadIdx.x / 255.0f, 2.2f);
}

// One warp per listed brick.  Phase 1 reads the brick's 512 albedo texels (coalesced rows) and compacts the occupied
// ones through shared memory; phase 2 shades 32 OCCUPIED voxels at a time (in Sponza only ~18 % of the voxels of
// a listed brick are occupied, so shading in place would run the expensive path at 18 % lane efficiency);
// phase 3 writes the whole brick (zeros included: this is also what clears bricks that became empty) as
// 16-byte stores, to the linear level-0 volume and through the surface to the 3D array the tracer filters.
__global__ void __launch_bounds__(INJECT_WARPS * 32) k_inject_n(const InjectParams P, const float* __restrict__ gamma_table)
{
    __shared__ float gam[256];
    __shared__ __align__(16) uint32_t sAlb[INJECT_WARPS][512];     // albedo in, radiance out (same slot)
    __shared__ uint16_t sIdx[INJECT_WARPS][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    gam[threadIdx.x] = gamma_table[threadIdx.x];
    __syncthreads();
    const uint32_t warp_global = blockIdx.x * INJECT_WARPS + warp, n_warps = gridDim.x * INJECT_WARPS;
    const uint32_t count = (uint32_t)*P.brick_count;
    const int N = P.N, NB = N >> 3;
    const float Nf = (float)N;
    uint32_t* sa = sAlb[warp];
    uint16_t* si = sIdx[warp];
    for (uint32_t i = warp_global; i < count; i += n_warps)
    {
        const uint32_t entry = P.brick_list[i];                // bit 31 = touched this frame (normalise)
        const uint32_t b = entry & 0x7fffffffu;
        // this texture set now holds the brick iff it is occupied this frame (an emptied brick is overwritten with zeros below)
        if (lane == 0) P.brick_prev[b] = (P.brick_prev[b] & ~P.set_bit) | ((entry >> 31) ? P.set_bit : 0u);
        const int bx = (b % NB) << 3, by = ((b / NB) % NB) << 3, bz = (b / (NB * NB)) << 3;
        // phase 1: 128 uint4 = 64 rows of 8 texels
        uint32_t nocc = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const int q = lane + 32 * k, row = q >> 1, half = q & 1;
            const int y = row & 7, z = row >> 3;
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(P.alb + ((size_t)(bz + z) * N + (by + y)) * N + bx + half * 4));
            const uint32_t av[4] = {a.x, a.y, a.z, a.w};
            *reinterpret_cast<uint4*>(sa + q * 4) = a;
#pragma unroll
            for (int e = 0; e < 4; e++)
            {
                const bool occ = (av[e] >> 24) != 0;
                const unsigned int m = __ballot_sync(0xffffffffu, occ);
                if (occ) si[nocc + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(q * 4 + e);
                nocc += __popc(m);
            }
        }
        __syncwarp();
        // phase 2
        for (uint32_t j = lane; j < nocc; j += 32)
        {
            const int local = si[j];                              // (z&7)<<6 | (y&7)<<3 | (x&7)
            const int x = bx + (local & 7), y = by + ((local >> 3) & 7), z = bz + (local >> 6);
            const size_t o = ((size_t)z * N + y) * N + x;
            const uint32_t a = sa[local];
            const char4 nq = __ldg(P.nrm + o);
            const float nx = ((float)x + 0.5f) / Nf * 2.0f - 1.0f, ny = ((float)y + 0.5f) / Nf * 2.0f - 1.0f, nz = ((float)z + 0.5f) / Nf;
            const f4 p4 = mul44(P.v2w, f4{nx, ny, nz, 1.0f});
            const f3 p = {p4.x / p4.w, p4.y / p4.w, p4.z / p4.w};
            f3 n = {(float)nq.x / 127.0f, (float)nq.y / 127.0f, (float)nq.z / 127.0f};
            const float nl = length3(n);
            if (nl > 0.0f) n = n / nl;
            const f3 col = {gam[a & 0xffu], gam[(a >> 8) & 0xffu], gam[(a >> 16) & 0xffu]};
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
            const uint32_t q0 = (uint32_t)floorf(dm_clamp(r0 * P.inv_exposure, 0.0f, 1.0f) * 255.0f + 0.5f);
            const uint32_t q1 = (uint32_t)floorf(dm_clamp(r1 * P.inv_exposure, 0.0f, 1.0f) * 255.0f + 0.5f);
            const uint32_t q2 = (uint32_t)floorf(dm_clamp(r2 * P.inv_exposure, 0.0f, 1.0f) * 255.0f + 0.5f);
            sa[local] = q0 | (q1 << 8) | (q2 << 16) | 0xff000000u;
        }
        __syncwarp();
        // phase 3 (unoccupied texels of sa still hold their albedo, whose alpha is 0: they must become 0)
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const int q = lane + 32 * k, row = q >> 1, half = q & 1;
            const int y = row & 7, z = row >> 3;
            uint4 v = *reinterpret_cast<const uint4*>(sa + q * 4);
            v.x = (v.x >> 24) ? v.x : 0u; v.y = (v.y >> 24) ? v.y : 0u; v.z = (v.z >> 24) ? v.z : 0u; v.w = (v.w >> 24) ? v.w : 0u;
            *reinterpret_cast<uint4*>(P.rad + ((size_t)(bz + z) * N + (by + y)) * N + bx + half * 4) = v;
            surf3Dwrite(v, P.rad_surf, (bx + half * 4) * 4, by + y, bz + z);
        }
        __syncwarp();
    }
}

__global__ void k_clear_array(cudaSurfaceObject_t s, int n, int nz)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= n) return;
    for (int z = blockIdx.z; z < nz; z += gridDim.z) surf3Dwrite(make_uchar4(0, 0, 0, 0), s, x * 4, y, z);
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

// Texture-side storage: level 0 as a 3D array, levels >= 1 as ONE mipmapped 3D array holding the six directions as
// z-slabs with zero padding between them (the atlas, f184_internal.h); all surface-writable, sampled with normalised
// coordinates, trilinear, border addressing; two texture objects on the atlas: nearest mip level (DESIGN.md B.5 as amended)
// and linear between levels (SURVEY.md Appendix B as written, F184_FLAG_SPEC_APPENDIX_B).  Two such sets when the frame
// pipeline is on (frame f+1 is built while frame f is traced), one otherwise.
static int alloc_set(f184_ctx* c, VolumeSet& v, int N, uint32_t levels)
{
    cudaChannelFormatDesc ch = cudaCreateChannelDesc<uchar4>();
    CK(c, cudaMalloc3DArray(&v.rad_array, &ch, make_cudaExtent(N, N, N), cudaArraySurfaceLoadStore));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = v.rad_array;
    CK(c, cudaCreateSurfaceObject(&v.rad_surf, &rd));
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    CK(c, cudaCreateTextureObject(&v.rad_tex, &rd, &td, nullptr));
    {
        dim3 g((N + 127) / 128, N, std::min(N, 1024));
        k_clear_array<<<g, 128, 0, c->stream>>>(v.rad_surf, N, N);
        CK_LAUNCH(c);
    }
    // the six-direction atlas (f184_internal.h): extent (N/2, N/2, 12 * N/2) at its level 0, halving with the chain
    CK(c, cudaMallocMipmappedArray(&v.dir_atlas, &ch, make_cudaExtent(N / 2, N / 2, 6 * (size_t)N), levels, cudaArraySurfaceLoadStore));
    for (uint32_t l = 0; l < levels; l++)
    {
        cudaArray_t la;
        CK(c, cudaGetMipmappedArrayLevel(&la, v.dir_atlas, l));
        cudaResourceDesc lrd{};
        lrd.resType = cudaResourceTypeArray;
        lrd.res.array.array = la;
        CK(c, cudaCreateSurfaceObject(&v.dir_surf[l], &lrd));
        const int n = std::max(1, (N / 2) >> l);      // the sparse mip builder never writes unlisted bricks, nobody writes the padding: start from zero
        k_clear_array<<<dim3((n + 127) / 128, n, std::min(12 * n, 1024)), 128, 0, c->stream>>>(v.dir_surf[l], n, 12 * n);
        CK_LAUNCH(c);
    }
    cudaResourceDesc mrd{};
    mrd.resType = cudaResourceTypeMipmappedArray;
    mrd.res.mipmap.mipmap = v.dir_atlas;
    cudaTextureDesc mtd = td;
    mtd.mipmapFilterMode = cudaFilterModePoint;       // nearest level
    mtd.minMipmapLevelClamp = 0.0f;
    mtd.maxMipmapLevelClamp = (float)(levels - 1);
    CK(c, cudaCreateTextureObject(&v.dir_tex, &mrd, &mtd, nullptr));
    mtd.mipmapFilterMode = cudaFilterModeLinear;      // "hardware trilinear across level floor(lod), ceil(lod)"
    CK(c, cudaCreateTextureObject(&v.dir_tex_lin, &mrd, &mtd, nullptr));
    CK(c, cudaEventCreateWithFlags(&v.ev_built, cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&v.ev_traced, cudaEventDisableTiming));
    return F184_OK;
}

int f184_mode_n_alloc(f184_ctx* c)
{
    if (c->n_sets) return F184_OK;
    const int N = (int)c->cfg.grid_n;
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
    const int sets = f184_pipelined(c) ? 2 : 1;
    for (int i = 0; i < sets; i++)
    {
        int rc = alloc_set(c, c->vs[i], N, levels);
        if (rc) return rc;
    }
    CK(c, cudaStreamSynchronize(c->stream));          // first use only: the clears have landed before any stream of the pipeline touches the sets
    c->n_sets = sets;
    c->build_set = c->trace_set = 0;
    return F184_OK;
}

// test hook: what the TEXTURE UNITS return for every texel centre of one direction of one atlas level — the fetch the cone tracer
// makes, at weights 1/0.  (A copy-out with cudaMemcpy3D from the array of a mip level is not used: for the 7.4 GB atlas of a 1024^3
// volume it returns bytes of the wrong level — levels that start beyond 4 GiB inside the mipmapped allocation — while surface
// writes and texture fetches address them correctly: measured, tools/debug_arrays_1024.py.)
__global__ void k_read_atlas_tex(cudaTextureObject_t atlas, float lod, int dir, int n, uchar4* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= n) return;
    const float4 t = tex3DLod<float4>(atlas, ((float)x + 0.5f) / (float)n, ((float)y + 0.5f) / (float)n, ((float)(2 * dir * n + z) + 0.5f) / (float)(12 * n), lod);
    out[((size_t)z * n + y) * n + x] = make_uchar4((unsigned char)(t.x * 255.0f + 0.5f), (unsigned char)(t.y * 255.0f + 0.5f), (unsigned char)(t.z * 255.0f + 0.5f), (unsigned char)(t.w * 255.0f + 0.5f));
}

// test hook: copy one level of the texture-side storage back (dir < 0: the level-0 radiance array) — of the set the next trace samples
extern "C" int f184_debug_read_array(f184_ctx* c, int32_t dir, uint32_t level, void* host, size_t bytes)
{
    if (!c || !host || dir >= 6) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "debug_read_array: bad argument");
    if (!c->n_sets) return f184_fail(c, F184_ERR_NOT_READY, "debug_read_array: no volume yet");
    int rc = f184_join_internal(c);
    if (rc) return rc;
    const VolumeSet& v = c->vs[c->trace_set];
    cudaArray_t a = v.rad_array;
    size_t n = c->cfg.grid_n;
    if (dir >= 0 && !(level & 0x40000000u))
    {
        level &= 0x3fffffffu;
        if (level >= c->n_mip_levels) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "debug_read_array: bad level");
        n = c->mip_levels[level].n;
        if (bytes != n * n * n * 4) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "debug_read_array: size mismatch");
        uchar4* dev = nullptr;
        CK(c, cudaMalloc(&dev, bytes));
        const int bx = (int)std::min<size_t>(n, 128);
        k_read_atlas_tex<<<dim3((unsigned)((n + bx - 1) / bx), (unsigned)n, (unsigned)n), bx, 0, c->stream>>>(v.dir_tex, (float)level, dir, (int)n, dev);
        CK_LAUNCH(c);
        CK(c, cudaStreamSynchronize(c->stream));
        CK(c, cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost));
        cudaFree(dev);
        return F184_OK;
    }
    if (dir >= 0)
    {   // level | 0x40000000: the cudaMemcpy3D copy-out, kept to demonstrate the discrepancy above
        level &= 0x3fffffffu;
        if (level >= c->n_mip_levels) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "debug_read_array: bad level");
        CK(c, cudaGetMipmappedArrayLevel(&a, v.dir_atlas, level));
        n = c->mip_levels[level].n;
    }
    if (bytes != n * n * n * 4) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "debug_read_array: size mismatch");
    CK(c, cudaStreamSynchronize(c->stream));
    cudaMemcpy3DParms p{};
    p.srcArray = a;
    if (dir >= 0) p.srcPos = make_cudaPos(0, 0, 2 * (size_t)dir * n);      // the direction's slab inside the atlas level
    p.dstPtr = make_cudaPitchedPtr(host, n * 4, n, n);
    p.extent = make_cudaExtent(n, n, n);
    p.kind = cudaMemcpyDeviceToHost;
    CK(c, cudaMemcpy3D(&p));
    return F184_OK;
}

int f184_mode_n_release(f184_ctx* c)
{
    for (VolumeSet& v : c->vs)
    {
        if (v.rad_tex) cudaDestroyTextureObject(v.rad_tex);
        if (v.rad_surf) cudaDestroySurfaceObject(v.rad_surf);
        if (v.rad_array) cudaFreeArray(v.rad_array);
        if (v.dir_tex) cudaDestroyTextureObject(v.dir_tex);
        if (v.dir_tex_lin) cudaDestroyTextureObject(v.dir_tex_lin);
        for (int l = 0; l < 12; l++)
            if (v.dir_surf[l]) cudaDestroySurfaceObject(v.dir_surf[l]);
        if (v.dir_atlas) cudaFreeMipmappedArray(v.dir_atlas);
        if (v.ev_built) cudaEventDestroy(v.ev_built);
        if (v.ev_traced) cudaEventDestroy(v.ev_traced);
        v = VolumeSet{};
    }
    c->n_sets = 0;
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
    rc = f184_volume_begin_write(c);       // the trace of two frames ago may still sample the set this frame is built into
    if (rc) return rc;
    c->inject_in_volume = true;
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
    P.rad_surf = c->vs[c->build_set].rad_surf;
    P.brick_prev = c->brick_prev;
    P.set_bit = 2u << c->build_set;
    P.brick_list = c->brick_list;
    P.brick_count = c->counters_dev + F184_COUNTER_COUNT;
    P.N = (int)c->cfg.grid_n; P.S = (int)c->cfg.shadow_res;
    if (!c->gamma_table) CK(c, cudaMalloc(&c->gamma_table, 256 * sizeof(float)));
    if (!c->gamma_ready)
    {
        k_gamma_table<<<1, 256, 0, c->stream>>>(c->gamma_table);
        CK_LAUNCH(c);
        c->gamma_ready = true;
    }
    rc = f184_stage_begin(c, F184_STAGE_INJECT);
    if (rc) return rc;
    k_inject_n<<<148 * 4, INJECT_WARPS * 32, 0, c->stream>>>(P, c->gamma_table);
    CK_LAUNCH(c);
    return f184_stage_end(c, F184_STAGE_INJECT);
}
