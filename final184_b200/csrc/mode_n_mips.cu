// mode_n_mips.cu — six-direction anisotropic mip chain of the radiance volume (DESIGN.md "Mode N" B.4).
//
// The reference has no mip levels at all (Foreground/Renderer/MegaPipeline.cpp:478-482 creates one level and
// no compute shader exists); this is the north-star stage BASELINE.json names: a bandwidth-bound 2x2x2
// reduction, one volume per ray direction (+x,-x,+y,-y,+z,-z), "what a ray travelling along d sees":
//   per 2x2x2 block and axis, each of the 4 columns composites front over back (premultiplied alpha),
//   the 4 columns are averaged.  Integer arithmetic on the 8-bit texels, so it is bit-exact by construction:
//       col = f*255 + (255 - f.a)*b          out = (sum of 4 cols + 510) / 1020
//
// B200 design (HBM-bound: level 1 reads N^3 texels once and writes 6 x (N/2)^3):
//   * TMA path (levels whose source edge >= 64): persistent CTAs, one per SM slot, walk output tiles of
//     32x4x4; the 64x8x8 source box (16 KB) is fetched by ONE cp.async.bulk.tensor.3d into shared memory,
//     3-stage ring, mbarrier complete_tx; every thread then reduces a 8x2x2 strip into 4 x-adjacent
//     outputs per direction and writes them as one 16-byte vector store.  Level 1 produces all six
//     directions from the same tile (the isotropic source is read once).
//   * the small tail levels (source edge < 64, < 0.1 % of the bytes) use a plain per-thread kernel.
//   * each output is written twice: to the linear chain (source of the next level's TMA) and through a
//     surface into the mipmapped 3D atlas the tracer's texture units filter (+12 % write traffic; level 0
//     is never copied).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "f184_device.cuh"

namespace {

__device__ __forceinline__ void acc_col(uint32_t f, uint32_t b, uint32_t sum[4])
{
    const uint32_t fa = f >> 24, ia = 255u - fa;
    sum[0] += (f & 0xffu) * 255u + ia * (b & 0xffu);
    sum[1] += ((f >> 8) & 0xffu) * 255u + ia * ((b >> 8) & 0xffu);
    sum[2] += ((f >> 16) & 0xffu) * 255u + ia * ((b >> 16) & 0xffu);
    sum[3] += fa * 255u + ia * (b >> 24);
}
__device__ __forceinline__ uint32_t finish(const uint32_t sum[4])
{
    return ((sum[0] + 510u) / 1020u) | (((sum[1] + 510u) / 1020u) << 8) | (((sum[2] + 510u) / 1020u) << 16) | (((sum[3] + 510u) / 1020u) << 24);
}
// t[z][y][x] = the 2x2x2 block; direction d = 2*axis + (travelling negative)
__device__ __forceinline__ uint32_t reduce_dir(const uint32_t t[2][2][2], int d)
{
    uint32_t sum[4] = {0, 0, 0, 0};
    const int axis = d >> 1, neg = d & 1;
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int i = 0; i < 2; i++)
        {
            uint32_t f, b;
            if (axis == 0) { f = t[j][i][neg]; b = t[j][i][neg ^ 1]; }          // columns along x: (y=i, z=j)
            else if (axis == 1) { f = t[i][neg][j]; b = t[i][neg ^ 1][j]; }     // along y: (z=i, x=j)
            else { f = t[neg][j][i]; b = t[neg ^ 1][j][i]; }                    // along z: (x=i, y=j)
            acc_col(f, b, sum);
        }
    return finish(sum);
}

struct MipOut
{
    uint32_t* lin[6];
    cudaSurfaceObject_t surf;          // the level's slice of the six-direction atlas (f184_internal.h): direction d at z + 2 d n
};

// ---- plain kernel: one thread per output texel (tail levels, and the cross-check for the TMA path) ----
// iso = 1: all six directions from one source (level 1); iso = 0: direction d reads src[d].
__global__ void __launch_bounds__(256) k_mips_simple(const uint32_t* __restrict__ src0, uint64_t src_dir_stride, MipOut out, int n, int iso)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= n) return;
    const int sn = 2 * n;
    for (int d0 = 0; d0 < 6; d0 += (iso ? 6 : 1))
    {
        const uint32_t* src = src0 + (iso ? 0 : (size_t)d0 * src_dir_stride);
        uint32_t t[2][2][2];
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int j = 0; j < 2; j++)
            {
                const uint2 r = *reinterpret_cast<const uint2*>(src + (((size_t)(2 * z + k)) * sn + (2 * y + j)) * sn + 2 * x);
                t[k][j][0] = r.x; t[k][j][1] = r.y;
            }
        for (int d = d0; d < (iso ? 6 : d0 + 1); d++)
        {
            const uint32_t v = reduce_dir(t, d);
            out.lin[d][((size_t)z * n + y) * n + x] = v;
            surf3Dwrite(v, out.surf, x * 4, y, atlas_z(d, n, z));
        }
    }
}

// ---- TMA path -----------------------------------------------------------------------------------------
constexpr int TX = 32, TY = 4, TZ = 4;                 // output tile
constexpr int SX = 2 * TX, SY = 2 * TY, SZ = 2 * TZ;   // source box 64x8x8 texels
constexpr int TILE_BYTES = SX * SY * SZ * 4;           // 16 KB
constexpr int STAGES = 3;
constexpr int MIPS_THREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z, int w)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// Source tensor: 4-D (x, y, z, dir) of uint32 texels; dir extent is 1 for the isotropic level 0.
// Work items: (dir_src, tile).  iso: one item yields six outputs; else item's dir yields one.
__global__ void __launch_bounds__(MIPS_THREADS)
k_mips_tma(const __grid_constant__ CUtensorMap src_map, MipOut out, int n, int iso, int tiles_x, int tiles_y, int tiles_z, int n_items)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t* tiles = reinterpret_cast<uint32_t*>(smem_raw);                     // STAGES x 16 KB
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * TILE_BYTES);

    const int tid = threadIdx.x;
    if (tid == 0)
    {
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int tiles_per_dir = tiles_x * tiles_y * tiles_z;
    auto issue = [&](int item, int stage) {
        const int dir = item / tiles_per_dir, tl = item % tiles_per_dir;
        const int tx = tl % tiles_x, ty = (tl / tiles_x) % tiles_y, tz = tl / (tiles_x * tiles_y);
        mbar_expect_tx(&full[stage], TILE_BYTES);
        tma_load_4d(tiles + stage * (TILE_BYTES / 4), &src_map, &full[stage], tx * SX, ty * SY, tz * SZ, dir);
    };
    // prologue: fill the ring
    if (tid == 0)
        for (int s = 0; s < STAGES; s++)
        {
            const int item = blockIdx.x + s * gridDim.x;
            if (item < n_items) issue(item, s);
        }
    // thread -> 4 x-adjacent outputs: ox = 4*(tid & 7), oy = (tid >> 3) & 3, oz = tid >> 5
    const int ox = (tid & 7) * 4, oy = (tid >> 3) & 3, oz = tid >> 5;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++)
    {
        const int stage = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1;
        mbar_wait(&full[stage], parity);
        const uint32_t* tile = tiles + stage * (TILE_BYTES / 4);
        // 8x2x2 source strip -> registers
        uint32_t srcv[2][2][8];
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int j = 0; j < 2; j++)
            {
                const uint4* row = reinterpret_cast<const uint4*>(tile + ((2 * oz + k) * SY + (2 * oy + j)) * SX + 2 * ox);
                const uint4 a = row[0], b = row[1];
                srcv[k][j][0] = a.x; srcv[k][j][1] = a.y; srcv[k][j][2] = a.z; srcv[k][j][3] = a.w;
                srcv[k][j][4] = b.x; srcv[k][j][5] = b.y; srcv[k][j][6] = b.z; srcv[k][j][7] = b.w;
            }
        const int dir = item / tiles_per_dir, tl = item % tiles_per_dir;
        const int tx = tl % tiles_x, ty = (tl / tiles_x) % tiles_y, tz = tl / (tiles_x * tiles_y);
        const int gx = tx * TX + ox, gy = ty * TY + oy, gz = tz * TZ + oz;
        const int d_begin = iso ? 0 : dir, d_end = iso ? 6 : dir + 1;
        for (int d = d_begin; d < d_end; d++)
        {
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                uint32_t t[2][2][2];
#pragma unroll
                for (int k = 0; k < 2; k++)
#pragma unroll
                    for (int j = 0; j < 2; j++) { t[k][j][0] = srcv[k][j][2 * q]; t[k][j][1] = srcv[k][j][2 * q + 1]; }
                o[q] = reduce_dir(t, d);
            }
            if (gx < n && gy < n && gz < n)
            {
                *reinterpret_cast<uint4*>(out.lin[d] + ((size_t)gz * n + gy) * n + gx) = make_uint4(o[0], o[1], o[2], o[3]);   // 16-byte store
                surf3Dwrite(make_uint4(o[0], o[1], o[2], o[3]), out.surf, gx * 4, gy, atlas_z(d, n, gz));   // one 16-byte surface store = 4 texels
            }
        }
        __syncthreads();                       // everyone is done with this stage's tile
        if (tid == 0)
        {
            const int next = item + STAGES * gridDim.x;
            if (next < n_items) issue(next, stage);
        }
    }
}


// ---- sparse path: levels 1-3 straight from the listed 8^3 bricks ------------------------------------------
// A 2x2x2 reduction never crosses an 8^3 brick boundary for three levels (8 -> 4 -> 2 -> 1), so one warp takes a
// listed brick (touched this frame or last frame, normalise's list) from the level-0 radiance to its 4^3 x 6
// level-1 texels, 2^3 x 6 level-2 texels and 1 x 6 level-3 texels without any halo; everything not listed is
// zero at every level and is never read or written.  Sponza at 512^3 lists 12 % of the bricks, so this moves
// ~0.1 GB where the dense chain moves 1.4 GB.  Levels >= 4 (edge N/16 and smaller) stay dense.
struct BrickMipOut
{
    uint32_t* lin[3][6];               // level 1..3, six directions, linear chain
    cudaSurfaceObject_t surf[3];       // atlas levels 1..3
};
constexpr int BRICK_WARPS = 4;                                    // 128 threads x 128 registers = 16 K: fits the hole a retiring trace CTA leaves
constexpr int BRICK_STAGES = 3;                                   // TMA ring per warp: 3 x 2 KB bricks in flight
constexpr int BRICK_RING_BYTES = BRICK_WARPS * BRICK_STAGES * 2048;

// TMA = true (default): each warp stages its bricks through a ring in shared memory — one cp.async.bulk.tensor.3d per 8x8x8 box
// of the level-0 volume, completion on an mbarrier — so the 64 row fetches of a brick cost the warp one instruction and the next
// bricks are in flight while this one is reduced.  TMA = false: the same kernel with per-lane 16-byte loads (F184_MIPS_BRICKS_LDG=1;
// also what runs when the driver has no cuTensorMapEncodeTiled).
template <bool TMA>
__global__ void __launch_bounds__(BRICK_WARPS * 32)
k_mips_bricks(const __grid_constant__ CUtensorMap level0_map, const uint32_t* __restrict__ level0, const uint32_t* __restrict__ brick_list,
              const unsigned long long* __restrict__ brick_count, BrickMipOut out, int N, uint32_t* __restrict__ export_buf, uint32_t export_cap,
              uint32_t* __restrict__ brick_prev, uint32_t set_bit)
{
    extern __shared__ __align__(128) uint8_t ring_raw[];            // TMA: [warp][stage][512 texels]; LDG: [warp][512 texels]
    __shared__ __align__(16) uint32_t sh1[BRICK_WARPS][6][64];     // level 1 per direction, [z][y][x] 4^3
    __shared__ __align__(16) uint32_t sh2[BRICK_WARPS][6][8];      // level 2 per direction, 2^3
    __shared__ __align__(8) uint64_t bars[BRICK_WARPS][BRICK_STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warp_global = blockIdx.x * BRICK_WARPS + warp, n_warps = gridDim.x * BRICK_WARPS;
    const uint32_t count = (uint32_t)*brick_count;
    const int NB = N >> 3, n1 = N >> 1, n2 = N >> 2, n3 = N >> 3;
    uint32_t* ring = reinterpret_cast<uint32_t*>(ring_raw) + (size_t)warp * (TMA ? BRICK_STAGES : 1) * 512;
    auto issue = [&](uint32_t i, int stage) {       // lane 0: fetch brick i of the list into `stage`
        const uint32_t b = __ldg(brick_list + i) & 0x7fffffffu;
        mbar_expect_tx(&bars[warp][stage], 2048);
        tma_load_3d(ring + stage * 512, &level0_map, &bars[warp][stage], (int)(b % NB) * 8, (int)((b / NB) % NB) * 8, (int)(b / (NB * NB)) * 8);
    };
    if (TMA)
    {
        if (lane == 0)
        {
            for (int s_ = 0; s_ < BRICK_STAGES; s_++) mbar_init(&bars[warp][s_], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (lane == 0)
            for (int s_ = 0; s_ < BRICK_STAGES; s_++)
            {
                const uint32_t i = warp_global + (uint32_t)s_ * n_warps;
                if (i < count) issue(i, s_);
            }
    }
    uint32_t it = 0;
    for (uint32_t i = warp_global; i < count; i += n_warps, it++)
    {
        const uint32_t entry = __ldg(brick_list + i);
        const uint32_t b = entry & 0x7fffffffu;
        // levels 1-3 of this texture set now hold the brick iff it is occupied this frame (same assignment as k_inject_n makes for level 0)
        if (lane == 0 && brick_prev) brick_prev[b] = (brick_prev[b] & ~set_bit) | ((entry >> 31) ? set_bit : 0u);
        const int bx = b % NB, by = (b / NB) % NB, bz = b / (NB * NB);
        const int stage = TMA ? (int)(it % BRICK_STAGES) : 0;
        uint32_t* s0 = ring + stage * 512;                          // the brick, [z][y][x]
        if (TMA)
        {
            mbar_wait(&bars[warp][stage], (it / BRICK_STAGES) & 1u);
            // multi-GPU: the brick's level 0 into the export array the other ranks pull from (see below)
            if (export_buf)
#pragma unroll
                for (int k = 0; k < 4; k++)
                    reinterpret_cast<uint4*>(export_buf + (size_t)i * 512)[lane + 32 * k] = reinterpret_cast<const uint4*>(s0)[lane + 32 * k];
        }
        else
        {
            // level 0: 64 rows of 8 texels (one 32-byte sector each) = 128 uint4, four per lane
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                const int q = lane + 32 * k, row = q >> 1, half = q & 1;
                const int y = row & 7, z = row >> 3;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(level0 + ((size_t)(bz * 8 + z) * N + (by * 8 + y)) * N + bx * 8 + half * 4));
                *reinterpret_cast<uint4*>(s0 + row * 8 + half * 4) = v;
                if (export_buf) reinterpret_cast<uint4*>(export_buf + (size_t)i * 512)[q] = v;
            }
        }
        // multi-GPU: what the other ranks fetch over NVLink (mode_n_shard.cu), three arrays indexed by the brick's position in the list —
        // level 0 (512 words per brick; only glossy scenes make anybody read it), level 1 (`fine`, 384 words; read by the ranks whose
        // cones sample it) and a 64-word `coarse` block everybody reads: level 2, level 3, the brick's index.  Coarse blocks of
        // consecutive bricks are contiguous, so a reader moves 64 of them with one bulk copy.
        uint32_t* fine = export_buf ? export_buf + (size_t)export_cap * 512 + (size_t)i * 384 : nullptr;
        uint32_t* coarse = export_buf ? export_buf + (size_t)export_cap * 896 + (size_t)i * 64 : nullptr;
        if (coarse && lane == 0) coarse[54] = b;
        __syncwarp();
        {   // level 1: lane -> output row (oy, oz) = lane & 15 and three of the six directions
            const int p = lane & 15, oy = p & 3, oz = p >> 2, d0 = (lane >> 4) * 3;
            uint32_t src[2][2][8];
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int j = 0; j < 2; j++)
                {
                    const uint4* row = reinterpret_cast<const uint4*>(s0 + ((2 * oz + k) * 8 + (2 * oy + j)) * 8);
                    const uint4 a = row[0], c = row[1];
                    src[k][j][0] = a.x; src[k][j][1] = a.y; src[k][j][2] = a.z; src[k][j][3] = a.w;
                    src[k][j][4] = c.x; src[k][j][5] = c.y; src[k][j][6] = c.z; src[k][j][7] = c.w;
                }
            const int gx = bx * 4, gy = by * 4 + oy, gz = bz * 4 + oz;
#pragma unroll
            for (int dd = 0; dd < 3; dd++)
            {
                const int d = d0 + dd;
                uint32_t o[4];
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    uint32_t t[2][2][2];
#pragma unroll
                    for (int k = 0; k < 2; k++)
#pragma unroll
                        for (int j = 0; j < 2; j++) { t[k][j][0] = src[k][j][2 * q]; t[k][j][1] = src[k][j][2 * q + 1]; }
                    o[q] = reduce_dir(t, d);
                }
                const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(&sh1[warp][d][(oz * 4 + oy) * 4]) = v;
                if (fine) *reinterpret_cast<uint4*>(fine + d * 64 + (oz * 4 + oy) * 4) = v;
                *reinterpret_cast<uint4*>(out.lin[0][d] + ((size_t)gz * n1 + gy) * n1 + gx) = v;
                surf3Dwrite(v, out.surf[0], gx * 4, gy, atlas_z(d, n1, gz));
            }
        }
        __syncwarp();
        if (lane < 24)
        {   // level 2: lane -> (direction, output row of 2): reads only its own direction's level 1
            const int d = lane >> 2, p = lane & 3, oy = p & 1, oz = p >> 1;
            uint32_t o[2];
            uint32_t src[2][2][4];
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int j = 0; j < 2; j++)
                {
                    const uint4 a = *reinterpret_cast<const uint4*>(&sh1[warp][d][((2 * oz + k) * 4 + (2 * oy + j)) * 4]);
                    src[k][j][0] = a.x; src[k][j][1] = a.y; src[k][j][2] = a.z; src[k][j][3] = a.w;
                }
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
                uint32_t t[2][2][2];
#pragma unroll
                for (int k = 0; k < 2; k++)
#pragma unroll
                    for (int j = 0; j < 2; j++) { t[k][j][0] = src[k][j][2 * q]; t[k][j][1] = src[k][j][2 * q + 1]; }
                o[q] = reduce_dir(t, d);
            }
            const int gx = bx * 2, gy = by * 2 + oy, gz = bz * 2 + oz;
            const uint2 v = make_uint2(o[0], o[1]);
            *reinterpret_cast<uint2*>(&sh2[warp][d][(oz * 2 + oy) * 2]) = v;
            if (coarse) *reinterpret_cast<uint2*>(coarse + d * 8 + (oz * 2 + oy) * 2) = v;
            *reinterpret_cast<uint2*>(out.lin[1][d] + ((size_t)gz * n2 + gy) * n2 + gx) = v;
            surf3Dwrite(v, out.surf[1], gx * 4, gy, atlas_z(d, n2, gz));
        }
        __syncwarp();
        if (lane < 6)
        {   // level 3: one texel per direction
            const int d = lane;
            uint32_t t[2][2][2];
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int j = 0; j < 2; j++) { t[k][j][0] = sh2[warp][d][(k * 2 + j) * 2]; t[k][j][1] = sh2[warp][d][(k * 2 + j) * 2 + 1]; }
            const uint32_t v = reduce_dir(t, d);
            out.lin[2][d][((size_t)bz * n3 + by) * n3 + bx] = v;
            if (coarse) coarse[48 + d] = v;
            surf3Dwrite(v, out.surf[2], bx * 4, by, atlas_z(d, n3, bz));
        }
        __syncwarp();                              // every lane is done with this stage's brick (and with sh1 / sh2)
        if (TMA && lane == 0)
        {
            const uint32_t next = i + (uint32_t)BRICK_STAGES * n_warps;
            if (next < count)
            {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the warp's reads of the stage before the async proxy overwrites it
                issue(next, stage);
            }
        }
    }
}

// ---- tail: every level below level 3 in ONE launch -------------------------------------------------------------
// Level 3 has edge N/8 (64 at 512^3): the rest of the chain is < 0.1 % of the bytes, and as one launch per level it was
// six launch latencies (42 us of a 4 ms frame).  Here a CTA takes a T^3 tile (T = min(16, N/8)) of one direction's
// level 3 through shared memory down log2(T) levels; whatever is left (edge N/128: 4^3 at 512^3) is finished by the
// last CTA to arrive (ticket counter, self-resetting).
struct TailOut
{
    cudaSurfaceObject_t surf[12];           // atlas level l
    uint64_t level_off[12];                 // texel offset of level l+1, direction 0, inside the chain allocation
    uint32_t level_n[12];
    uint32_t n_levels;
};

__global__ void __launch_bounds__(256)
k_mips_tail(uint32_t* __restrict__ chain, TailOut out, int T, int tiles, unsigned long long* __restrict__ ticket)
{
    __shared__ uint32_t bufA[4096], bufB[512];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int d = blockIdx.x / (tiles * tiles * tiles), tl = blockIdx.x % (tiles * tiles * tiles);
    const int tx = tl % tiles, ty = (tl / tiles) % tiles, tz = tl / (tiles * tiles);
    const int n3 = (int)out.level_n[2];
    {
        const uint32_t* src = chain + out.level_off[2] + (uint64_t)d * n3 * n3 * n3;
        for (int i = tid; i < T * T * T; i += 256)
        {
            const int x = i % T, y = (i / T) % T, z = i / (T * T);
            bufA[i] = src[((size_t)(tz * T + z) * n3 + (ty * T + y)) * n3 + tx * T + x];
        }
    }
    __syncthreads();
    uint32_t *a = bufA, *b = bufB;
    int lvl = 3;                                            // index of the level being WRITTEN (0 = level 1)
    for (int s = T; s > 1; s >>= 1, lvl++)
    {
        const int m = s >> 1, n = (int)out.level_n[lvl];
        uint32_t* dst = chain + out.level_off[lvl] + (uint64_t)d * n * n * n;
        for (int i = tid; i < m * m * m; i += 256)
        {
            const int x = i % m, y = (i / m) % m, z = i / (m * m);
            uint32_t t[2][2][2];
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int q = 0; q < 2; q++) t[k][j][q] = a[((2 * z + k) * s + (2 * y + j)) * s + 2 * x + q];
            const uint32_t v = reduce_dir(t, d);
            b[i] = v;
            const int gx = tx * m + x, gy = ty * m + y, gz = tz * m + z;
            dst[((size_t)gz * n + gy) * n + gx] = v;
            surf3Dwrite(v, out.surf[lvl], gx * 4, gy, atlas_z(d, n, gz));
        }
        __syncthreads();
        uint32_t* tmp = a; a = b; b = tmp;
    }
    if ((uint32_t)lvl >= out.n_levels) return;             // the tile reached 1^3: the chain is complete
    // hand-over: the last CTA finishes the remaining levels from global memory
    __threadfence();
    __syncthreads();
    if (tid == 0)
    {
        const unsigned long long tk = atomicAdd(ticket, 1ull);
        is_last = (tk == (unsigned long long)gridDim.x - 1ull);
        if (is_last) *ticket = 0ull;
    }
    __syncthreads();
    if (!is_last) return;
    for (; (uint32_t)lvl < out.n_levels; lvl++)
    {
        const int n = (int)out.level_n[lvl], sn = 2 * n;
        for (int i = tid; i < 6 * n * n * n; i += 256)
        {
            const int dd = i / (n * n * n), r = i % (n * n * n);
            const int x = r % n, y = (r / n) % n, z = r / (n * n);
            const uint32_t* src = chain + out.level_off[lvl - 1] + (uint64_t)dd * sn * sn * sn;
            uint32_t t[2][2][2];
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int q = 0; q < 2; q++) t[k][j][q] = __ldcg(src + ((size_t)(2 * z + k) * sn + (2 * y + j)) * sn + 2 * x + q);
            uint32_t v = 0;
#pragma unroll
            for (int e = 0; e < 6; e++)
                if (e == dd) v = reduce_dir(t, e);
            chain[out.level_off[lvl] + (uint64_t)dd * n * n * n + r] = v;
            surf3Dwrite(v, out.surf[lvl], x * 4, y, atlas_z(dd, n, z));
        }
        __threadfence();
        __syncthreads();
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn)
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

}  // namespace

// every level below level 3, one launch (timed with the mips stage on one GPU, after the gather on several)
int f184_mips_tail_n(f184_ctx* c, bool own_stage)
{
    if (c->n_mip_levels <= 3) return F184_OK;
    uint32_t* mips = img_ptr<uint32_t>(c, F184_SLOT_MIPS);
    TailOut to{};
    to.n_levels = c->n_mip_levels;
    for (uint32_t l = 0; l < c->n_mip_levels; l++) { to.level_off[l] = c->mip_levels[l].offset_texels; to.level_n[l] = c->mip_levels[l].n; }
    for (uint32_t l = 0; l < c->n_mip_levels; l++) to.surf[l] = c->vs[c->build_set].dir_surf[l];
    const int n3 = (int)c->mip_levels[2].n, T = std::min(16, n3), tiles = n3 / T;
    if (own_stage)
    {
        int rc = f184_stage_begin(c, F184_STAGE_MIPS);
        if (rc) return rc;
    }
    k_mips_tail<<<6 * tiles * tiles * tiles, 256, 0, c->stream>>>(mips, to, T, tiles, c->counters_dev + F184_COUNTER_COUNT + 2);
    CK_LAUNCH(c);
    return own_stage ? f184_stage_end(c, F184_STAGE_MIPS) : F184_OK;
}

int f184_mips_init_n(f184_ctx* c)
{
    static bool done = false;
    if (done) return F184_OK;
    if (get_encode() != nullptr)
    {
        CK(c, cudaFuncSetAttribute(k_mips_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * TILE_BYTES + 64));
        CK(c, cudaFuncSetAttribute(k_mips_bricks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BRICK_RING_BYTES));
    }
    done = true;
    return F184_OK;
}

int f184_mips_n(f184_ctx* c)
{
    int rc = f184_ensure_image(c, F184_SLOT_RADIANCE); if (rc) return rc;
    rc = f184_ensure_image(c, F184_SLOT_MIPS); if (rc) return rc;
    rc = f184_mode_n_alloc(c); if (rc) return rc;
    rc = f184_volume_begin_write(c); if (rc) return rc;
    VolumeSet& vs = c->vs[c->build_set];
    const int N = (int)c->cfg.grid_n;
    uint32_t* mips = img_ptr<uint32_t>(c, F184_SLOT_MIPS);
    const uint32_t* level0 = img_ptr<uint32_t>(c, F184_SLOT_RADIANCE);
    const bool use_tma = !(c->cfg.flags & F184_FLAG_NO_TMA) && get_encode() != nullptr;
    const int smem = STAGES * TILE_BYTES + 64;
    rc = f184_mips_init_n(c); if (rc) return rc;
    rc = f184_stage_begin(c, F184_STAGE_MIPS);
    if (rc) return rc;
    uint32_t first_dense = 0;
    const bool sparse = !(c->cfg.flags & (F184_FLAG_NO_TMA | F184_FLAG_DENSE_MIPS)) && c->brick_list && c->n_mip_levels >= 3;
    if (sparse)
    {   // levels 1-3 from the brick list (F184_FLAG_NO_TMA keeps the all-dense plain path as the cross-check)
        BrickMipOut bo;
        for (int l = 0; l < 3; l++)
        {
            bo.surf[l] = vs.dir_surf[l];
            for (int d = 0; d < 6; d++)
            {
                const uint64_t n = c->mip_levels[l].n;
                bo.lin[l][d] = mips + c->mip_levels[l].offset_texels + (uint64_t)d * n * n * n;
            }
        }
        uint32_t* export_buf = nullptr;
        const uint32_t export_cap = ((uint32_t)(N / 8) * (N / 8) * (N / 8)) / (c->cfg.nranks ? c->cfg.nranks : 1);       // bricks the export arrays hold
        if (c->cfg.nranks > 1)
        {
            void* e = nullptr;
            rc = f184_ipc_buffer_ptr(c, F184_IPC_EXPORT, &e);
            if (rc) return rc;
            export_buf = reinterpret_cast<uint32_t*>(e);
        }
        // (the history bit is assigned only when f184_inject wrote level 0 of the same set from the same list: a build_mips on its
        // own must not declare a stale level 0 clean)
        static const bool force_ldg = [] { const char* e = getenv("F184_MIPS_BRICKS_LDG"); return e && atoi(e) != 0; }();
        uint32_t* prev = c->inject_in_volume ? c->brick_prev : nullptr;
        CUtensorMap map{};
        bool tma_bricks = get_encode() != nullptr && !force_ldg;
        if (tma_bricks)
        {   // the level-0 volume as a 3-D tensor of texels, fetched in 8 x 8 x 8 boxes
            const cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)N};
            const cuuint64_t gstride[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * N * 4};
            const cuuint32_t box[3] = {8, 8, 8};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = get_encode()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)level0, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) tma_bricks = false;
        }
        if (tma_bricks)
        {
            k_mips_bricks<true><<<148 * 4, BRICK_WARPS * 32, BRICK_RING_BYTES, c->stream>>>(map, level0, c->brick_list, c->counters_dev + F184_COUNTER_COUNT, bo, N,
                                                                                           export_buf, export_cap, prev, 2u << c->build_set);
        }
        else
            k_mips_bricks<false><<<148 * 8, BRICK_WARPS * 32, BRICK_WARPS * 2048, c->stream>>>(map, level0, c->brick_list, c->counters_dev + F184_COUNTER_COUNT, bo, N,
                                                                                              export_buf, export_cap, prev, 2u << c->build_set);
        CK_LAUNCH(c);
        first_dense = c->n_mip_levels;          // the tail kernel takes every remaining level
        if (c->cfg.nranks <= 1)                 // (multi-GPU: level 3 is complete only after f184_gather_volume, which runs the tail)
        {
            rc = f184_mips_tail_n(c, false);
            if (rc) return rc;
        }
    }
    for (uint32_t li = first_dense; li < c->n_mip_levels; li++)
    {
        const int n = (int)c->mip_levels[li].n, sn = 2 * n;
        const int iso = (li == 0);
        const uint32_t* src = iso ? level0 : mips + c->mip_levels[li - 1].offset_texels;
        const uint64_t src_dir_stride = (uint64_t)sn * sn * sn;
        MipOut out;
        out.surf = vs.dir_surf[li];
        for (int d = 0; d < 6; d++) out.lin[d] = mips + c->mip_levels[li].offset_texels + (uint64_t)d * n * n * n;
        if (use_tma && sn >= 64)
        {
            CUtensorMap map;
            const cuuint64_t gdim[4] = {(cuuint64_t)sn, (cuuint64_t)sn, (cuuint64_t)sn, (cuuint64_t)(iso ? 1 : 6)};
            const cuuint64_t gstride[3] = {(cuuint64_t)sn * 4, (cuuint64_t)sn * sn * 4, (cuuint64_t)sn * sn * sn * 4};
            const cuuint32_t box[4] = {SX, SY, SZ, 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = get_encode()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, (void*)src, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return f184_fail(c, F184_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for level %u", (int)r, li + 1);
            const int tiles_x = (n + TX - 1) / TX, tiles_y = (n + TY - 1) / TY, tiles_z = (n + TZ - 1) / TZ;
            const int n_items = tiles_x * tiles_y * tiles_z * (iso ? 1 : 6);
            const int grid = std::min(n_items, 148 * 4);
            k_mips_tma<<<grid, MIPS_THREADS, smem, c->stream>>>(map, out, n, iso, tiles_x, tiles_y, tiles_z, n_items);
            CK_LAUNCH(c);
        }
        else
        {
            const int bx = std::min(n, 256);
            dim3 grid((n + bx - 1) / bx, n, n);
            k_mips_simple<<<grid, bx, 0, c->stream>>>(src, src_dir_stride, out, n, iso);
            CK_LAUNCH(c);
        }
    }
    rc = f184_stage_end(c, F184_STAGE_MIPS);
    if (rc) return rc;
    // one GPU: the volume of this frame is complete — it becomes the set the next trace samples.  One NVLink box: only after
    // f184_gather_volume has fetched the other ranks' bricks.
    return c->cfg.nranks <= 1 ? f184_volume_publish(c) : F184_OK;
}
