// mode_n_shard.cu — one NVLink box, one process per GPU: the pieces of the north-star schedule that touch peer memory
// (include/f184.h "one NVLink box"; DESIGN.md "Multi-GPU").  The reference is single-GPU and single-queue
// (RHI/Private/Vulkan/DeviceVk.cpp:301-304); nothing here has a counterpart in it.
//
//   * the reduce-scatter of the partial volumes is not a separate collective: mode_n_voxelize.cu reduces every fragment
//     straight into the accumulators of the rank that owns the fragment's Z-slab (red.global.add.v4.f32 on a
//     peer-mapped pointer, carried by NVLink), so after one barrier each owner holds the exact sum for its slab;
//   * k_peer_barrier: device-side flag barrier — each rank stores its epoch into every peer's flag array
//     (st.release.sys over NVLink) and spins on its own array (ld.acquire.sys, local memory): no host round trip and
//     no NCCL launch on the frame's critical path; a timeout (5 s; F184_BARRIER_TIMEOUT_MS) turns a missing peer into an
//     error instead of a hang: the kernel sets a sticky device error bit and the next synchronous call of the context
//     (f184_sync, f184_counter_get, f184_stage_time_*) returns F184_ERR_PEER_TIMEOUT;
//   * k_gather_bricks: the all-gather before tracing — every rank pulls the other ranks' finished bricks, packed by
//     mode_n_mips.cu into contiguous 4 KB records (level 0 + levels 1-3 of one 8^3 brick), with coalesced 16-byte peer
//     loads, and writes them through surfaces into its own texture storage; only listed bricks move (Sponza at 512^3:
//     ~0.12 GB for the whole volume instead of 1.0 GB dense).  Levels >= 4 are then finished locally (k_mips_tail);
//   * level 0 travels only when a cone of this rank's rows can sample it: k_need_level0 evaluates the tracer's own level
//     selection for the first (finest) sample of every pixel's cones; with the 60-degree diffuse cones and materials of
//     roughness >= 0.6 no cone ever reads level 0 and the gather moves 2 KB per brick instead of 4.
#include <algorithm>
#include <cstdlib>

#include "f184_device.cuh"
#include "f184_cone.cuh"

int f184_mips_tail_n(f184_ctx* c, bool own_stage);      // mode_n_mips.cu

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct BarrierArgs
{
    uint32_t* flags[8];        // flags[p] = rank p's flag array (own entry: local memory)
    int rank, nranks;
    uint32_t epoch;
    unsigned long long timeout_ns;
    uint32_t* dev_state;
};

__global__ void k_peer_barrier(const BarrierArgs B)
{
    const int p = threadIdx.x;
    if (p >= B.nranks || p == B.rank) return;
    __threadfence_system();                                    // everything this GPU wrote (incl. peer atomics) before the flag
    st_release_sys(B.flags[p] + B.rank, B.epoch);
    const uint32_t* mine = B.flags[B.rank] + p;
    const unsigned long long t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - B.epoch) < 0)
    {
        if (globaltimer_ns() - t0 > B.timeout_ns)
        {   // no hang: a sticky error bit that the next synchronous call of the context reports (F184_ERR_PEER_TIMEOUT), and a marker
            // naming the peer
            atomicOr(B.dev_state + F184_DEV_ERROR, F184_DEVERR_BARRIER_TIMEOUT);
            B.flags[B.rank][8 + p] = B.epoch;
            break;
        }
        __nanosleep(200);
    }
}

// Does any cone of this rank's rows sample level 0 of the volume?  Evaluates the tracer's level selection (f184_cone.cuh) at the
// first sample of each cone — the finest one: the footprint only grows with distance.  The diffuse cones are the same for
// every pixel; the specular cone depends on the pixel's roughness.
struct NeedArgs
{
    const float* depth;
    const uchar4* material;
    uint32_t W, y0, y1, tile0, tile_stride, n_tiles;
    float h;
    uint32_t spec_b;
    uint32_t* dev_state;
};
__global__ void __launch_bounds__(256) k_need_level0(const NeedArgs A)
{
    bool need = false;
    if (blockIdx.x == 0 && threadIdx.x == 0) need = cone_samples_level0(kTanHalfDiffuse, A.h, A.spec_b != 0);
    for (uint32_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x)
    {
        const uint32_t yb = A.y0 + (A.tile0 + t * A.tile_stride) * 8;
        for (uint32_t i = threadIdx.x; i < 8 * A.W; i += blockDim.x)
        {
            const uint32_t y = yb + i / A.W, x = i % A.W;
            if (y >= A.y1) break;
            if (__ldg(A.depth + (size_t)y * A.W + x) >= 1.0f) continue;
            const float rough = (float)__ldg(A.material + (size_t)y * A.W + x).y / 255.0f;
            need |= cone_samples_level0(cone_specular_tan(rough), A.h, A.spec_b != 0);
        }
    }
    if (__syncthreads_or(need) && threadIdx.x == 0) A.dev_state[F184_DEV_NEED_L0] = 1u;
}

// Level 0 is about to be gathered into a set whose level 0 was skipped by an earlier gather: the bricks of the OTHER ranks may hold
// anything (stale frames), and bricks that have emptied since are no longer on any list.  Clear them all once; own bricks are
// always current (f184_inject writes them).  Exits at once in every other frame.
__global__ void __launch_bounds__(256) k_clear_foreign_level0(cudaSurfaceObject_t rad_surf, int N, uint32_t G, uint32_t rank, const uint32_t* __restrict__ dev_state, int set)
{
    if (!dev_state[F184_DEV_NEED_L0] || dev_state[F184_DEV_L0_FULL + set]) return;
    const uint32_t NB = (uint32_t)N >> 3, n_bricks = NB * NB * NB;
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (uint32_t b = warp_global; b < n_bricks; b += n_warps)
    {
        const uint32_t bx = b % NB, by = (b / NB) % NB, bz = b / (NB * NB);
        if (((bx + by + bz) & (G - 1)) == rank) continue;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const int q = lane + 32 * k, row = q >> 1, half = q & 1, y = row & 7, z = row >> 3;
            surf3Dwrite(zero, rad_surf, (bx * 8 + half * 4) * 4, by * 8 + y, bz * 8 + z);
        }
    }
}


struct GatherArgs
{
    const uint32_t* peer_export[8];
    const unsigned long long* peer_counters[8];
    const uint32_t* peer_list[8];
    const uint32_t* dev_state;
    int rank, nranks, N, write_linear;
    cudaSurfaceObject_t rad_surf;
    uint32_t* rad_lin;
    uint32_t* lin[3][6];
    cudaSurfaceObject_t surf[3];       // atlas levels 1..3 (direction d at z + 2 d n)
};

// behind the gather: the set's level-0 bookkeeping, and the bytes that crossed NVLink (F184_COUNTER_GATHER_BYTES)
__global__ void k_gather_state(const GatherArgs G, uint32_t* dev_state, int set, unsigned long long* counters, uint32_t bytes_with_l0, uint32_t bytes_without)
{
    const bool level0 = dev_state[F184_DEV_NEED_L0] != 0;
    dev_state[F184_DEV_L0_FULL + set] = level0 ? 1u : 0u;
    unsigned long long records = 0;
    for (int p = 0; p < G.nranks; p++)
        if (p != G.rank) records += G.peer_counters[p][F184_COUNTER_COUNT];
    counters[F184_COUNTER_GATHER_BYTES] = records * (unsigned long long)(level0 ? bytes_with_l0 : bytes_without);
}

constexpr int GATHER_WARPS = 8;

__global__ void __launch_bounds__(GATHER_WARPS * 32) k_gather_bricks(const GatherArgs G)
{
    // grid = (nranks, slices): consecutive CTAs read from DIFFERENT peers, and the peer order is rotated by the reader's rank.
    // With the peer in the slow grid dimension every GPU of the box read from peer 0 first, then from peer 1, ... : seven
    // readers on one GPU's NVLink egress at a time while the other links idled (111 MB per rank at ~320 GB/s at 8 GPUs against
    // 630 GB/s at 2).  Now each wave of CTAs covers all peers, and reader r starts at peer r+1.
    const int p = (int)((blockIdx.x + (unsigned)G.rank) % (unsigned)G.nranks);
    if (p == G.rank) return;                                   // blockIdx.x == 0
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warp_global = blockIdx.y * GATHER_WARPS + warp, n_warps = gridDim.y * GATHER_WARPS;
    const uint32_t count = (uint32_t)G.peer_counters[p][F184_COUNTER_COUNT];     // that rank's brick-list cursor
    const int N = G.N, NB = N >> 3, n1 = N >> 1, n2 = N >> 2, n3 = N >> 3;
    const bool level0 = G.dev_state[F184_DEV_NEED_L0] != 0;       // block-uniform: level 0 travels only when a cone of this rank's rows samples it
    for (uint32_t i = warp_global; i < count; i += n_warps)
    {
        const uint32_t* rec = G.peer_export[p] + (size_t)i * 1024;
        const uint4* rec4 = reinterpret_cast<const uint4*>(rec);
        const uint32_t b = G.peer_list[p][i] & 0x7fffffffu;
        const int bx = b % NB, by = (b / NB) % NB, bz = b / (NB * NB);
        uint4 l0[4], l1[3];
        if (level0)
        {
#pragma unroll
            for (int k = 0; k < 4; k++) l0[k] = rec4[lane + 32 * k];           // 7 independent 16-byte peer loads in flight
        }
#pragma unroll
        for (int k = 0; k < 3; k++) l1[k] = rec4[128 + lane + 32 * k];
        uint2 l2 = make_uint2(0, 0);
        if (lane < 24) l2 = reinterpret_cast<const uint2*>(rec + 896)[lane];
        uint32_t l3 = 0;
        if (lane < 6) l3 = rec[944 + lane];
        if (level0)
        {
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                const int q = lane + 32 * k, row = q >> 1, half = q & 1, y = row & 7, z = row >> 3;
                surf3Dwrite(l0[k], G.rad_surf, (bx * 8 + half * 4) * 4, by * 8 + y, bz * 8 + z);
                if (G.write_linear) *reinterpret_cast<uint4*>(G.rad_lin + ((size_t)(bz * 8 + z) * N + (by * 8 + y)) * N + bx * 8 + half * 4) = l0[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            const int j = lane + 32 * k, d = j >> 4, r = j & 15, oy = r & 3, oz = r >> 2;
            const int gx = bx * 4, gy = by * 4 + oy, gz = bz * 4 + oz;
            surf3Dwrite(l1[k], G.surf[0], gx * 4, gy, atlas_z(d, n1, gz));
            if (G.write_linear) *reinterpret_cast<uint4*>(G.lin[0][d] + ((size_t)gz * n1 + gy) * n1 + gx) = l1[k];
        }
        if (lane < 24)
        {
            const int d = lane >> 2, r = lane & 3, oy = r & 1, oz = r >> 1;
            const int gx = bx * 2, gy = by * 2 + oy, gz = bz * 2 + oz;
            surf3Dwrite(l2, G.surf[1], gx * 4, gy, atlas_z(d, n2, gz));
            if (G.write_linear) *reinterpret_cast<uint2*>(G.lin[1][d] + ((size_t)gz * n2 + gy) * n2 + gx) = l2;
        }
        if (lane < 6)
        {
            surf3Dwrite(l3, G.surf[2], bx * 4, by, atlas_z(lane, n3, bz));
            G.lin[2][lane][((size_t)bz * n3 + by) * n3 + bx] = l3;             // level 3 is the source of the local tail: always
        }
    }
}

// ---- the gather, TMA-fed (default) ---------------------------------------------------------------------------------------------
// The per-lane peer loads above keep 7 x 16 bytes per lane in flight and need every warp of the GPU to cover NVLink's ~2 us: fine
// when the gather has the GPU to itself (630 GB/s at 2 GPUs), but inside the frame pipeline it shares the SMs with the cone trace
// of the previous frame and ran at 70 GB/s.  Here ONE thread per CTA issues bulk copies (cp.async.bulk: the TMA engine moves a
// brick's whole record, 1.7 KB or 3.7 KB, peer HBM -> shared memory over NVLink, completion on an mbarrier) into a 16-slot ring —
// 28-60 KB in flight per CTA whatever else runs on the SM — and four consumer warps write the landed records through surfaces into
// the texture storage.  Records carry their brick index (word 950, written by k_mips_bricks), so nothing but bulk copies crosses
// NVLink.
constexpr int G2_SLOTS = 16;
constexpr int G2_CONSUMERS = 4;                  // warps; + 1 producer warp
constexpr int G2_SLOT_BYTES = 4096;
constexpr int G2_REC_BYTES = 1760;               // levels 1-3 + brick index: words [512, 952)
constexpr int G2_REC0_BYTES = 3808;              // with level 0: words [0, 952)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// bounded wait: a protocol error or a dead peer becomes a sticky device error (reported by the next synchronous call), not a hung GPU
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, uint32_t* dev_state)
{
    unsigned long long t0 = 0;
    for (uint32_t tries = 0;; tries++)
    {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
        if ((tries & 255u) == 255u)
        {
            const unsigned long long now = globaltimer_ns();
            if (!t0) t0 = now;
            else if (now - t0 > 2000000000ull) break;               // 2 s
        }
    }
    atomicOr(dev_state + F184_DEV_ERROR, F184_DEVERR_BARRIER_TIMEOUT);
    return false;
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__((G2_CONSUMERS + 1) * 32) k_gather_bricks_tma(const GatherArgs G, uint32_t* dev_state)
{
    extern __shared__ __align__(128) uint8_t ring[];                // G2_SLOTS x 4 KB
    __shared__ __align__(8) uint64_t full[G2_SLOTS], empty[G2_SLOTS];
    __shared__ uint32_t counts[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
    {
        for (int s_ = 0; s_ < G2_SLOTS; s_++) { mbar_init(&full[s_], 1); mbar_init(&empty[s_], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 8) counts[threadIdx.x] = ((int)threadIdx.x < G.nranks && (int)threadIdx.x != G.rank) ? (uint32_t)G.peer_counters[threadIdx.x][F184_COUNTER_COUNT] : 0u;
    __syncthreads();
    const bool level0 = G.dev_state[F184_DEV_NEED_L0] != 0;
    uint32_t max_count = 0;
    for (int p = 0; p < G.nranks; p++) max_count = max(max_count, counts[p]);
    const int N = G.N, NB = N >> 3, n1 = N >> 1, n2 = N >> 2, n3 = N >> 3;
    // both roles walk the same item sequence: j = blockIdx.x, + gridDim.x, ...; for each j the peers in an order rotated by the
    // reader's rank (so the box's readers do not all start on the same peer); an item exists where j < that peer's count
    uint32_t n = 0;                                                  // sequence number of the next item
    if (warp == G2_CONSUMERS)
    {   // ---- producer
        if (lane == 0)
            for (uint32_t j = blockIdx.x; j < max_count; j += gridDim.x)
                for (int q = 1; q < G.nranks; q++)
                {
                    const int p = (G.rank + q) % G.nranks;
                    if (j >= counts[p]) continue;
                    const int slot = (int)(n % G2_SLOTS);
                    if (!mbar_wait_bounded(&empty[slot], ((n / G2_SLOTS) & 1u) ^ 1u, dev_state)) return;
                    const uint8_t* rec = reinterpret_cast<const uint8_t*>(G.peer_export[p]) + (size_t)j * 4096;
                    uint8_t* dst = ring + (size_t)slot * G2_SLOT_BYTES;
                    if (level0) { mbar_expect_tx(&full[slot], G2_REC0_BYTES); bulk_load(dst, rec, G2_REC0_BYTES, &full[slot]); }
                    else { mbar_expect_tx(&full[slot], G2_REC_BYTES); bulk_load(dst + 2048, rec + 2048, G2_REC_BYTES, &full[slot]); }
                    n++;
                }
        return;
    }
    // ---- consumers: warp w takes the items with n % G2_CONSUMERS == w
    for (uint32_t j = blockIdx.x; j < max_count; j += gridDim.x)
        for (int q = 1; q < G.nranks; q++)
        {
            const int p = (G.rank + q) % G.nranks;
            if (j >= counts[p]) continue;
            const uint32_t mine = n++;
            if ((int)(mine % G2_CONSUMERS) != warp) continue;
            const int slot = (int)(mine % G2_SLOTS);
            if (!mbar_wait_bounded(&full[slot], (mine / G2_SLOTS) & 1u, dev_state)) return;
            const uint32_t* rec = reinterpret_cast<const uint32_t*>(ring + (size_t)slot * G2_SLOT_BYTES);
            const uint4* rec4 = reinterpret_cast<const uint4*>(rec);
            const uint32_t b = rec[950] & 0x7fffffffu;
            const int bx = b % NB, by = (b / NB) % NB, bz = b / (NB * NB);
            if (level0)
            {
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    const int qq = lane + 32 * k, row = qq >> 1, half = qq & 1, y = row & 7, z = row >> 3;
                    const uint4 v = rec4[qq];
                    surf3Dwrite(v, G.rad_surf, (bx * 8 + half * 4) * 4, by * 8 + y, bz * 8 + z);
                    if (G.write_linear) *reinterpret_cast<uint4*>(G.rad_lin + ((size_t)(bz * 8 + z) * N + (by * 8 + y)) * N + bx * 8 + half * 4) = v;
                }
            }
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                const int jj = lane + 32 * k, d = jj >> 4, r = jj & 15, oy = r & 3, oz = r >> 2;
                const int gx = bx * 4, gy = by * 4 + oy, gz = bz * 4 + oz;
                const uint4 v = rec4[128 + jj];
                surf3Dwrite(v, G.surf[0], gx * 4, gy, atlas_z(d, n1, gz));
                if (G.write_linear) *reinterpret_cast<uint4*>(G.lin[0][d] + ((size_t)gz * n1 + gy) * n1 + gx) = v;
            }
            if (lane < 24)
            {
                const int d = lane >> 2, r = lane & 3, oy = r & 1, oz = r >> 1;
                const int gx = bx * 2, gy = by * 2 + oy, gz = bz * 2 + oz;
                const uint2 v = reinterpret_cast<const uint2*>(rec + 896)[lane];
                surf3Dwrite(v, G.surf[1], gx * 4, gy, atlas_z(d, n2, gz));
                if (G.write_linear) *reinterpret_cast<uint2*>(G.lin[1][d] + ((size_t)gz * n2 + gy) * n2 + gx) = v;
            }
            if (lane < 6)
            {
                const uint32_t v = rec[944 + lane];
                surf3Dwrite(v, G.surf[2], bx * 4, by, atlas_z(lane, n3, bz));
                G.lin[2][lane][((size_t)bz * n3 + by) * n3 + bx] = v;             // level 3 is the source of the local tail: always
            }
            __syncwarp();                                            // the whole warp has read the slot
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
}

}  // namespace

// own pointer of a shareable buffer (allocating it on first use)
int f184_ipc_buffer_ptr(f184_ctx* c, uint32_t buffer, void** out)
{
    const uint32_t N = c->cfg.grid_n, n_bricks = (N / 8) * (N / 8) * (N / 8);
    switch (buffer)
    {
    case F184_IPC_ACCUM_COLOR:
    case F184_IPC_ACCUM_NORMAL:
    case F184_IPC_BRICK_FLAGS:
    {
        const int slot = buffer == F184_IPC_ACCUM_COLOR ? F184_SLOT_ACCUM_COLOR : (buffer == F184_IPC_ACCUM_NORMAL ? F184_SLOT_ACCUM_NORMAL : F184_SLOT_BRICK_FLAGS);
        int rc = f184_ensure_image(c, slot);
        if (rc) return rc;
        if (!c->img[slot].owned) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "ipc: slot %d is bound to caller memory", slot);
        *out = c->img[slot].ptr;
        return F184_OK;
    }
    case F184_IPC_EXPORT:
        if (!c->export_buf)
        {
            const uint32_t own = n_bricks / (c->cfg.nranks ? c->cfg.nranks : 1);
            CK(c, cudaMalloc(&c->export_buf, 4096ull * own));
        }
        *out = c->export_buf;
        return F184_OK;
    case F184_IPC_COUNTERS: *out = c->counters_dev; return F184_OK;
    case F184_IPC_BRICK_LIST:
        if (!c->brick_list)
        {
            CK(c, cudaMalloc(&c->brick_prev, 4ull * n_bricks));
            CK(c, cudaMemsetAsync(c->brick_prev, 0, 4ull * n_bricks, c->stream));
            CK(c, cudaMalloc(&c->brick_list, 4ull * n_bricks));
        }
        *out = c->brick_list;
        return F184_OK;
    case F184_IPC_FRAG_QUEUE:
    case F184_IPC_FRAG_COUNTS:
        if (!c->frag_queue)
        {   // capacity per sender: generous (4 records per triangle of the scene, 1 Mi..8 Mi); a full region is not an error, the
            // sender falls back to remote reductions
            static const long env_cap = [] { const char* e = getenv("F184_FRAG_QUEUE_RECORDS"); return e ? atol(e) : 0l; }();
            uint64_t cap = env_cap > 0 ? (uint64_t)env_cap : std::min<uint64_t>(std::max<uint64_t>(4ull * c->n_tris, 1ull << 20), 8ull << 20);
            cap = (cap + F184_FRAG_SUBQUEUES - 1) / F184_FRAG_SUBQUEUES * F184_FRAG_SUBQUEUES;
            c->frag_cap = (uint32_t)cap;
            const uint32_t G = c->cfg.nranks ? c->cfg.nranks : 1;
            CK(c, cudaMalloc(&c->frag_queue, sizeof(uint4) * cap * G));
            CK(c, cudaMalloc(&c->frag_counts, 8 * F184_FRAG_SUBQUEUES * sizeof(uint32_t)));
            CK(c, cudaMalloc(&c->frag_cursor, 8 * F184_FRAG_SUBQUEUES * sizeof(uint32_t)));
            CK(c, cudaMemsetAsync(c->frag_counts, 0, 8 * F184_FRAG_SUBQUEUES * sizeof(uint32_t), c->stream));
            CK(c, cudaMemsetAsync(c->frag_cursor, 0, 8 * F184_FRAG_SUBQUEUES * sizeof(uint32_t), c->stream));
        }
        *out = buffer == F184_IPC_FRAG_QUEUE ? (void*)c->frag_queue : (void*)c->frag_counts;
        return F184_OK;
    case F184_IPC_SYNC:
        if (!c->sync_flags)
        {
            CK(c, cudaMalloc(&c->sync_flags, 16 * sizeof(uint32_t)));
            CK(c, cudaMemsetAsync(c->sync_flags, 0, 16 * sizeof(uint32_t), c->stream));
        }
        *out = c->sync_flags;
        return F184_OK;
    }
    return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "ipc: unknown buffer %u", buffer);
}

static int peer_ptr(f184_ctx* c, uint32_t p, uint32_t buffer, void** out)
{
    if (p == c->cfg.rank) return f184_ipc_buffer_ptr(c, buffer, out);
    if (!c->peer[p].buf[buffer]) return f184_fail(c, F184_ERR_NOT_READY, "rank %u: buffer %u of rank %u was not imported (f184_ipc_import)", c->cfg.rank, buffer, p);
    *out = c->peer[p].buf[buffer];
    return F184_OK;
}

extern "C" int f184_peer_barrier(f184_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    if (c->cfg.nranks <= 1) return F184_OK;
    CK(c, cudaSetDevice(c->cfg.device));
    BarrierArgs B{};
    for (uint32_t p = 0; p < c->cfg.nranks; p++)
    {
        void* q = nullptr;
        int rc = peer_ptr(c, p, F184_IPC_SYNC, &q);
        if (rc) return rc;
        B.flags[p] = reinterpret_cast<uint32_t*>(q);
    }
    B.rank = (int)c->cfg.rank; B.nranks = (int)c->cfg.nranks;
    B.epoch = ++c->barrier_epoch;
    static const unsigned long long timeout_ms = [] { const char* e = getenv("F184_BARRIER_TIMEOUT_MS"); return e && atoll(e) > 0 ? (unsigned long long)atoll(e) : 5000ull; }();
    B.timeout_ns = timeout_ms * 1000000ull;
    B.dev_state = c->dev_state;
    F184Section sec;
    int rc = f184_enter(c, F184_SID_BUILD, &sec);      // the barriers belong to the build stream's chain (DESIGN.md "Frame pipeline")
    if (rc) return rc;
    auto body = [&]() -> int {
        int rc = f184_build_wait_vox(c);               // this rank's fragments (possibly still in flight on vox_stream) leave before its flag does
        if (rc) return rc;
        if (c->cur_sid == F184_SID_PASS && (rc = f184_join_internal(c))) return rc;
        rc = f184_stage_begin(c, F184_STAGE_BARRIER);
        if (rc) return rc;
        k_peer_barrier<<<1, 32, 0, c->stream>>>(B);
        CK_LAUNCH(c);
        rc = f184_stage_end(c, F184_STAGE_BARRIER);
        if (rc) return rc;
        if (!c->ev_barrier) CK(c, cudaEventCreateWithFlags(&c->ev_barrier, cudaEventDisableTiming));
        CK(c, cudaEventRecord(c->ev_barrier, c->stream));
        c->barrier_recorded = true;
        c->frag_sent_applied = true;
        return F184_OK;
    };
    return f184_leave(c, sec, body());
}

int f184_gather_n(f184_ctx* c)
{
    if (c->cfg.nranks <= 1) return F184_OK;
    int rc = f184_mode_n_alloc(c); if (rc) return rc;
    rc = f184_ensure_image(c, F184_SLOT_RADIANCE); if (rc) return rc;
    rc = f184_ensure_image(c, F184_SLOT_MIPS); if (rc) return rc;
    for (int s_ : {F184_SLOT_DEPTH, F184_SLOT_MATERIAL}) { rc = f184_ensure_image(c, s_); if (rc) return rc; }
    rc = f184_volume_begin_write(c); if (rc) return rc;
    VolumeSet& vs = c->vs[c->build_set];
    const int set = c->build_set;
    GatherArgs G{};
    for (uint32_t p = 0; p < c->cfg.nranks; p++)
    {
        void *e = nullptr, *k = nullptr, *l = nullptr;
        if ((rc = peer_ptr(c, p, F184_IPC_EXPORT, &e)) || (rc = peer_ptr(c, p, F184_IPC_COUNTERS, &k)) || (rc = peer_ptr(c, p, F184_IPC_BRICK_LIST, &l))) return rc;
        G.peer_export[p] = reinterpret_cast<const uint32_t*>(e);
        G.peer_counters[p] = reinterpret_cast<const unsigned long long*>(k);
        G.peer_list[p] = reinterpret_cast<const uint32_t*>(l);
    }
    G.rank = (int)c->cfg.rank; G.nranks = (int)c->cfg.nranks; G.N = (int)c->cfg.grid_n;
    G.write_linear = (c->cfg.flags & F184_FLAG_GATHER_LINEAR) ? 1 : 0;
    G.dev_state = c->dev_state;
    G.rad_surf = vs.rad_surf;
    G.rad_lin = img_ptr<uint32_t>(c, F184_SLOT_RADIANCE);
    uint32_t* mips = img_ptr<uint32_t>(c, F184_SLOT_MIPS);
    for (int l = 0; l < 3; l++)
    {
        G.surf[l] = vs.dir_surf[l];
        for (int d = 0; d < 6; d++)
        {
            const uint64_t n = c->mip_levels[l].n;
            G.lin[l][d] = mips + c->mip_levels[l].offset_texels + (uint64_t)d * n * n * n;
        }
    }
    rc = f184_stage_begin(c, F184_STAGE_EXCHANGE);
    if (rc) return rc;
    {   // does level 0 have to travel this frame?  (F184_FLAG_GATHER_LINEAR — the tests' full comparison — and F184_GATHER_LEVEL0=1 force it)
        static const bool force_env = [] { const char* e = getenv("F184_GATHER_LEVEL0"); return e && atoi(e) != 0; }();
        const bool force = force_env || (c->cfg.flags & F184_FLAG_GATHER_LINEAR);
        if ((rc = f184_fill_async(c, c->dev_state + F184_DEV_NEED_L0, force ? 1u : 0u, 4, c->stream))) return rc;
        if (!force)
        {
            NeedArgs A{};
            A.depth = img_ptr<float>(c, F184_SLOT_DEPTH);
            A.material = img_ptr<uchar4>(c, F184_SLOT_MATERIAL);
            A.W = c->cfg.width;
            A.n_tiles = f184_trace_tiles(c, c->cfg.height, &A.y0, &A.y1, &A.tile0, &A.tile_stride);
            A.h = c->voxel_h;
            A.spec_b = (c->cfg.flags & F184_FLAG_SPEC_APPENDIX_B) ? 1u : 0u;
            A.dev_state = c->dev_state;
            k_need_level0<<<148, 256, 0, c->stream>>>(A);
            CK_LAUNCH(c);
        }
        k_clear_foreign_level0<<<148 * 4, 256, 0, c->stream>>>(vs.rad_surf, (int)c->cfg.grid_n, c->cfg.nranks, c->cfg.rank, c->dev_state, set);
        CK_LAUNCH(c);
    }
    static const bool gather_ldg = [] { const char* e = getenv("F184_GATHER_LDG"); return e && atoi(e) != 0; }();
    if (gather_ldg) k_gather_bricks<<<dim3(c->cfg.nranks, 148), GATHER_WARPS * 32, 0, c->stream>>>(G);
    else
    {
        static bool attr = false;
        if (!attr)
        {
            CK(c, cudaFuncSetAttribute(k_gather_bricks_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SLOTS * G2_SLOT_BYTES));
            attr = true;
        }
        k_gather_bricks_tma<<<148 * 2, (G2_CONSUMERS + 1) * 32, G2_SLOTS * G2_SLOT_BYTES, c->stream>>>(G, c->dev_state);
    }
    CK_LAUNCH(c);
    // bytes per record that cross NVLink: the bulk copies of the TMA-fed kernel, or list entry + the record parts the per-lane loads read
    k_gather_state<<<1, 1, 0, c->stream>>>(G, c->dev_state, set, c->counters_dev, gather_ldg ? 4 + 2048 + 4 * (384 + 48 + 6) : G2_REC0_BYTES,
                                           gather_ldg ? 4 + 4 * (384 + 48 + 6) : G2_REC_BYTES);
    CK_LAUNCH(c);
    rc = f184_stage_end(c, F184_STAGE_EXCHANGE);
    if (rc) return rc;
    rc = f184_mips_tail_n(c, true);
    if (rc) return rc;
    return f184_volume_publish(c);        // every rank's bricks are in: this set is what the next trace samples
}
