// mode_n_shard.cu — one NVLink box, one process per GPU: the pieces of the north-star schedule that touch peer memory
// (include/f184.h "one NVLink box"; DESIGN.md "Multi-GPU").  The reference is single-GPU and single-queue
// (RHI/Private/Vulkan/DeviceVk.cpp:301-304); nothing here has a counterpart in it.
//
//   * the reduce-scatter of the partial volumes is an all-to-all of FRAGMENTS (mode_n_voxelize.cu): a fragment of a brick another
//     rank owns — bricks are owned along diagonals, (bx + by + bz) % G — becomes a 16-byte record in a queue in the sender's own
//     memory; behind one barrier the owner pulls its queues over NVLink and reduces locally.  The queues and their cursors are
//     allocated here (f184_ipc_buffer_ptr);
//   * k_peer_barrier: device-side flag barrier — each rank stores its epoch into every peer's flag array
//     (st.release.sys over NVLink) and spins on its own array (ld.acquire.sys, local memory): no host round trip and
//     no NCCL launch on the frame's critical path; a timeout (5 s; F184_BARRIER_TIMEOUT_MS) turns a missing peer into an
//     error instead of a hang: the kernel sets a sticky device error bit and the next synchronous call of the context
//     (f184_sync, f184_counter_get, f184_stage_time_*) returns F184_ERR_PEER_TIMEOUT;
//   * k_need_bricks: what do the cones of THIS rank's rows sample?  Level 0 at all (glossy pixels only), and level 1 of which
//     bricks (a conservative box around every pixel's world position; the specular cone marched where it stays fine longer);
//   * k_gather_bricks_tma: the all-gather before tracing — every rank pulls the other ranks' finished bricks out of the export
//     arrays mode_n_mips.cu writes (level 0 | level 1 | "coarse": levels 2, 3 + brick index) with bulk copies (cp.async.bulk
//     into shared memory, mbarrier completion): one 16 KB copy per 64 bricks for the coarse blocks, level 1 / level 0 per marked
//     brick, and stores them through surfaces into its own texture set.  Levels >= 4 are then finished locally (k_mips_tail).
//     C4 at 8 GPUs: 73 MB per rank instead of 1.1 GB, 0.19 ms.
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

// What does the cone trace of this rank's rows need from the OTHER ranks?  Replays the tracer's per-pixel set-up and level selection
// (mode_n_trace.cu; same expressions) without touching the volume:
//   * level 0 at all?  (dev_state[F184_DEV_NEED_L0]: some cone's first — finest — sample selects it.)  With 60-degree diffuse cones and
//     rough materials none does, and level 0 (2 KB of every 4 KB record) stays home;
//   * level 1 of WHICH bricks?  A cone reads level 1 only while its footprint is under ~3 voxels: the first sample of a diffuse
//     cone, 2.3 voxels above the surface the pixel shows.  So level 1 is needed only around the surfaces this rank's pixels see —
//     a fraction of the scene (the interiors of seven of C4's eight buildings are invisible from the camera).  The mask may be a
//     SUPERSET of what the cones read (a brick too many costs 1.5 KB of NVLink; a brick too few is a wrong image), so the six diffuse
//     cones are not marched: their fine samples lie within (normal offset h) + (the diffuse cone's fine reach) of the pixel's world
//     position, and every brick under that box — plus the trilinear footprint and a quarter texel of slack — is marked.  The
//     specular cone is marched sample by sample, and only where its fine reach exceeds the diffuse one (glossy materials).
//     (Marching all seven cones was ~2600 instructions per pixel: 0.3 ms per frame at 2 GPUs, 0.14 at 8 — a third of the trace itself.)
// No early termination on opacity here (it would need the volume): the marching stops where the tracer's level selection passes 1.
struct NeedArgs
{
    M4 InvProj, InvModelView, w2v;
    const float* depth;
    const uint16_t* normals;
    const uchar4* material;
    uint32_t W, H, y0, y1, tile0, tile_stride, n_tiles;
    float h, max_dist;
    f3 cam;
    uint32_t spec_b, N, no_view;       // no_view: no camera given — only the level-0 question is answered (it needs the roughness alone)
    float texel_per_world[3];          // level-1 texels per world unit along each volume axis (row norms of w2v)
    uint32_t* dev_state;
    uint32_t* need1;                   // bit per brick: level 1 of the brick is sampled by a cone of this rank's rows
};

// A pixel's seven cones take their few fine samples inside the same one to three bricks, and so do its neighbours — and a word of
// the mask is 32 bricks along x, half the width of a 512^3 volume: a wall along x sends every pixel that shows it to the same few
// words.  Reductions on one address are serialised by the L2 slice that owns it (the kernel took 0.14-0.2 ms per frame, slower the
// COARSER the volume).  So: the thread remembers the last two bricks it marked; new ones go into a small per-CTA hash set in shared
// memory (16 x 8 pixels see a handful of bricks); at the end the CTA sends one reduction per brick of the set whose bit is not set yet.
constexpr int MARK_SET = 256;
struct MarkCache
{
    uint32_t last0 = 0xffffffffu, last1 = 0xffffffffu;
    uint32_t* set;                        // shared memory, MARK_SET entries, 0xffffffff = free
};
__device__ __forceinline__ void mark_brick(uint32_t* need1, int bx, int by, int bz, int NB, MarkCache& mc)
{
    if ((unsigned)bx >= (unsigned)NB || (unsigned)by >= (unsigned)NB || (unsigned)bz >= (unsigned)NB) return;
    const uint32_t b = ((uint32_t)bz * NB + by) * NB + bx;
    if (b == mc.last0 || b == mc.last1) return;
    mc.last1 = mc.last0; mc.last0 = b;
    uint32_t hsh = (b * 2654435761u) >> 24;
    for (int probe = 0; probe < 4; probe++, hsh = (hsh + 1u) & (MARK_SET - 1))
    {
        uint32_t cur = reinterpret_cast<volatile uint32_t*>(mc.set)[hsh];
        if (cur == b) return;
        if (cur == 0xffffffffu)
        {
            cur = atomicCAS(mc.set + hsh, 0xffffffffu, b);
            if (cur == 0xffffffffu || cur == b) return;
        }
    }
    atomicOr(need1 + (b >> 5), 1u << (b & 31u));                               // four occupied places in a row: straight to the mask
}

// how far along a cone its last sample at level <= 1 lies (0: none) — the loop of mark_cone without the marking, same arithmetic
__device__ float cone_fine_reach(const NeedArgs& A, float tan_half)
{
    const float h = A.h, inv_h = 1.0f / h;
    float t = 2.0f * h, last = 0.0f;
    while (t < A.max_dist)
    {
        float diam;
        const float lod = cone_lod(t, tan_half, h, inv_h, &diam);
        if (lod >= (A.spec_b ? 2.0f : 1.5f) + 1e-3f) break;
        last = t;
        t += A.spec_b ? 0.5f * diam : diam;
    }
    return last;
}

// one cone: every sample it takes while its level is <= 1; returns whether it touches level 0
__device__ bool mark_cone(const NeedArgs& A, f3 origin, f3 dir, float tan_half, MarkCache& mc)
{
    const float h = A.h, inv_h = 1.0f / h;
    const f3 dv = {(A.w2v.m[0] * dir.x + A.w2v.m[4] * dir.y + A.w2v.m[8] * dir.z) * 0.5f,
                   (A.w2v.m[1] * dir.x + A.w2v.m[5] * dir.y + A.w2v.m[9] * dir.z) * 0.5f,
                   (A.w2v.m[2] * dir.x + A.w2v.m[6] * dir.y + A.w2v.m[10] * dir.z)};
    const f3 o3 = mul43(A.w2v, origin, 1.0f);
    const f3 q0 = {o3.x * 0.5f + 0.5f, o3.y * 0.5f + 0.5f, o3.z};
    const float n1 = (float)(A.N >> 1);
    const int NB = (int)(A.N >> 3);
    bool level0 = false;
    float t = 2.0f * h;
    while (t < A.max_dist)
    {
        float diam;
        const float lod = cone_lod(t, tan_half, h, inv_h, &diam);
        // level 1 (atlas level 0) is read while: nearest spec L = floor(lod + .5) <= 1; Appendix-B spec: the mip-linear fetch at lod - 1 < 1
        if (lod >= (A.spec_b ? 2.0f : 1.5f) + 1e-3f) break;
        const float qx = q0.x + dv.x * t, qy = q0.y + dv.y * t, qz = q0.z + dv.z * t;
        if (!(qx >= 0.0f && qx <= 1.0f && qy >= 0.0f && qy <= 1.0f && qz >= 0.0f && qz <= 1.0f)) break;
        if (lod < (A.spec_b ? 1.0f : 0.5f) + 1e-3f) level0 = true;
        // bricks under the trilinear footprint in level-1 texel space (a brick = 4 level-1 texels per axis), with a margin
        const float px = qx * n1 - 0.5f, py = qy * n1 - 0.5f, pz = qz * n1 - 0.5f, e = 1.0f / 16.0f;
        const int x0 = (int)floorf(px - e) >> 2, x1 = ((int)floorf(px + e) + 1) >> 2;
        const int y0 = (int)floorf(py - e) >> 2, y1 = ((int)floorf(py + e) + 1) >> 2;
        const int z0 = (int)floorf(pz - e) >> 2, z1 = ((int)floorf(pz + e) + 1) >> 2;
        for (int bz = z0; bz <= z1; bz++)
            for (int by = y0; by <= y1; by++)
                for (int bx = x0; bx <= x1; bx++) mark_brick(A.need1, bx, by, bz, NB, mc);
        t += A.spec_b ? 0.5f * diam : diam;
    }
    return level0;
}

__global__ void __launch_bounds__(128) k_need_bricks(const NeedArgs A)
{
    // pixel mapping of the tracer's CTA, 128 threads = 16 x 8 pixels of one 8-row tile row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t y = A.y0 + (A.tile0 + blockIdx.y * A.tile_stride) * 8 + (warp >> 1) * 4 + (lane >> 3);
    bool level0 = false;
    __shared__ uint32_t mark_set[MARK_SET];
    for (int i = threadIdx.x; i < MARK_SET; i += blockDim.x) mark_set[i] = 0xffffffffu;
    __syncthreads();
    if (x < A.W && y < A.y1)
    {
        const float depth = __ldg(A.depth + (size_t)y * A.W + x);
        if (depth < 1.0f && A.no_view)
            level0 = cone_samples_level0(kTanHalfDiffuse, A.h, A.spec_b != 0) ||
                     cone_samples_level0(cone_specular_tan((float)__ldg(A.material + (size_t)y * A.W + x).y / 255.0f), A.h, A.spec_b != 0);
        else if (depth < 1.0f)
        {   // -- the tracer's set-up (k_trace_n), expression for expression
            const float uvx = ((float)x + 0.5f) / (float)A.W, uvy = ((float)y + 0.5f) / (float)A.H;
            const f4 cp = mul44(A.InvProj, f4{uvx * 2.0f - 1.0f, uvy * 2.0f - 1.0f, depth, 1.0f});
            const f3 cspos = {cp.x / cp.w, cp.y / cp.w, cp.z / cp.w};
            const f3 wpos = mul43(A.InvModelView, cspos, 1.0f);
            const float tan_s = cone_specular_tan((float)__ldg(A.material + (size_t)y * A.W + x).y / 255.0f);
            level0 = cone_samples_level0(kTanHalfDiffuse, A.h, A.spec_b != 0) || cone_samples_level0(tan_s, A.h, A.spec_b != 0);
            MarkCache mc;
            mc.set = mark_set;
            const float reach_d = cone_fine_reach(A, kTanHalfDiffuse);
            {   // the diffuse cones: the box around the pixel's position that holds all their fine samples
                const f3 o3 = mul43(A.w2v, wpos, 1.0f);
                const float n1 = (float)(A.N >> 1);
                const float px = (o3.x * 0.5f + 0.5f) * n1 - 0.5f, py = (o3.y * 0.5f + 0.5f) * n1 - 0.5f, pz = o3.z * n1 - 0.5f;
                const float r = (A.h + reach_d) * 1.0001f, e = 1.0f / 16.0f + 0.25f;
                const float rx = r * A.texel_per_world[0] + e, ry = r * A.texel_per_world[1] + e, rz = r * A.texel_per_world[2] + e;
                const int NB = (int)(A.N >> 3);
                const int x0 = max((int)floorf(px - rx) >> 2, 0), x1 = min(((int)floorf(px + rx) + 1) >> 2, NB - 1);
                const int y0 = max((int)floorf(py - ry) >> 2, 0), y1 = min(((int)floorf(py + ry) + 1) >> 2, NB - 1);
                const int z0 = max((int)floorf(pz - rz) >> 2, 0), z1 = min(((int)floorf(pz + rz) + 1) >> 2, NB - 1);
                // neighbouring pixels mostly land on the same box of bricks: one lane per distinct box of the warp walks it
                const bool small = x1 - x0 <= 3 && y1 - y0 <= 3 && z1 - z0 <= 3 && x0 <= x1 && y0 <= y1 && z0 <= z1;
                const uint32_t key = small ? ((uint32_t)x0 | ((uint32_t)y0 << 7) | ((uint32_t)z0 << 14) | ((uint32_t)(x1 - x0) << 21) | ((uint32_t)(y1 - y0) << 23) | ((uint32_t)(z1 - z0) << 25))
                                           : (0x80000000u | (uint32_t)lane);
                const uint32_t same = __match_any_sync(__activemask(), key);
                if (reach_d > 0.0f && (__ffs(same) - 1) == lane)
                    for (int bz = z0; bz <= z1; bz++)
                        for (int by = y0; by <= y1; by++)
                            for (int bx = x0; bx <= x1; bx++) mark_brick(A.need1, bx, by, bz, NB, mc);
            }
            if (cone_fine_reach(A, tan_s) > reach_d)
            {   // a glossy pixel: its specular cone stays fine beyond the box
                const ushort4 nq = __ldg(reinterpret_cast<const ushort4*>(A.normals) + (size_t)y * A.W + x);
                const f3 csnorm = normalize3(f3{fmaf((float)nq.x / 65535.0f, 2.0f, -1.0f), fmaf((float)nq.y / 65535.0f, 2.0f, -1.0f), fmaf((float)nq.z / 65535.0f, 2.0f, -1.0f)});
                const f3 z = normalize3(mul33(A.InvModelView, csnorm));
                const f3 origin = {wpos.x + z.x * A.h, wpos.y + z.y * A.h, wpos.z + z.z * A.h};
                const f3 I = normalize3(wpos - A.cam);
                const float ndi = dot3(z, I);
                const f3 R = {I.x - 2.0f * ndi * z.x, I.y - 2.0f * ndi * z.y, I.z - 2.0f * ndi * z.z};
                if (dot3(R, z) > -1e-3f) mark_cone(A, origin, R, tan_s, mc);
            }
        }
    }
    if (__syncthreads_or(level0) && threadIdx.x == 0) A.dev_state[F184_DEV_NEED_L0] = 1u;
    // (the barrier above also orders the insertions before these reads)
    for (int i = threadIdx.x; i < MARK_SET; i += blockDim.x)
    {
        const uint32_t b = mark_set[i];
        if (b != 0xffffffffu && !((__ldcg(A.need1 + (b >> 5)) >> (b & 31u)) & 1u)) atomicOr(A.need1 + (b >> 5), 1u << (b & 31u));
    }
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
    const uint32_t* peer_export[8];    // each rank's export arrays (k_mips_bricks): level 0 | fine (level 1) | coarse (levels 2, 3, brick index)
    const unsigned long long* peer_counters[8];
    const uint32_t* dev_state;
    unsigned long long* gather_bytes;  // F184_COUNTER_GATHER_BYTES: what the bulk copies move over NVLink (summed by the producers)
    uint32_t export_cap;               // bricks per export array
    int rank, nranks, N, write_linear;
    int dry;                           // measurement aid (F184_GATHER_DRY): 1 = skip the level 2/3 stores, 2 = skip the level 0/1 stores
    cudaSurfaceObject_t rad_surf;
    uint32_t* rad_lin;
    uint32_t* lin[3][6];
    cudaSurfaceObject_t surf[3];       // atlas levels 1..3 (direction d at z + 2 d n)
};

// behind the gather: the set's level-0 bookkeeping
__global__ void k_gather_state(uint32_t* dev_state, int set)
{
    dev_state[F184_DEV_L0_FULL + set] = dev_state[F184_DEV_NEED_L0] != 0 ? 1u : 0u;
}

// ---- the gather ------------------------------------------------------------------------------------------------------------------
// Nothing but bulk copies (cp.async.bulk: the TMA engine moves peer HBM -> shared memory over NVLink, completion on an mbarrier)
// crosses NVLink, issued by one producer warp per CTA, so the bytes in flight do not depend on how many warps of the SM the
// gather gets while the cone trace of the previous frame runs beside it.  (Per-lane peer loads ran at 70 GB/s inside the frame;
// one 1.7 KB record per brick in one copy ran at 460 GB/s where level 1 was needed, but the 224-byte copies of the bricks whose
// level 1 no cone needs — most of them, at 8 GPUs — were bound by copies in flight, not bytes: 86 GB/s.)  So the owner exports
// three arrays, and what every rank needs of EVERY brick — levels 2 and 3 and the brick's index, 256 bytes — is contiguous over
// consecutive bricks: one 16 KB copy moves a work unit's 64 bricks.  Level 1 (1536 B) and, in a glossy scene, level 0 (2048 B)
// follow per brick through a ring, only for the bricks k_need_bricks marked.
constexpr int G4_CONSUMERS = 8;                  // warps; + 1 producer warp
constexpr int G4_SLOTS = 16;                     // ring of per-brick fine copies
// A slot must always be consumed by the SAME warp (fine record n goes to warp n % G4_CONSUMERS and to slot n % G4_SLOTS): a parity wait
// only tells "one phase ago" from "now", so a warp that ran a whole lap ahead of the slot's previous consumer would see the phase
// it waits for as already complete and read the old record.  (Seven consumers did exactly that.)
static_assert(G4_SLOTS % G4_CONSUMERS == 0, "every ring slot is owned by one consumer warp");
constexpr int G4_SLOT_BYTES = 3584;              // level 0 at [0, 2048), level 1 at [2048, 3584)
constexpr int G4_UNIT = 64;                      // bricks per work unit
constexpr int G4_COARSE_BUFS = 3;                // unit k + 1 is prefetched while unit k is worked on and the consumers may still be on k - 1
constexpr int G4_SMEM_BYTES = G4_COARSE_BUFS * G4_UNIT * 256 + G4_SLOTS * G4_SLOT_BYTES + G4_COARSE_BUFS * G4_UNIT * 4;

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

// Work = units of G4_UNIT consecutive bricks of one peer's list, walked in the same order by both roles: u = 0, 1, ...; for each u the
// peers in an order rotated by the reader's rank (so the box's readers do not all start on the same peer); CTA c takes every
// gridDim.x-th unit number.  Per unit the producer fetches the coarse blocks, looks every brick up in the need mask (k_need_bricks) and in
// this rank's "my copy of its level 1 is not zero" bits, publishes per brick
//     info = level 1 travels | (my copy is non-zero) << 1 | (sequence number among this CTA's fine copies) << 2
// and issues the fine copies.  The consumers write what arrived; a brick whose level 1 did not travel gets zeros there IF this
// rank's copy of it is not zero already — so a foreign brick's level 1 on this rank is always either current or zero, never stale.
__global__ void __launch_bounds__((G4_CONSUMERS + 1) * 32) k_gather_bricks_tma(const GatherArgs G, uint32_t* dev_state, const uint32_t* __restrict__ need1,
                                                                               uint32_t* __restrict__ l1_nonzero)
{
    extern __shared__ __align__(128) uint8_t gsm[];
    __shared__ __align__(8) uint64_t full[G4_SLOTS], empty[G4_SLOTS], coarse_full[G4_COARSE_BUFS], coarse_free[G4_COARSE_BUFS], info_ready[G4_COARSE_BUFS];
    __shared__ uint32_t counts[8];
    uint32_t* coarse = reinterpret_cast<uint32_t*>(gsm);                                            // [buf][brick][64]
    uint8_t* ring = gsm + G4_COARSE_BUFS * G4_UNIT * 256;
    uint32_t* info = reinterpret_cast<uint32_t*>(ring + G4_SLOTS * G4_SLOT_BYTES);                  // [buf][brick]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
    {
        for (int s_ = 0; s_ < G4_SLOTS; s_++) { mbar_init(&full[s_], 1); mbar_init(&empty[s_], 1); }
        for (int s_ = 0; s_ < G4_COARSE_BUFS; s_++) { mbar_init(&coarse_full[s_], 1); mbar_init(&coarse_free[s_], G4_CONSUMERS); mbar_init(&info_ready[s_], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 8) counts[threadIdx.x] = ((int)threadIdx.x < G.nranks && (int)threadIdx.x != G.rank) ? (uint32_t)G.peer_counters[threadIdx.x][F184_COUNTER_COUNT] : 0u;
    __syncthreads();
    const bool level0 = G.dev_state[F184_DEV_NEED_L0] != 0;
    uint32_t max_units = 0;
    for (int p = 0; p < G.nranks; p++) max_units = max(max_units, (counts[p] + G4_UNIT - 1) / G4_UNIT);
    const int N = G.N, NB = N >> 3, n1 = N >> 1, n2 = N >> 2, n3 = N >> 3;
    // the walk: unit number g = u * (nranks - 1) + (q - 1) names unit u of the q-th peer after this rank; CTA c takes g = c, c + gridDim.x, ...
    // and skips the numbers past the end of a peer's list.  (Cost per CTA ~ its own units.  Walking ALL units in every warp and counting
    // the valid ones — two runtime divisions each — was 0.3-0.6 ms of serial latency at 1024^3: most of what the exchange took.)
    uint32_t g_next = blockIdx.x;
    const uint32_t g_end = max_units * (uint32_t)(G.nranks - 1);
    auto next_unit = [&](int& p_out, uint32_t& u_out) -> bool {
        for (; g_next < g_end; g_next += gridDim.x)
        {
            const uint32_t uu = g_next / (uint32_t)(G.nranks - 1), qq = 1u + g_next % (uint32_t)(G.nranks - 1);
            const int p = (int)((G.rank + qq) % (uint32_t)G.nranks);
            if (uu * G4_UNIT >= counts[p]) continue;
            p_out = p; u_out = uu; g_next += gridDim.x;
            return true;
        }
        return false;
    };
    uint32_t k = 0;                                                  // this CTA's units so far (coarse buffer = k % G4_COARSE_BUFS)
    if (warp == G4_CONSUMERS)
    {   // ---- producer
        unsigned long long sent = 0;
        uint32_t fine_n = 0;                                         // fine copies of this CTA so far
        bool fail = false;
        auto fetch_coarse = [&](int p, uint32_t uu, uint32_t kk) {   // lane 0: the coarse blocks of unit (p, uu) into buffer kk % G4_COARSE_BUFS
            const uint32_t buf = kk % G4_COARSE_BUFS, use = kk / G4_COARSE_BUFS;
            if (use && !mbar_wait_bounded(&coarse_free[buf], (use - 1u) & 1u, dev_state)) { fail = true; return; }
            const uint32_t first = uu * G4_UNIT, cnt = min((uint32_t)G4_UNIT, counts[p] - first);
            mbar_expect_tx(&coarse_full[buf], cnt * 256u);
            bulk_load(coarse + (size_t)buf * G4_UNIT * 64, G.peer_export[p] + (size_t)G.export_cap * 896 + (size_t)first * 64, cnt * 256u, &coarse_full[buf]);
            sent += cnt * 256u;
        };
        int p_cur = 0, p_nxt = 0; uint32_t u_cur = 0, u_nxt = 0;
        bool have = next_unit(p_cur, u_cur);
        if (have && lane == 0) fetch_coarse(p_cur, u_cur, 0);
        while (have)
        {
            const bool have_nxt = next_unit(p_nxt, u_nxt);
            if (have_nxt && lane == 0) fetch_coarse(p_nxt, u_nxt, k + 1);
            const uint32_t buf = k % G4_COARSE_BUFS, use = k / G4_COARSE_BUFS;
            {   // every lane waits for the coarse blocks itself (the wait is what makes the bulk copy's bytes visible to the waiting thread)
                const bool ok = !fail && mbar_wait_bounded(&coarse_full[buf], use & 1u, dev_state);
                if (!__all_sync(0xffffffffu, ok)) return;
            }
            const uint32_t* cb = coarse + (size_t)buf * G4_UNIT * 64;
            const uint32_t first = u_cur * G4_UNIT, cnt = min((uint32_t)G4_UNIT, counts[p_cur] - first);
            bool my_l1[G4_UNIT / 32]; uint32_t my_n[G4_UNIT / 32];
#pragma unroll
            for (int bi = 0; bi < G4_UNIT / 32; bi++)
            {   // the two bits the consumers need about each brick, looked up for 32 bricks in one round of loads
                const uint32_t j = bi * 32 + lane;
                bool l1 = false, nz = false;
                if (j < cnt)
                {
                    const uint32_t b = cb[j * 64 + 54] & 0x7fffffffu;
                    l1 = level0 || ((__ldg(need1 + (b >> 5)) >> (b & 31u)) & 1u);
                    nz = (l1_nonzero[b >> 5] >> (b & 31u)) & 1u;
                }
                const uint32_t ball = __ballot_sync(0xffffffffu, l1);
                my_l1[bi] = l1;
                my_n[bi] = fine_n + __popc(ball & ((1u << lane) - 1u));
                fine_n += __popc(ball);
                if (j < cnt) info[buf * G4_UNIT + j] = (l1 ? 1u : 0u) | (nz ? 2u : 0u) | (my_n[bi] << 2);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&info_ready[buf]);            // (before the copies: a ring slot this unit waits for may hold a record of this unit)
#pragma unroll
            for (int bi = 0; bi < G4_UNIT / 32; bi++)
            {
                const uint32_t base_n = __shfl_sync(0xffffffffu, my_n[bi], 0);
                for (uint32_t round = 0; round < 32 / G4_SLOTS; round++)
                {   // G4_SLOTS lanes at a time: their slots are distinct, and every slot's previous record was issued in an earlier round
                    if (my_l1[bi] && (my_n[bi] - base_n) / G4_SLOTS == round)
                    {
                        const uint32_t n = my_n[bi], slot = n % G4_SLOTS;
                        if (!mbar_wait_bounded(&empty[slot], ((n / G4_SLOTS) & 1u) ^ 1u, dev_state)) fail = true;
                        else
                        {
                            const size_t i = (size_t)first + bi * 32 + lane;
                            uint8_t* dst = ring + (size_t)slot * G4_SLOT_BYTES;
                            const uint32_t bytes = level0 ? 3584u : 1536u;
                            mbar_expect_tx(&full[slot], bytes);
                            if (level0) bulk_load(dst, G.peer_export[p_cur] + i * 512, 2048u, &full[slot]);
                            bulk_load(dst + 2048, G.peer_export[p_cur] + (size_t)G.export_cap * 512 + i * 384, 1536u, &full[slot]);
                            sent += bytes;
                        }
                    }
                    if (__any_sync(0xffffffffu, fail)) return;
                }
            }
            have = have_nxt; p_cur = p_nxt; u_cur = u_nxt; k++;
        }
        for (int o = 16; o; o >>= 1) sent += __shfl_xor_sync(0xffffffffu, sent, o);
        if (lane == 0 && sent) atomicAdd(G.gather_bytes, sent);
        return;
    }
    // ---- consumers: a brick with a fine copy goes to the warp that owns the copy's slot, any other to warp j % G4_CONSUMERS
    int p_cur = 0; uint32_t u_cur = 0;
    while (next_unit(p_cur, u_cur))
    {
        const uint32_t buf = k % G4_COARSE_BUFS, use = k / G4_COARSE_BUFS;
        const uint32_t cnt = min((uint32_t)G4_UNIT, counts[p_cur] - u_cur * G4_UNIT);
        if (!mbar_wait_bounded(&info_ready[buf], use & 1u, dev_state)) return;
        if (!mbar_wait_bounded(&coarse_full[buf], use & 1u, dev_state)) return;
        for (uint32_t j = 0; j < cnt; j++)
        {
            const uint32_t inf = info[buf * G4_UNIT + j];
            const bool l1 = (inf & 1u) != 0, was_nonzero = (inf & 2u) != 0;
            const uint32_t fn = inf >> 2;
            if ((int)((l1 ? fn : j) % G4_CONSUMERS) != warp) continue;
            const uint32_t* rec = coarse + ((size_t)buf * G4_UNIT + j) * 64;
            const uint32_t b = rec[54] & 0x7fffffffu;
            const int bx = b % NB, by = (b / NB) % NB, bz = b / (NB * NB);
            const uint32_t word = b >> 5, bit = 1u << (b & 31u);
            const uint32_t slot = fn % G4_SLOTS;
            const uint4* fine4 = reinterpret_cast<const uint4*>(ring + (size_t)slot * G4_SLOT_BYTES);
            if (l1 && !mbar_wait_bounded(&full[slot], (fn / G4_SLOTS) & 1u, dev_state)) return;
            if (level0 && !(G.dry & 2))
            {
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
                {
                    const int qq = lane + 32 * kk, row = qq >> 1, half = qq & 1, y = row & 7, z = row >> 3;
                    const uint4 v = fine4[qq];
                    surf3Dwrite(v, G.rad_surf, (bx * 8 + half * 4) * 4, by * 8 + y, bz * 8 + z);
                    if (G.write_linear) *reinterpret_cast<uint4*>(G.rad_lin + ((size_t)(bz * 8 + z) * N + (by * 8 + y)) * N + bx * 8 + half * 4) = v;
                }
            }
            if ((l1 || was_nonzero) && !(G.dry & 2))
            {   // the brick's level 1, or zeros over a copy of it that no cone of this rank needs any more
#pragma unroll
                for (int kk = 0; kk < 3; kk++)
                {
                    const int jj = lane + 32 * kk, d = jj >> 4, r = jj & 15, oy = r & 3, oz = r >> 2;
                    const int gx = bx * 4, gy = by * 4 + oy, gz = bz * 4 + oz;
                    const uint4 v = l1 ? fine4[128 + jj] : make_uint4(0, 0, 0, 0);
                    surf3Dwrite(v, G.surf[0], gx * 4, gy, atlas_z(d, n1, gz));
                    if (G.write_linear) *reinterpret_cast<uint4*>(G.lin[0][d] + ((size_t)gz * n1 + gy) * n1 + gx) = v;
                }
                if (lane == 0 && l1 != was_nonzero)
                {
                    if (l1) atomicOr(l1_nonzero + word, bit); else atomicAnd(l1_nonzero + word, ~bit);
                }
            }
            if (lane < 24 && !(G.dry & 1))
            {
                const int d = lane >> 2, r = lane & 3, oy = r & 1, oz = r >> 1;
                const int gx = bx * 2, gy = by * 2 + oy, gz = bz * 2 + oz;
                const uint2 v = reinterpret_cast<const uint2*>(rec)[lane];
                surf3Dwrite(v, G.surf[1], gx * 4, gy, atlas_z(d, n2, gz));
                if (G.write_linear) *reinterpret_cast<uint2*>(G.lin[1][d] + ((size_t)gz * n2 + gy) * n2 + gx) = v;
            }
            if (lane < 6 && !(G.dry & 1))
            {
                const uint32_t v = rec[48 + lane];
                surf3Dwrite(v, G.surf[2], bx * 4, by, atlas_z(lane, n3, bz));
                G.lin[2][lane][((size_t)bz * n3 + by) * n3 + bx] = v;             // level 3 is the source of the local tail: always
            }
            if (l1)
            {
                __syncwarp();                                        // the whole warp has read the slot
                if (lane == 0) mbar_arrive(&empty[slot]);
            }
        }
        __syncwarp();                                                // the whole warp has read the unit's coarse blocks and info words
        if (lane == 0) mbar_arrive(&coarse_free[buf]);
        k++;
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
            // per destination: 16 records per triangle of this rank's share of the scene (Sponza at 512^3 makes 18 fragments per triangle,
            // of which (G - 1) / G leave the rank, spread over G - 1 destinations), 1 Mi at least
            const uint64_t G_ = c->cfg.nranks ? c->cfg.nranks : 1;
            uint64_t cap = env_cap > 0 ? (uint64_t)env_cap : std::max<uint64_t>(16ull * c->n_tris / G_, 1ull << 20);
            cap = (cap + F184_FRAG_SUBQUEUES - 1) / F184_FRAG_SUBQUEUES * F184_FRAG_SUBQUEUES;
            c->frag_cap = (uint32_t)cap;
            const uint32_t G = c->cfg.nranks ? c->cfg.nranks : 1;
            CK(c, cudaMalloc(&c->frag_queue, sizeof(uint4) * cap * G));
            CK(c, cudaMalloc(&c->frag_counts, 8 * F184_FRAG_SUBQUEUES * sizeof(uint32_t)));
            CK(c, cudaMalloc(&c->frag_cursor, 8 * F184_FRAG_SUBQUEUES * F184_FRAG_CURSOR_STRIDE * sizeof(uint32_t)));
            CK(c, cudaMemsetAsync(c->frag_counts, 0, 8 * F184_FRAG_SUBQUEUES * sizeof(uint32_t), c->stream));
            CK(c, cudaMemsetAsync(c->frag_cursor, 0, 8 * F184_FRAG_SUBQUEUES * F184_FRAG_CURSOR_STRIDE * sizeof(uint32_t), c->stream));
        }
        *out = buffer == F184_IPC_FRAG_QUEUE ? (void*)c->frag_queue : (void*)c->frag_cursor;
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

// first-use set-up of the gather (called from f184_prepare_frame, i.e. before anything of a frame is enqueued: no allocation in the
// middle of a frame, where a peer barrier may be waiting)
int f184_gather_init_n(f184_ctx* c)
{
    static bool attr = false;
    if (!attr)
    {
        CK(c, cudaFuncSetAttribute(k_gather_bricks_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, G4_SMEM_BYTES));
        attr = true;
    }
    if (c->need1) return F184_OK;
    const uint32_t n_words = ((uint32_t)(c->cfg.grid_n / 8) * (c->cfg.grid_n / 8) * (c->cfg.grid_n / 8) + 31) / 32;
    CK(c, cudaMalloc(&c->need1, 4ull * n_words));
    for (int i = 0; i < 2; i++)
    {
        CK(c, cudaMalloc(&c->l1_nonzero[i], 4ull * n_words));
        int rc = f184_fill_async(c, c->l1_nonzero[i], 0u, 4ull * n_words, c->stream);
        if (rc) return rc;
    }
    CK(c, cudaStreamSynchronize(c->stream));
    return F184_OK;
}

int f184_gather_n(f184_ctx* c, const f184_trace_constants* view)
{
    if (c->cfg.nranks <= 1) return F184_OK;
    int rc = f184_mode_n_alloc(c); if (rc) return rc;
    rc = f184_ensure_image(c, F184_SLOT_RADIANCE); if (rc) return rc;
    rc = f184_ensure_image(c, F184_SLOT_MIPS); if (rc) return rc;
    for (int s_ : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_MATERIAL}) { rc = f184_ensure_image(c, s_); if (rc) return rc; }
    rc = f184_volume_begin_write(c); if (rc) return rc;
    VolumeSet& vs = c->vs[c->build_set];
    const int set = c->build_set;
    GatherArgs G{};
    for (uint32_t p = 0; p < c->cfg.nranks; p++)
    {
        void *e = nullptr, *k = nullptr;
        if ((rc = peer_ptr(c, p, F184_IPC_EXPORT, &e)) || (rc = peer_ptr(c, p, F184_IPC_COUNTERS, &k))) return rc;
        G.peer_export[p] = reinterpret_cast<const uint32_t*>(e);
        G.peer_counters[p] = reinterpret_cast<const unsigned long long*>(k);
    }
    G.export_cap = ((uint32_t)(c->cfg.grid_n / 8) * (c->cfg.grid_n / 8) * (c->cfg.grid_n / 8)) / c->cfg.nranks;
    G.rank = (int)c->cfg.rank; G.nranks = (int)c->cfg.nranks; G.N = (int)c->cfg.grid_n;
    G.write_linear = (c->cfg.flags & F184_FLAG_GATHER_LINEAR) ? 1 : 0;
    static const int dry = [] { const char* e = getenv("F184_GATHER_DRY"); return e ? atoi(e) : 0; }();
    G.dry = dry;
    G.dev_state = c->dev_state;
    G.gather_bytes = c->counters_dev + F184_COUNTER_GATHER_BYTES;
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
    rc = f184_stage_begin(c, F184_STAGE_NEED);
    if (rc) return rc;
    const uint32_t n_words = ((uint32_t)(c->cfg.grid_n / 8) * (c->cfg.grid_n / 8) * (c->cfg.grid_n / 8) + 31) / 32;
    if ((rc = f184_gather_init_n(c))) return rc;
    {   // What has to travel this frame?  Level 0: only if a cone of this rank's rows samples it.  Level 1: only the bricks such a cone
        // samples (needs the camera: f184_gather_volume_view).  F184_FLAG_GATHER_LINEAR — the tests' full comparison — and
        // F184_GATHER_LEVEL0=1 / F184_GATHER_ALL=1 force everything.
        static const bool force_l0 = [] { const char* e = getenv("F184_GATHER_LEVEL0"); return e && atoi(e) != 0; }();
        static const bool force_all = [] { const char* e = getenv("F184_GATHER_ALL"); return e && atoi(e) != 0; }();
        const bool force = force_l0 || (c->cfg.flags & F184_FLAG_GATHER_LINEAR);
        const bool all_l1 = force || force_all || !view;
        if ((rc = f184_fill_async(c, c->dev_state + F184_DEV_NEED_L0, force ? 1u : 0u, 4, c->stream))) return rc;
        if ((rc = f184_fill_async(c, c->need1, all_l1 ? 0xffffffffu : 0u, 4ull * n_words, c->stream))) return rc;
        if (!force)
        {
            NeedArgs A{};
            A.depth = img_ptr<float>(c, F184_SLOT_DEPTH);
            A.normals = img_ptr<uint16_t>(c, F184_SLOT_NORMALS);
            A.material = img_ptr<uchar4>(c, F184_SLOT_MATERIAL);
            A.W = c->cfg.width; A.H = c->cfg.height;
            A.n_tiles = f184_trace_tiles(c, c->cfg.height, &A.y0, &A.y1, &A.tile0, &A.tile_stride);
            A.h = c->voxel_h;
            A.max_dist = c->cfg.cone_max_distance;
            A.spec_b = (c->cfg.flags & F184_FLAG_SPEC_APPENDIX_B) ? 1u : 0u;
            A.N = c->cfg.grid_n;
            A.no_view = all_l1 ? 1u : 0u;
            A.dev_state = c->dev_state;
            A.need1 = c->need1;
            if (!all_l1)
            {
                memcpy(A.InvProj.m, view->view.InvProj, 64);
                memcpy(A.InvModelView.m, view->ext.InvModelView, 64);
                M4 vp, vv;
                memcpy(vp.m, view->ext.VoxelProj, 64);
                memcpy(vv.m, view->ext.VoxelView, 64);
                A.w2v = host_matmul(vp, vv);
                for (int i = 0; i < 3; i++)
                {   // |d(texel_i)| per unit of world distance, at most: the norm of row i of the linear part (x and y are mapped [-1, 1] -> [0, 1])
                    const float* m = A.w2v.m;
                    A.texel_per_world[i] = sqrtf(m[i] * m[i] + m[4 + i] * m[4 + i] + m[8 + i] * m[8 + i]) * (i < 2 ? 0.5f : 1.0f) * (float)(c->cfg.grid_n >> 1);
                }
                A.h = f184_voxel_h(view->ext.VoxelProj, view->ext.VoxelView, c->cfg.grid_n);      // the tracer's own h
                A.cam = {A.InvModelView.m[12], A.InvModelView.m[13], A.InvModelView.m[14]};
            }
            if (A.n_tiles)
            {
                k_need_bricks<<<dim3((A.W + 15) / 16, A.n_tiles), 128, 0, c->stream>>>(A);
                CK_LAUNCH(c);
            }
        }
        k_clear_foreign_level0<<<148 * 4, 256, 0, c->stream>>>(vs.rad_surf, (int)c->cfg.grid_n, c->cfg.nranks, c->cfg.rank, c->dev_state, set);
        CK_LAUNCH(c);
    }
    if ((rc = f184_stage_end(c, F184_STAGE_NEED))) return rc;
    if ((rc = f184_stage_begin(c, F184_STAGE_EXCHANGE))) return rc;
    if ((rc = f184_zero_counters(c, 1u << F184_COUNTER_GATHER_BYTES))) return rc;
    {
        static const int gather_ctas = [] { const char* e = getenv("F184_GATHER_CTAS"); return e && atoi(e) > 0 ? atoi(e) : 148 * 2; }();     // (tests: few CTAs = many units each)
        k_gather_bricks_tma<<<gather_ctas, (G4_CONSUMERS + 1) * 32, G4_SMEM_BYTES, c->stream>>>(G, c->dev_state, c->need1, c->l1_nonzero[set]);
    }
    CK_LAUNCH(c);
    k_gather_state<<<1, 1, 0, c->stream>>>(c->dev_state, set);
    CK_LAUNCH(c);
    rc = f184_stage_end(c, F184_STAGE_EXCHANGE);
    if (rc) return rc;
    if ((rc = f184_stage_begin(c, F184_STAGE_TAIL))) return rc;
    rc = f184_mips_tail_n(c, false);
    if (rc) return rc;
    if ((rc = f184_stage_end(c, F184_STAGE_TAIL))) return rc;
    return f184_volume_publish(c);        // every rank's bricks are in: this set is what the next trace samples
}
