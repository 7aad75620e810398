// microbench.cu — the two device peaks this path is bounded by that MEASURED_PEAKS.json does not hold
// ("L2-atomic and texture-fetch peaks ... must be established by our own microbenchmarks"):
//   f184_microbench(ctx, 0, &rate)   trilinear RGBA8 3D texture fetches per second (the cone tracer's bound)
//   f184_microbench(ctx, 1, &rate)   16-byte vector reductions (red.global.add.v4.f32) per second with the voxelizer's locality:
//                                    a warp's lanes inside one 8 KB brick, bricks scattered over a buffer larger than L2
// Each runs its kernel a few times and reports the best CUDA-event time.  Measurement aids only: bench.py calls them to
// put a denominator under the trace / voxelize numbers.
#include "f184_device.cuh"

namespace {

// Every thread walks a short diagonal through the volume, neighbouring threads start one texel apart: the access pattern
// of a coherent warp of cones, L1-resident after the first touch.  ILP 4 so the issue side never limits the TEX pipe.
__global__ void __launch_bounds__(256) k_tex_rate(cudaTextureObject_t tex, int iters, float inv_n, float4* __restrict__ sink)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const float x0 = (float)(tid & 63) * inv_n, y0 = (float)((tid >> 6) & 63) * inv_n, z0 = (float)((tid >> 12) & 63) * inv_n;
    float4 acc[4] = {};
    for (int i = 0; i < iters; i++)
    {
        const float t = (float)i * inv_n * 0.37f;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const float4 s = tex3D<float4>(tex, x0 + t + (float)k * 0.013f, y0 + t * 0.5f, z0 + t * 0.25f + (float)k * 0.007f);
            acc[k].x += s.x; acc[k].y += s.y; acc[k].z += s.z; acc[k].w += s.w;
        }
    }
    float4 r = acc[0];
    r.x += acc[1].x + acc[2].x + acc[3].x; r.y += acc[1].y + acc[2].y + acc[3].y;
    if (r.x == -1.0f) sink[tid] = r;          // never true: keeps the fetches alive
}

// The voxelizer's accumulation pattern: the 32 fragments a warp emits together land in ONE 8^3 brick (8 KB of contiguous
// float4), at scattered places inside it, and consecutive bricks of a warp are far apart (a triangle's columns walk through the
// volume).  So: per iteration a warp picks a brick out of 1 GiB by hashing, its lanes pick voxels of that brick by hashing.
// (A first version scattered every lane over the whole GiB: the real kernel beat that "peak".)
__global__ void __launch_bounds__(256) k_red_rate(float4* __restrict__ buf, uint32_t brick_mask, int iters)
{
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    uint32_t hb = gwarp * 2654435761u + 12345u, hv = (gwarp * 32u + lane) * 2246822519u + 977u;
    for (int i = 0; i < iters; i++)
    {
        hb = hb * 1664525u + 1013904223u;
        hv = hv * 1664525u + 1013904223u;
        const size_t o = (size_t)((hb >> 7) & brick_mask) * 512u + ((hv >> 9) & 511u);
        atomicAdd(buf + o, make_float4(1.0f, 2.0f, 3.0f, 1.0f));      // red.global.add.v4.f32
    }
}


// ---- peer reads over NVLink: what shape of read does the gather (mode_n_shard.cu) want? -------------------------------------------
__device__ __forceinline__ uint32_t mb_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mb_smem(bar)), "r"(parity) : "memory");
}

// one warp per CTA; `issuers` lanes (1 or 32) issue bulk copies of `copy_bytes` into a ring of `depth` slots and recycle a slot as soon
// as its copy has landed (nobody reads the data: this is the transfer alone)
__global__ void __launch_bounds__(32) k_peer_bulk(const uint8_t* __restrict__ src, uint32_t n_copies, uint32_t copy_bytes, uint32_t depth, uint32_t issuers)
{
    extern __shared__ __align__(128) uint8_t mb_ring[];
    __shared__ __align__(8) uint64_t bar[512];
    const uint32_t lane = threadIdx.x;
    for (uint32_t s_ = lane; s_ < depth; s_ += 32) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb_smem(&bar[s_])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (lane >= issuers) return;
    uint32_t k = lane;
    for (;; k += issuers)
    {
        const uint32_t i = blockIdx.x + k * gridDim.x;
        if (i >= n_copies) break;
        const uint32_t slot = k % depth, use = k / depth;
        if (use) mb_wait(&bar[slot], (use - 1u) & 1u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb_smem(&bar[slot])), "r"(copy_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(mb_smem(mb_ring + (size_t)slot * copy_bytes)), "l"(src + (size_t)i * copy_bytes), "r"(copy_bytes), "r"(mb_smem(&bar[slot])) : "memory");
    }
    // drain: the last use of every slot this lane owns
    for (uint32_t slot = lane; slot < depth; slot += issuers)
    {
        // uses of this slot = number of k' < k_end with k' % depth == slot and k' % issuers == lane (depth % issuers == 0: all of the slot's uses are this lane's)
        const uint32_t per_cta = (n_copies > blockIdx.x) ? (n_copies - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
        if (slot >= per_cta) continue;
        const uint32_t uses = (per_cta - slot + depth - 1) / depth;
        mb_wait(&bar[slot], (uses - 1u) & 1u);
    }
}

// per-lane 16-byte loads, `U` in flight per thread
__global__ void __launch_bounds__(256) k_peer_ldg(const uint4* __restrict__ src, size_t n_vec, uint4* __restrict__ sink)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (size_t i = tid; i + 7 * stride < n_vec; i += 8 * stride)
    {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __ldg(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; u++) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
    }
    if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) sink[tid & 255] = acc;      // (keeps the loads alive)
}

}  // namespace

extern "C" int f184_microbench(f184_ctx* c, uint32_t which, double* out_per_second)
{
    if (!c || !out_per_second || which > 1) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "microbench: bad argument");
    CK(c, cudaSetDevice(c->cfg.device));
    cudaEvent_t e0, e1;
    CK(c, cudaEventCreate(&e0));
    CK(c, cudaEventCreate(&e1));
    float best = 1e30f;
    double ops = 0;
    if (which == 0)
    {
        const int n = 64;
        cudaArray_t arr = nullptr;
        cudaChannelFormatDesc ch = cudaCreateChannelDesc<uchar4>();
        CK(c, cudaMalloc3DArray(&arr, &ch, make_cudaExtent(n, n, n)));
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
        td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
        cudaTextureObject_t tex = 0;
        CK(c, cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        float4* sink = nullptr;
        CK(c, cudaMalloc(&sink, sizeof(float4) * 148 * 8 * 256));
        const int iters = 256, blocks = 148 * 8;
        ops = (double)blocks * 256 * iters * 4;
        for (int rep = 0; rep < 5; rep++)
        {
            CK(c, cudaEventRecord(e0, c->stream));
            k_tex_rate<<<blocks, 256, 0, c->stream>>>(tex, iters, 1.0f / n, sink);
            CK_LAUNCH(c);
            CK(c, cudaEventRecord(e1, c->stream));
            CK(c, cudaEventSynchronize(e1));
            float ms; CK(c, cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        cudaDestroyTextureObject(tex); cudaFreeArray(arr); cudaFree(sink);
    }
    else
    {
        const uint32_t n = 1u << 26;                    // 64 Mi float4 = 1 GiB  (> 126 MB of L2)
        float4* buf = nullptr;
        CK(c, cudaMalloc(&buf, sizeof(float4) * (size_t)n));
        CK(c, cudaMemsetAsync(buf, 0, sizeof(float4) * (size_t)n, c->stream));
        const int iters = 64, blocks = 148 * 16;
        ops = (double)blocks * 256 * iters;
        for (int rep = 0; rep < 4; rep++)
        {
            CK(c, cudaEventRecord(e0, c->stream));
            k_red_rate<<<blocks, 256, 0, c->stream>>>(buf, n / 512 - 1, iters);
            CK_LAUNCH(c);
            CK(c, cudaEventRecord(e1, c->stream));
            CK(c, cudaEventSynchronize(e1));
            float ms; CK(c, cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        cudaFree(buf);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *out_per_second = ops / ((double)best * 1e-3);
    return F184_OK;
}

// Peer-read shapes over NVLink (measurement aid; the context must have rank `peer`'s export buffer imported): mode 0 = bulk copies
// issued by one lane per CTA, 1 = by 32 lanes per CTA, 2 = per-lane 16-byte loads (256-thread CTAs, 8 loads in flight per thread).
// Reads `total_bytes` (<= the export buffer's size) once; reports GB/s of the best of three runs.
extern "C" int f184_microbench_peer(f184_ctx* c, uint32_t peer, uint32_t mode, uint32_t copy_bytes, uint32_t depth, uint32_t ctas, uint64_t total_bytes, double* out_gbs)
{
    if (!c || !out_gbs || mode > 2 || peer >= c->cfg.nranks || !ctas) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "microbench_peer: bad argument");
    if (mode < 2 && (copy_bytes % 16 || !copy_bytes || !depth || depth > 512 || depth % (mode == 0 ? 1u : 32u) || (uint64_t)depth * copy_bytes > 200 * 1024))
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "microbench_peer: copy size / ring depth");
    CK(c, cudaSetDevice(c->cfg.device));
    const void* src = peer == c->cfg.rank ? (const void*)c->export_buf : (const void*)c->peer[peer].buf[F184_IPC_EXPORT];
    if (!src) return f184_fail(c, F184_ERR_NOT_READY, "microbench_peer: export buffer of rank %u not imported", peer);
    const uint64_t N = c->cfg.grid_n, cap_bytes = 4096ull * ((N / 8) * (N / 8) * (N / 8) / (c->cfg.nranks ? c->cfg.nranks : 1));
    if (total_bytes > cap_bytes) total_bytes = cap_bytes;
    cudaEvent_t e0, e1;
    CK(c, cudaEventCreate(&e0));
    CK(c, cudaEventCreate(&e1));
    uint4* sink = nullptr;
    CK(c, cudaMalloc(&sink, 256 * sizeof(uint4)));
    if (mode < 2) CK(c, cudaFuncSetAttribute(k_peer_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    float best = 1e30f;
    double moved = 0;
    for (int rep = 0; rep < 4; rep++)
    {
        CK(c, cudaEventRecord(e0, c->stream));
        if (mode < 2)
        {
            const uint32_t n_copies = (uint32_t)(total_bytes / copy_bytes);
            moved = (double)n_copies * copy_bytes;
            k_peer_bulk<<<ctas, 32, (size_t)depth * copy_bytes, c->stream>>>(reinterpret_cast<const uint8_t*>(src), n_copies, copy_bytes, depth, mode == 0 ? 1u : 32u);
        }
        else
        {
            const size_t n_vec = total_bytes / 16;
            moved = (double)(n_vec / ((size_t)ctas * 256 * 8) * ((size_t)ctas * 256 * 8)) * 16.0;
            k_peer_ldg<<<ctas, 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(src), n_vec, sink);
        }
        CK_LAUNCH(c);
        CK(c, cudaEventRecord(e1, c->stream));
        CK(c, cudaEventSynchronize(e1));
        float ms; CK(c, cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;
    }
    cudaFree(sink); cudaEventDestroy(e0); cudaEventDestroy(e1);
    *out_gbs = moved / ((double)best * 1e-3) * 1e-9;
    return F184_OK;
}
