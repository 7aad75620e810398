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
