// debug_hooks.cu — test hook: evaluate f184_detmath.h ON THE DEVICE so tests can show it is bit-identical
// to the host evaluation (the premise of the bit-exact parity tests).  Host pointers in, host pointers out.
#include "f184_device.cuh"

namespace {
__global__ void k_detmath(uint32_t op, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r = 0.f;
    switch (op)
    {
    case 0: r = dm_sin(x[i]); break;
    case 1: r = dm_cos(x[i]); break;
    case 2: r = dm_log(x[i]); break;
    case 3: r = dm_log2(x[i]); break;
    case 4: r = dm_exp2(x[i]); break;
    case 5: r = dm_pow(x[i], y[i]); break;
    case 6: r = dm_f16_to_f32(dm_f32_to_f16(x[i])); break;
    case 7: r = dm_u2f((uint32_t)dm_f2i(x[i])); break;                    // conversions: results returned as raw bits
    case 8: r = dm_u2f(dm_f2uint(x[i])); break;
    case 9: r = dm_u2f((uint32_t)dm_f32_to_f16(x[i])); break;
    case 10: r = dm_f16_to_f32((uint16_t)(dm_f2u(x[i]) & 0xffffu)); break;
    }
    out[i] = r;
}
}  // namespace

extern "C" int f184_debug_detmath(f184_ctx* c, uint32_t op, const float* x, const float* y, float* out, size_t n)
{
    if (!c || op > 10 || !x || !out) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "debug_detmath: bad argument");
    float *dx = nullptr, *dy = nullptr, *dout = nullptr;
    CK(c, cudaMalloc(&dx, n * 4));
    CK(c, cudaMalloc(&dy, n * 4));
    CK(c, cudaMalloc(&dout, n * 4));
    CK(c, cudaMemcpy(dx, x, n * 4, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(dy, y ? y : x, n * 4, cudaMemcpyHostToDevice));
    k_detmath<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(op, dx, dy, dout, n);
    CK_LAUNCH(c);
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return F184_OK;
}

// Test hook: install `device_ptr` as rank `peer_rank`'s shareable buffer `buffer` (f184_ipc_buffer) without CUDA IPC — for
// contexts that live in ONE process (cudaIpcOpenMemHandle refuses handles of the same process).  Lets the test-suite drive the
// multi-rank slab schedule with several contexts on a single GPU ("loopback ranks"), and a barrier with a peer that never arrives.
extern "C" int f184_debug_set_peer(f184_ctx* c, uint32_t peer_rank, uint32_t buffer, void* device_ptr)
{
    if (!c || peer_rank >= 8 || peer_rank >= c->cfg.nranks || peer_rank == c->cfg.rank || buffer >= F184_IPC_COUNT || !device_ptr)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "debug_set_peer: bad argument");
    c->peer[peer_rank].buf[buffer] = device_ptr;
    c->peer[peer_rank].imported[buffer] = false;       // not ours to close
    return F184_OK;
}
// own pointer of a shareable buffer (allocated on first use), the counterpart of f184_debug_set_peer
extern "C" int f184_debug_get_ipc_ptr(f184_ctx* c, uint32_t buffer, void** out)
{
    if (!c || buffer >= F184_IPC_COUNT || !out) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "debug_get_ipc_ptr: bad argument");
    CK(c, cudaSetDevice(c->cfg.device));
    int rc = f184_prepare_frame(c);        // loopback ranks share a device: no first-use allocation may happen while another rank waits in a barrier
    if (rc) return rc;
    rc = f184_ipc_buffer_ptr(c, buffer, out);
    if (rc) return rc;
    rc = f184_join_internal(c);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    return F184_OK;
}
