// f184_api.cu — the C-ABI of libf184 (include/f184.h): context, scene/texture upload, image slots,
// stream/interop plumbing, stage timing.  The passes themselves live in the mode_*.cu / gtao.cu / blur.cu
// translation units.  There is deliberately no host fallback anywhere in this library.
#include <algorithm>
#include <cstdarg>

#include "f184_internal.h"

static std::string g_create_err;

int f184_fail(f184_ctx* c, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_err = buf;
    return code;
}

static uint32_t fmt_bytes(uint32_t f)
{
    switch (f)
    {
    case F184_FMT_R32_SFLOAT: return 4;
    case F184_FMT_R16G16B16A16_UNORM: return 8;
    case F184_FMT_R8G8B8A8_UNORM: return 4;
    case F184_FMT_R16G16B16A16_SFLOAT: return 8;
    case F184_FMT_R16G16_UINT: return 4;
    case F184_FMT_R32G32B32A32_SFLOAT: return 16;
    case F184_FMT_R8G8B8A8_SNORM: return 4;
    case F184_FMT_R32_UINT: return 4;
    }
    return 0;
}

static uint64_t mips_total_texels(uint32_t n)
{
    uint64_t t = 0;
    for (uint32_t s = n / 2; s >= 1; s /= 2) t += 6ull * s * s * s;
    return t;
}

static bool default_desc(const f184_config& c, int slot, f184_image_desc* d)
{
    memset(d, 0, sizeof(*d));
    const uint32_t W = c.width, H = c.height, S = c.shadow_res, N = c.grid_n;
    auto set = [&](uint32_t f, uint32_t w, uint32_t h, uint32_t dep) {
        d->format = f; d->width = w; d->height = h; d->depth = dep;
        d->row_pitch_bytes = w * fmt_bytes(f);
        d->size_bytes = (uint64_t)w * h * dep * fmt_bytes(f);
    };
    switch (slot)
    {
    case F184_SLOT_DEPTH: set(F184_FMT_R32_SFLOAT, W, H, 1); break;
    case F184_SLOT_NORMALS: set(F184_FMT_R16G16B16A16_UNORM, W, H, 1); break;
    case F184_SLOT_ALBEDO:
    case F184_SLOT_MATERIAL: set(F184_FMT_R8G8B8A8_UNORM, W, H, 1); break;
    case F184_SLOT_SHADOW: set(F184_FMT_R32_SFLOAT, S, S, 1); break;
    case F184_SLOT_VOXELS: set(F184_FMT_R16G16_UINT, N, N, N); break;
    case F184_SLOT_INDIRECT_OUT:
    case F184_SLOT_INDIRECT_HISTORY:
    case F184_SLOT_AO_RAW:
    case F184_SLOT_AO_OUT:
    case F184_SLOT_INDIRECT_BLUR_X:
    case F184_SLOT_LIGHTING:
    case F184_SLOT_TAA_HISTORY:
    case F184_SLOT_TAA_OUT:
    case F184_SLOT_COLOR_OUT:
    case F184_SLOT_INDIRECT_FINAL: set(F184_FMT_R16G16B16A16_SFLOAT, W, H, 1); break;
    case F184_SLOT_ACCUM_COLOR:
    case F184_SLOT_ACCUM_NORMAL: set(F184_FMT_R32G32B32A32_SFLOAT, N, N, N); break;
    case F184_SLOT_VOX_ALBEDO: set(F184_FMT_R8G8B8A8_UNORM, N, N, N); break;
    case F184_SLOT_VOX_NORMAL: set(F184_FMT_R8G8B8A8_SNORM, N, N, N); break;
    case F184_SLOT_RADIANCE: set(F184_FMT_R8G8B8A8_UNORM, N, N, N); break;
    case F184_SLOT_MIPS:
        d->format = F184_FMT_R8G8B8A8_UNORM; d->width = (uint32_t)mips_total_texels(N); d->height = 1; d->depth = 1;
        d->row_pitch_bytes = 0; d->size_bytes = mips_total_texels(N) * 4; break;
    case F184_SLOT_BRICK_FLAGS: set(F184_FMT_R32_UINT, N / 8, N / 8, N / 8); break;
    default: return false;
    }
    return true;
}

int f184_ensure_image(f184_ctx* c, int slot)
{
    DevImage& im = c->img[slot];
    if (im.upload_pending & (1u << c->cur_sid))
    {   // a pass (or a read-back) is about to touch the slot: order the stream it runs on after the upload that is in flight on
        // the copy stream (each of the three streams joins the upload once)
        CK(c, cudaStreamWaitEvent(c->stream, im.ev_up[im.cur], 0));
        im.upload_pending &= ~(1u << c->cur_sid);
    }
    if (im.ptr) return F184_OK;
    f184_image_desc d;
    if (!default_desc(c->cfg, slot, &d)) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "bad slot %d", slot);
    CK(c, cudaMalloc(&im.ptr, d.size_bytes));
    CK(c, cudaMemsetAsync(im.ptr, 0, d.size_bytes, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));     // first use only: the zero-fill has landed before ANY stream of the pipeline can touch the slot
    im.owned = true;
    im.buf[0] = im.ptr; im.cur = 0;
    im.desc = d;
    im.desc.device_ptr = im.ptr;
    return F184_OK;
}

int f184_stage_begin(f184_ctx* c, int stage)
{
    if (c->ev_accumulate)
    {   // a fresh event pair per run so the whole timed region can be summed afterwards without a sync inside it
        auto& pool = c->ev_pool[stage];
        const uint32_t run = c->ev_runs[stage];
        if (pool.size() < 2ull * (run + 1))
        {
            cudaEvent_t a, b;
            CK(c, cudaEventCreate(&a));
            CK(c, cudaEventCreate(&b));
            pool.push_back(a); pool.push_back(b);
        }
        c->ev[stage][0] = pool[2 * run];
        c->ev[stage][1] = pool[2 * run + 1];
        c->ev_runs[stage] = run + 1;
    }
    CK(c, cudaEventRecord(c->ev[stage][0], c->stream));
    return F184_OK;
}
int f184_stage_end(f184_ctx* c, int stage)
{
    CK(c, cudaEventRecord(c->ev[stage][1], c->stream));
    c->ev_valid[stage] = true;
    return F184_OK;
}

// ---- frame pipeline (f184_internal.h; DESIGN.md "Frame pipeline") ------------------------------------
bool f184_pipelined(const f184_ctx* c)
{
    return c->vox_stream && c->build_stream && c->cfg.mode == F184_MODE_NORTHSTAR && !(c->cfg.flags & F184_FLAG_NO_OVERLAP);
}
int f184_enter(f184_ctx* c, int sid, F184Section* s)
{
    s->saved = c->stream; s->saved_sid = c->cur_sid; s->switched = false;
    if (!f184_pipelined(c) || c->cur_sid != F184_SID_PASS) return F184_OK;
    cudaStream_t target = sid == F184_SID_VOX ? c->vox_stream : c->build_stream;
    if (c->pass_dirty & (1u << sid))
    {   // the pass stream wrote something this stream reads (first use: allocation clears, scene/table uploads; later: uploads into the
        // voxelizer's slots, an imported semaphore wait, caller-owned input images): wait for THAT point, not for the traces queued since
        CK(c, cudaStreamWaitEvent(target, c->ev_pass_point, 0));
        c->pass_dirty &= ~(1u << sid);
    }
    c->stream = target; c->cur_sid = sid; s->switched = true;
    return F184_OK;
}
int f184_leave(f184_ctx* c, const F184Section& s, int rc)
{
    if (!s.switched) return rc;
    cudaError_t e = cudaSuccess;
    if (c->cur_sid == F184_SID_VOX) { e = cudaEventRecord(c->ev_vox_done, c->stream); c->vox_pending = true; c->vox_to_build = true; }
    else { e = cudaEventRecord(c->ev_build_tail, c->stream); c->build_pending = true; }
    c->stream = s.saved; c->cur_sid = s.saved_sid;
    if (rc) return rc;
    CK(c, e);
    return F184_OK;
}
int f184_build_wait_vox(f184_ctx* c)
{
    if (c->vox_to_build && c->cur_sid == F184_SID_BUILD)
    {
        CK(c, cudaStreamWaitEvent(c->stream, c->ev_vox_done, 0));
        c->vox_to_build = false;
    }
    return F184_OK;
}
int f184_join_internal(f184_ctx* c)
{
    if (c->cur_sid != F184_SID_PASS) return F184_OK;
    if (c->vox_pending) { CK(c, cudaStreamWaitEvent(c->stream, c->ev_vox_done, 0)); c->vox_pending = false; }
    if (c->build_pending) { CK(c, cudaStreamWaitEvent(c->stream, c->ev_build_tail, 0)); c->build_pending = false; }
    return F184_OK;
}
int f184_pass_wrote(f184_ctx* c)
{
    if (!c->ev_pass_point || c->cur_sid != F184_SID_PASS) return F184_OK;
    CK(c, cudaEventRecord(c->ev_pass_point, c->stream));
    c->pass_dirty = (1u << F184_SID_VOX) | (1u << F184_SID_BUILD);
    return F184_OK;
}
int f184_volume_begin_write(f184_ctx* c)
{
    if (c->volume_open) return F184_OK;
    VolumeSet& v = c->vs[c->build_set];
    if (f184_pipelined(c) && v.traced_valid) CK(c, cudaStreamWaitEvent(c->stream, v.ev_traced, 0));
    c->volume_open = true;
    c->inject_in_volume = false;
    return F184_OK;
}
int f184_volume_publish(f184_ctx* c)
{
    VolumeSet& v = c->vs[c->build_set];
    if (f184_pipelined(c)) { CK(c, cudaEventRecord(v.ev_built, c->stream)); v.built_valid = true; }
    c->trace_set = c->build_set;
    c->build_set = (c->build_set + 1) % (c->n_sets ? c->n_sets : 1);
    c->volume_open = false;
    return F184_OK;
}
int f184_volume_acquire(f184_ctx* c, VolumeSet** out)
{
    VolumeSet& v = c->vs[c->trace_set];
    if (f184_pipelined(c) && v.built_valid) CK(c, cudaStreamWaitEvent(c->stream, v.ev_built, 0));
    *out = &v;
    return F184_OK;
}
int f184_volume_release(f184_ctx* c, VolumeSet* v)
{
    if (f184_pipelined(c)) { CK(c, cudaEventRecord(v->ev_traced, c->stream)); v->traced_valid = true; }
    return F184_OK;
}
int f184_check_device_errors(f184_ctx* c)
{
    if (!c->dev_state) return F184_OK;
    uint32_t e = 0;
    CK(c, cudaMemcpy(&e, c->dev_state + F184_DEV_ERROR, sizeof(e), cudaMemcpyDeviceToHost));
    if (!e) return F184_OK;
    CK(c, cudaMemset(c->dev_state + F184_DEV_ERROR, 0, sizeof(e)));       // reported once
    if (e & F184_DEVERR_BARRIER_TIMEOUT)
        return f184_fail(c, F184_ERR_PEER_TIMEOUT, "rank %u: a peer did not reach f184_peer_barrier within the timeout; the volumes of the frames since are undefined",
                         c->cfg.rank);
    return f184_fail(c, F184_ERR_NOT_READY, "rank %u: a cone sampled level 0 of the volume, which the last f184_gather_volume did not fetch from the peers "
                                            "(the G-buffer changed between the gather and the trace)", c->cfg.rank);
}

// First frame: everything a frame allocates on first use (images, texture sets, lists, export records) is allocated HERE, on the
// pass stream, before anything is enqueued — first-use allocations zero-fill and synchronise their stream, and a host-side wait in
// the middle of a frame (behind a peer barrier that waits for another rank) has no place in an asynchronous schedule.
int f184_prepare_frame(f184_ctx* c)
{
    if (c->prepared || c->cfg.mode != F184_MODE_NORTHSTAR) return F184_OK;
    int rc;
    for (int slot : {F184_SLOT_ACCUM_COLOR, F184_SLOT_ACCUM_NORMAL, F184_SLOT_VOX_ALBEDO, F184_SLOT_VOX_NORMAL, F184_SLOT_BRICK_FLAGS,
                     F184_SLOT_RADIANCE, F184_SLOT_MIPS, F184_SLOT_SHADOW, F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_MATERIAL,
                     F184_SLOT_INDIRECT_OUT, F184_SLOT_INDIRECT_HISTORY})
        if ((rc = f184_ensure_image(c, slot))) return rc;
    if ((rc = f184_mode_n_alloc(c))) return rc;
    void* dummy = nullptr;
    if ((rc = f184_ipc_buffer_ptr(c, F184_IPC_BRICK_LIST, &dummy))) return rc;
    if (c->cfg.nranks > 1)
    {
        for (uint32_t b : {(uint32_t)F184_IPC_EXPORT, (uint32_t)F184_IPC_SYNC, (uint32_t)F184_IPC_FRAG_QUEUE})
            if ((rc = f184_ipc_buffer_ptr(c, b, &dummy))) return rc;
        if ((rc = f184_gather_init_n(c))) return rc;
    }
    if ((rc = f184_mips_init_n(c))) return rc;         // kernel attributes (dynamic shared memory, carve-out): set before the first frame, not inside it
    if ((rc = f184_trace_init_n(c))) return rc;
    if (!c->gamma_table) CK(c, cudaMalloc(&c->gamma_table, 256 * sizeof(float)));
    if (!c->ev_barrier) CK(c, cudaEventCreateWithFlags(&c->ev_barrier, cudaEventDisableTiming));
    if (c->n_tris)
    {   // scene-sized scratch and the material / texture tables, when the scene is already there (it normally is)
        if ((rc = f184_voxelizer_scratch_n(c))) return rc;
        if ((rc = f184_sync_tables(c))) return rc;
    }
    CK(c, cudaStreamSynchronize(c->stream));
    c->prepared = true;
    return F184_OK;
}

__global__ void k_zero_counters(unsigned long long* __restrict__ counters, uint32_t mask)
{
    if ((mask >> threadIdx.x) & 1u) counters[threadIdx.x] = 0ull;
}
__global__ void __launch_bounds__(256) k_fill_words(uint32_t* __restrict__ p, uint32_t v, size_t n_words)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)
    {
        uint4* p4 = reinterpret_cast<uint4*>(p);
        const size_t n4 = n_words / 4;
        const uint4 v4 = make_uint4(v, v, v, v);
        for (size_t j = i; j < n4; j += stride) p4[j] = v4;
        for (size_t j = n4 * 4 + i; j < n_words; j += stride) p[j] = v;
    }
    else
        for (; i < n_words; i += stride) p[i] = v;
}
int f184_zero_counters(f184_ctx* c, uint32_t mask)
{
    k_zero_counters<<<1, 32, 0, c->stream>>>(c->counters_dev, mask);
    CK_LAUNCH(c);
    return F184_OK;
}
int f184_fill_async(f184_ctx* c, void* ptr, uint32_t value32, size_t bytes, cudaStream_t stream)
{
    const size_t n_words = bytes / 4;
    if (!n_words) return F184_OK;
    const unsigned blocks = (unsigned)std::min<size_t>((n_words / 4 + 255) / 256 + 1, 148 * 8);
    k_fill_words<<<blocks, 256, 0, stream>>>(reinterpret_cast<uint32_t*>(ptr), value32, n_words);
    CK_LAUNCH(c);
    return F184_OK;
}

// ---- texture mip chain on the device ------------------------------------------------------------
// One level per launch: mean of the 2x2 source block, rounded half-to-even back to UNORM8 — what a linear
// 2:1 vkCmdBlitImage computes (RHI/Private/Vulkan/DeviceVk.cpp:437-465).
__global__ void k_tex_downsample(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t sw, uint32_t dw, uint32_t dh)
{
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const uchar4* s = reinterpret_cast<const uchar4*>(src);
    uchar4 a = s[(size_t)(2 * y) * sw + 2 * x], b = s[(size_t)(2 * y) * sw + 2 * x + 1];
    uchar4 cc = s[(size_t)(2 * y + 1) * sw + 2 * x], d = s[(size_t)(2 * y + 1) * sw + 2 * x + 1];
    auto avg = [](uint32_t p, uint32_t q, uint32_t r, uint32_t t) {
        uint32_t sum = p + q + r + t, v = sum >> 2, rem = sum & 3;
        if (rem > 2 || (rem == 2 && (v & 1))) v++;
        return (unsigned char)v;
    };
    uchar4 o = make_uchar4(avg(a.x, b.x, cc.x, d.x), avg(a.y, b.y, cc.y, d.y), avg(a.z, b.z, cc.z, d.z), avg(a.w, b.w, cc.w, d.w));
    reinterpret_cast<uchar4*>(dst)[(size_t)y * dw + x] = o;
}

int f184_sync_tables(f184_ctx* c)
{
    if (!c->tables_dirty) return F184_OK;
    // the voxelizer of a frame still in flight on vox_stream reads the tables: let it finish before they change (a material
    // or texture change is rare; the blocking copies below then land before anything enqueued later)
    if (c->vox_stream) CK(c, cudaStreamSynchronize(c->vox_stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (c->tex_host.size() > c->tex_dev_cap)
    {
        if (c->tex_dev) cudaFree(c->tex_dev);
        c->tex_dev_cap = (uint32_t)c->tex_host.size() + 16;
        CK(c, cudaMalloc(&c->tex_dev, sizeof(TexDev) * c->tex_dev_cap));
    }
    if (c->mat_host.size() > c->mat_dev_cap)
    {
        if (c->mat_dev) cudaFree(c->mat_dev);
        c->mat_dev_cap = (uint32_t)c->mat_host.size() + 16;
        CK(c, cudaMalloc(&c->mat_dev, sizeof(MatDev) * c->mat_dev_cap));
    }
    if (!c->tex_host.empty()) CK(c, cudaMemcpy(c->tex_dev, c->tex_host.data(), sizeof(TexDev) * c->tex_host.size(), cudaMemcpyHostToDevice));
    if (!c->mat_host.empty()) CK(c, cudaMemcpy(c->mat_dev, c->mat_host.data(), sizeof(MatDev) * c->mat_host.size(), cudaMemcpyHostToDevice));
    c->tables_dirty = false;
    return F184_OK;
}

extern "C" {

int f184_abi_version(void) { return F184_ABI_VERSION; }

int f184_create(const f184_config* config, f184_ctx** out)
{
    if (!config || !out || config->struct_size != sizeof(f184_config)) return f184_fail(nullptr, F184_ERR_INVALID_ARGUMENT, "bad config");
    const uint32_t n = config->grid_n;
    if (n < 8 || n > 1024 || (n & (n - 1))) return f184_fail(nullptr, F184_ERR_INVALID_ARGUMENT, "grid_n must be a power of two in [8,1024]");
    if (!config->width || !config->height) return f184_fail(nullptr, F184_ERR_INVALID_ARGUMENT, "width/height must be non-zero");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return f184_fail(nullptr, F184_ERR_NO_DEVICE, "no CUDA device (%s); libf184 has no CPU fallback", cudaGetErrorString(e));
    if (config->device < 0 || config->device >= ndev) return f184_fail(nullptr, F184_ERR_INVALID_ARGUMENT, "device %d out of range", config->device);
    e = cudaSetDevice(config->device);
    if (e != cudaSuccess) return f184_fail(nullptr, F184_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    f184_ctx* c = new f184_ctx();
    c->cfg = *config;
    if (c->cfg.march_steps == 0) c->cfg.march_steps = 60;
    if (c->cfg.step_size == 0.f) c->cfg.step_size = 0.2f;
    if (c->cfg.shadow_res == 0) c->cfg.shadow_res = 2048;
    if (c->cfg.cone_max_distance == 0.f) c->cfg.cone_max_distance = 32.f;
    if (c->cfg.nranks == 0) c->cfg.nranks = 1;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess)
    {
        delete c;
        return f184_fail(nullptr, F184_ERR_CUDA, "cudaStreamCreate failed");
    }
    c->stream = c->own_stream;
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess)
    {
        delete c;
        return f184_fail(nullptr, F184_ERR_CUDA, "cudaStreamCreate failed");
    }
    // highest priority: the block scheduler hands free SM slots to a higher-priority stream's CTAs first, so the voxelizer's
    // CTAs move in beside the resident cone-trace CTAs instead of queueing behind that kernel's whole grid
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&c->vox_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->build_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_vox_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_normalised, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_build_tail, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_pass_point, cudaEventDisableTiming) != cudaSuccess)
    {
        delete c;
        return f184_fail(nullptr, F184_ERR_CUDA, "cudaStreamCreate failed");
    }
    c->pass_dirty = (1u << F184_SID_VOX) | (1u << F184_SID_BUILD);       // first use of either: behind everything enqueued on the pass stream so far
    cudaEventRecord(c->ev_pass_point, c->stream);
    for (int s = 0; s < F184_STAGE_COUNT; s++)
    {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        c->ev_pool[s].push_back(a); c->ev_pool[s].push_back(b);
        c->ev[s][0] = a; c->ev[s][1] = b;
    }
    if (cudaMalloc(&c->counters_dev, sizeof(unsigned long long) * 32) != cudaSuccess)
    {
        delete c;
        return f184_fail(nullptr, F184_ERR_OUT_OF_MEMORY, "cudaMalloc counters");
    }
    cudaMemset(c->counters_dev, 0, sizeof(unsigned long long) * 32);       // public counters, internal cursors (F184_COUNTER_COUNT + 0..3), diagnostics
    if (cudaMalloc(&c->dev_state, sizeof(uint32_t) * F184_DEV_WORDS) != cudaSuccess)
    {
        delete c;
        return f184_fail(nullptr, F184_ERR_OUT_OF_MEMORY, "cudaMalloc device state");
    }
    {
        const uint32_t init[F184_DEV_WORDS] = {0, 0, 1, 1, 0, 0, 0, 0};      // both texture sets start complete (all zero)
        cudaMemcpy(c->dev_state, init, sizeof(init), cudaMemcpyHostToDevice);
    }
    *out = c;
    return F184_OK;
}

void f184_destroy(f184_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->vox_stream) cudaStreamSynchronize(c->vox_stream);
    if (c->build_stream) cudaStreamSynchronize(c->build_stream);
    if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
    f184_mode_n_release(c);
    for (auto& im : c->img)
    {
        if (im.owned)
        {
            for (int i = 0; i < 2; i++)
            {
                if (im.buf[i]) cudaFree(im.buf[i]);
                if (im.ev_up[i]) cudaEventDestroy(im.ev_up[i]);
                for (int s = 0; s < F184_SID_COUNT; s++)
                    if (im.ev_release[i][s]) cudaEventDestroy(im.ev_release[i][s]);
            }
            im.ptr = nullptr;
        }
        if (im.ext) { if (im.ptr) cudaFree(im.ptr); cudaDestroyExternalMemory(im.ext); }
    }
    for (void* p : {(void*)c->pos, (void*)c->nrm, (void*)c->uv, (void*)c->model_mats, (void*)c->idx, (void*)c->tri_mat,
                    (void*)c->tri_model, (void*)c->tex_dev, (void*)c->mat_dev, (void*)c->vox_keys, (void*)c->counters_dev,
                    (void*)c->brick_prev, (void*)c->brick_list, (void*)c->vox_queue, (void*)c->vm_dev, (void*)c->gamma_table, c->gtao_phi_table,
                    (void*)c->dev_state, (void*)c->chunk_list, (void*)c->need1, (void*)c->l1_nonzero[0], (void*)c->l1_nonzero[1]})
        if (p) cudaFree(p);
    for (uint8_t* p : c->tex_alloc) if (p) cudaFree(p);
    for (int s = 0; s < F184_STAGE_COUNT; s++)
        for (cudaEvent_t e : c->ev_pool[s]) cudaEventDestroy(e);
    for (auto& pr : c->peer)
        for (int b = 0; b < F184_IPC_COUNT; b++)
            if (pr.imported[b] && pr.buf[b]) cudaIpcCloseMemHandle(pr.buf[b]);
    if (c->export_buf) cudaFree(c->export_buf);
    for (void* p : {(void*)c->cacheC, (void*)c->cacheN, (void*)c->cache_slot, (void*)c->cache_occ})
        if (p) cudaFree(p);
    for (void* p : {(void*)c->frag_queue, (void*)c->frag_counts, (void*)c->frag_cursor})
        if (p) cudaFree(p);
    if (c->sync_flags) cudaFree(c->sync_flags);
    if (c->sem_wait) cudaDestroyExternalSemaphore(c->sem_wait);
    if (c->sem_signal) cudaDestroyExternalSemaphore(c->sem_signal);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->vox_stream) cudaStreamDestroy(c->vox_stream);
    if (c->build_stream) cudaStreamDestroy(c->build_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for (int i = 0; i < 4; i++)
    {
        if (c->view_streams[i]) cudaStreamDestroy(c->view_streams[i]);
        if (c->ev_view_done[i]) cudaEventDestroy(c->ev_view_done[i]);
    }
    if (c->ev_view_fork) cudaEventDestroy(c->ev_view_fork);
    for (int i = 0; i < 2; i++)
    {
        if (c->rb_stage[i]) cudaFree(c->rb_stage[i]);
        if (c->ev_rb_snap[i]) cudaEventDestroy(c->ev_rb_snap[i]);
        if (c->ev_rb_done[i]) cudaEventDestroy(c->ev_rb_done[i]);
    }
    for (cudaEvent_t e : {c->ev_vox_done, c->ev_normalised, c->ev_build_tail, c->ev_pass_point})
        if (e) cudaEventDestroy(e);
    if (c->ev_barrier) cudaEventDestroy(c->ev_barrier);
    if (c->lights_dev) cudaFree(c->lights_dev);
    if (c->r_queue) cudaFree(c->r_queue);
    delete c;
}

const char* f184_last_error(const f184_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int f184_scene_upload(f184_ctx* c, const f184_scene_desc* s)
{
    if (!c || !s || !s->positions || !s->normals || !s->uvs || !s->indices || !s->tri_material || !s->tri_model || !s->model_mats)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "scene_upload: null argument");
    if (!s->n_tris || !s->n_verts || !s->n_models) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "scene_upload: empty scene");
    for (uint64_t i = 0; i < 3ull * s->n_tris; i++)
        if (s->indices[i] >= s->n_verts) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "scene_upload: index %llu out of range", (unsigned long long)i);
    uint32_t max_mat = 0;
    for (uint32_t i = 0; i < s->n_tris; i++)
    {
        if (s->tri_model[i] >= s->n_models) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "scene_upload: tri_model out of range");
        if (s->tri_material[i] > max_mat) max_mat = s->tri_material[i];
    }
    CK(c, cudaSetDevice(c->cfg.device));
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaStreamSynchronize(c->vox_stream));     // the voxelizer in flight reads the buffers freed below
    CK(c, cudaStreamSynchronize(c->build_stream));
    // free and forget: if an upload below fails the context is left with NO scene (n_tris = 0), not with dangling pointers
    for (void** p : {(void**)&c->pos, (void**)&c->nrm, (void**)&c->uv, (void**)&c->model_mats, (void**)&c->idx, (void**)&c->tri_mat, (void**)&c->tri_model})
        if (*p) { cudaFree(*p); *p = nullptr; }
    c->n_verts = c->n_tris = c->n_models = 0;
    auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    CK(c, up((void**)&c->pos, s->positions, 12ull * s->n_verts));
    CK(c, up((void**)&c->nrm, s->normals, 12ull * s->n_verts));
    CK(c, up((void**)&c->uv, s->uvs, 8ull * s->n_verts));
    CK(c, up((void**)&c->idx, s->indices, 12ull * s->n_tris));
    CK(c, up((void**)&c->tri_mat, s->tri_material, 2ull * s->n_tris));
    CK(c, up((void**)&c->tri_model, s->tri_model, 2ull * s->n_tris));
    CK(c, up((void**)&c->model_mats, s->model_mats, 64ull * s->n_models));
    c->model_mats_host.resize(s->n_models);
    memcpy(c->model_mats_host.data(), s->model_mats, 64ull * s->n_models);
    c->n_verts = s->n_verts; c->n_tris = s->n_tris; c->n_models = s->n_models;
    c->max_material = max_mat;
    return F184_OK;
}

int f184_texture_upload(f184_ctx* c, uint32_t id, const uint8_t* rgba, uint32_t w, uint32_t h)
{
    if (!c || !rgba || !w || !h || (w & (w - 1)) || (h & (h - 1)) || id > 65535)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "texture_upload: need power-of-two RGBA8 image");
    CK(c, cudaSetDevice(c->cfg.device));
    if (c->tex_host.size() <= id) { c->tex_host.resize(id + 1, TexDev{}); c->tex_alloc.resize(id + 1, nullptr); }
    TexDev t{};
    t.w = w; t.h = h;
    uint32_t nl = 1;
    for (uint32_t m = (w < h ? w : h); m > 1; m >>= 1) nl++;      // DeviceVk.cpp:504-509
    if (nl > F184_MAX_TEX_LEVELS) nl = F184_MAX_TEX_LEVELS;
    t.nlevels = nl;
    uint64_t total = 0;
    for (uint32_t l = 0; l < nl; l++)
    {
        t.off[l] = (uint32_t)total;
        total += 4ull * ((w >> l) ? (w >> l) : 1) * ((h >> l) ? (h >> l) : 1);
    }
    uint8_t* dev = nullptr;
    CK(c, cudaMalloc(&dev, total));
    auto build_chain = [&]() -> int {
        CK(c, cudaMemcpyAsync(dev, rgba, 4ull * w * h, cudaMemcpyHostToDevice, c->stream));
        uint32_t sw = w, sh = h;
        for (uint32_t l = 1; l < nl; l++)
        {
            uint32_t dw = sw > 1 ? sw / 2 : 1, dh = sh > 1 ? sh / 2 : 1;
            dim3 b(16, 16), g((dw + 15) / 16, (dh + 15) / 16);
            k_tex_downsample<<<g, b, 0, c->stream>>>(dev + t.off[l - 1], dev + t.off[l], sw, dw, dh);
            CK_LAUNCH(c);
            sw = dw; sh = dh;
        }
        CK(c, cudaStreamSynchronize(c->stream));     // `rgba` may be freed by the caller on return
        CK(c, cudaStreamSynchronize(c->vox_stream)); // the voxelizer in flight may still sample the texture replaced below
        return F184_OK;
    };
    if (int rc = build_chain()) { cudaFree(dev); return rc; }
    if (c->tex_alloc[id]) cudaFree(c->tex_alloc[id]);
    c->tex_alloc[id] = dev;
    t.base = dev;
    c->tex_host[id] = t;
    c->tables_dirty = true;
    return F184_OK;
}

int f184_texture_readback(f184_ctx* c, uint32_t id, uint32_t level, uint8_t* out, size_t bytes)
{
    if (!c || id >= c->tex_host.size() || !c->tex_host[id].base || level >= c->tex_host[id].nlevels)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "texture_readback: bad id/level");
    const TexDev& t = c->tex_host[id];
    size_t want = 4ull * ((t.w >> level) ? (t.w >> level) : 1) * ((t.h >> level) ? (t.h >> level) : 1);
    if (bytes != want) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "texture_readback: size mismatch");
    CK(c, cudaMemcpy(out, t.base + t.off[level], bytes, cudaMemcpyDeviceToHost));
    return F184_OK;
}

int f184_material_set(f184_ctx* c, uint32_t id, const float factor[4], int32_t tex, uint32_t use_textures)
{
    if (!c || !factor || id > 65535) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "material_set: bad argument");
    if (c->mat_host.size() <= id) c->mat_host.resize(id + 1, MatDev{{1, 1, 1, 1}, -1, 1});
    MatDev m;
    memcpy(m.factor, factor, 16);
    m.tex = tex; m.use_textures = use_textures;
    c->mat_host[id] = m;
    c->tables_dirty = true;
    return F184_OK;
}

int f184_bind_image(f184_ctx* c, uint32_t slot, const f184_image_desc* d)
{
    if (!c || slot >= F184_SLOT_COUNT || !d) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "bind_image: bad argument");
    f184_image_desc want;
    default_desc(c->cfg, slot, &want);
    DevImage& im = c->img[slot];
    if (d->device_ptr == nullptr)
    {   // unbind: back to a context-owned image on next use
        if (im.owned && im.ptr) return F184_OK;
        im = DevImage{};
        return F184_OK;
    }
    if (d->format != want.format || d->width != want.width || d->height != want.height || d->depth != want.depth)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "bind_image: slot %u expects format %u %ux%ux%u", slot, want.format, want.width, want.height, want.depth);
    if (d->row_pitch_bytes != 0 && d->row_pitch_bytes != want.row_pitch_bytes)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "bind_image: only tightly packed rows are supported");
    if (im.owned && im.ptr)
    {
        cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->vox_stream); cudaStreamSynchronize(c->build_stream);
        for (int i = 0; i < 2; i++) { if (im.buf[i]) cudaFree(im.buf[i]); im.buf[i] = nullptr; }
        im.upload_pending = 0;
    }
    im.ptr = d->device_ptr;
    im.owned = false;
    im.desc = want;
    im.desc.device_ptr = d->device_ptr;
    return F184_OK;
}

int f184_image_info(f184_ctx* c, uint32_t slot, f184_image_desc* out)
{
    if (!c || slot >= F184_SLOT_COUNT || !out) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "image_info: bad argument");
    int rc = f184_ensure_image(c, slot);
    if (rc) return rc;
    *out = c->img[slot].desc;
    return F184_OK;
}

// Copy only the rows this rank traces (f184_set_trace_rows / f184_set_trace_tiles) of a W x H image between two buffers with
// the same layout: the selected 8-row tile rows form a regular 2D pattern (8 rows wide, every `stride`-th tile), so one
// cudaMemcpy2DAsync moves them; a partial last tile goes separately.  Falls back to the whole image for other shapes.
static int copy_selected_rows(f184_ctx* c, const DevImage& im, void* dst, const void* src, cudaMemcpyKind kind, cudaStream_t st)
{
    const uint32_t H = c->cfg.height;
    const size_t rb = im.desc.row_pitch_bytes;
    uint32_t y0, y1, tile0, stride;
    const uint32_t ntiles = f184_trace_tiles(c, H, &y0, &y1, &tile0, &stride);
    if (im.desc.height != H || im.desc.depth != 1 || rb == 0 || (stride > 1 && (y0 & 7)))
    {
        CK(c, cudaMemcpyAsync(dst, src, im.desc.size_bytes, kind, st));
        return F184_OK;
    }
    if (!ntiles) return F184_OK;
    const uint32_t first_row = y0 + tile0 * 8;
    // tiles i = 0..ntiles-1 start at row first_row + i * 8 * stride; the last one may be cut by y1
    const uint32_t last_start = first_row + (ntiles - 1) * 8 * stride;
    const uint32_t last_rows = (y1 - last_start) < 8 ? (y1 - last_start) : 8;
    const uint32_t full = last_rows == 8 ? ntiles : ntiles - 1;
    char* d = static_cast<char*>(dst);
    const char* s_ = static_cast<const char*>(src);
    if (full) CK(c, cudaMemcpy2DAsync(d + (size_t)first_row * rb, (size_t)8 * stride * rb, s_ + (size_t)first_row * rb, (size_t)8 * stride * rb,
                                      (size_t)8 * rb, full, kind, st));
    if (full != ntiles) CK(c, cudaMemcpyAsync(d + (size_t)last_start * rb, s_ + (size_t)last_start * rb, (size_t)last_rows * rb, kind, st));
    return F184_OK;
}

// slots the internal streams of the frame pipeline write (and the pass stream only reads back / uploads in tests)
static bool volume_slot(uint32_t slot)
{
    return slot == F184_SLOT_ACCUM_COLOR || slot == F184_SLOT_ACCUM_NORMAL || slot == F184_SLOT_VOX_ALBEDO || slot == F184_SLOT_VOX_NORMAL ||
           slot == F184_SLOT_BRICK_FLAGS || slot == F184_SLOT_RADIANCE || slot == F184_SLOT_MIPS;
}

static int upload_impl(f184_ctx* c, uint32_t slot, const void* host, size_t bytes, bool selected_rows);
int f184_upload_image(f184_ctx* c, uint32_t slot, const void* host, size_t bytes) { return upload_impl(c, slot, host, bytes, false); }
int f184_upload_image_rows(f184_ctx* c, uint32_t slot, const void* host, size_t bytes) { return upload_impl(c, slot, host, bytes, true); }

static int upload_impl(f184_ctx* c, uint32_t slot, const void* host, size_t bytes, bool selected_rows)
{
    if (!c || slot >= F184_SLOT_COUNT || !host) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "upload_image: bad argument");
    int rc = f184_ensure_image(c, slot);
    if (rc) return rc;
    DevImage& im = c->img[slot];
    if (bytes != im.desc.size_bytes) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "upload_image: slot %u is %llu bytes, got %zu", slot, (unsigned long long)im.desc.size_bytes, bytes);
    if (volume_slot(slot))
    {   // the pipeline's own slots: written in stream order behind whatever is in flight on the internal streams, which then
        // wait for this copy before their next kernel
        rc = f184_join_internal(c);
        if (rc) return rc;
        CK(c, cudaMemcpyAsync(im.ptr, host, bytes, cudaMemcpyHostToDevice, c->stream));
        return f184_pass_wrote(c);
    }
    if (!im.owned || im.ext)
    {   // caller-owned memory: plain stream-ordered copy on the pass stream; the internal streams order themselves behind it
        if (selected_rows) rc = copy_selected_rows(c, im, im.ptr, host, cudaMemcpyHostToDevice, c->stream);
        else { CK(c, cudaMemcpyAsync(im.ptr, host, bytes, cudaMemcpyHostToDevice, c->stream)); rc = F184_OK; }
        if (rc) return rc;
        return f184_pass_wrote(c);
    }
    const int nb = im.cur ^ 1;
    if (!im.buf[nb])
    {
        CK(c, cudaMalloc(&im.buf[nb], im.desc.size_bytes));
        for (int i = 0; i < 2; i++)
        {
            CK(c, cudaEventCreateWithFlags(&im.ev_up[i], cudaEventDisableTiming));
            for (int s = 0; s < F184_SID_COUNT; s++) CK(c, cudaEventCreateWithFlags(&im.ev_release[i][s], cudaEventDisableTiming));
        }
    }
    // everything enqueued so far — on the pass stream and, for the slots the build stream reads (the shadow map), on the
    // build stream — may still read buf[cur]; nothing enqueued later will
    {
        const bool pipe = f184_pipelined(c);
        cudaStream_t streams[F184_SID_COUNT] = {c->stream, pipe ? c->vox_stream : nullptr, pipe ? c->build_stream : nullptr};
        const uint32_t readers = slot == F184_SLOT_SHADOW ? ((1u << F184_SID_PASS) | (1u << F184_SID_BUILD)) : (1u << F184_SID_PASS);
        im.release_valid[im.cur] = 0;
        for (int s = 0; s < F184_SID_COUNT; s++)
            if ((readers & (1u << s)) && streams[s])
            {
                CK(c, cudaEventRecord(im.ev_release[im.cur][s], streams[s]));
                im.release_valid[im.cur] |= 1u << s;
            }
        for (int s = 0; s < F184_SID_COUNT; s++)
            if (im.release_valid[nb] & (1u << s)) CK(c, cudaStreamWaitEvent(c->copy_stream, im.ev_release[nb][s], 0));
    }
    if (selected_rows) { rc = copy_selected_rows(c, im, im.buf[nb], host, cudaMemcpyHostToDevice, c->copy_stream); if (rc) return rc; }
    else CK(c, cudaMemcpyAsync(im.buf[nb], host, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    CK(c, cudaEventRecord(im.ev_up[nb], c->copy_stream));
    im.cur = nb;
    im.ptr = im.buf[nb];
    im.desc.device_ptr = im.ptr;
    im.upload_pending = (1u << F184_SID_COUNT) - 1u;        // every stream joins the upload before its next use of the slot
    return F184_OK;
}

static int readback_impl(f184_ctx* c, uint32_t slot, void* host, size_t bytes, bool selected_rows);
int f184_readback_async(f184_ctx* c, uint32_t slot, void* host, size_t bytes) { return readback_impl(c, slot, host, bytes, false); }
int f184_readback_async_rows(f184_ctx* c, uint32_t slot, void* host, size_t bytes) { return readback_impl(c, slot, host, bytes, true); }

static int readback_impl(f184_ctx* c, uint32_t slot, void* host, size_t bytes, bool selected_rows)
{
    if (!c || slot >= F184_SLOT_COUNT || !host) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "readback: bad argument");
    int rc = f184_ensure_image(c, slot);
    if (rc) return rc;
    if (bytes > c->img[slot].desc.size_bytes) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "readback: slot %u is %llu bytes, got %zu", slot, (unsigned long long)c->img[slot].desc.size_bytes, bytes);
    const bool vox_slot = volume_slot(slot);
    if (vox_slot) { rc = f184_join_internal(c); if (rc) return rc; }
    if (bytes > (256ull << 20) || vox_slot)
    {   // volume-sized slots (tests): on the pass stream; the internal streams wait for the copy before they overwrite the slot
        CK(c, cudaMemcpyAsync(host, c->img[slot].ptr, bytes, cudaMemcpyDeviceToHost, c->stream));
        if (vox_slot) return f184_pass_wrote(c);
        return F184_OK;
    }
    if (!c->d2h_stream)
    {
        CK(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++)
        {
            CK(c, cudaEventCreateWithFlags(&c->ev_rb_snap[i], cudaEventDisableTiming));
            CK(c, cudaEventCreateWithFlags(&c->ev_rb_done[i], cudaEventDisableTiming));
        }
    }
    const int b = c->rb_cur ^ 1;
    if (c->rb_valid[b]) CK(c, cudaStreamWaitEvent(c->stream, c->ev_rb_done[b], 0));      // its previous PCIe copy has drained
    if (c->rb_cap[b] < bytes)
    {
        if (c->rb_stage[b]) { CK(c, cudaEventSynchronize(c->ev_rb_done[b])); CK(c, cudaFree(c->rb_stage[b])); c->rb_stage[b] = nullptr; c->rb_cap[b] = 0; }
        CK(c, cudaMalloc(&c->rb_stage[b], bytes));
        c->rb_cap[b] = bytes;
    }
    const bool rows = selected_rows && bytes == c->img[slot].desc.size_bytes;
    if (rows) { rc = copy_selected_rows(c, c->img[slot], c->rb_stage[b], c->img[slot].ptr, cudaMemcpyDeviceToDevice, c->stream); if (rc) return rc; }
    else CK(c, cudaMemcpyAsync(c->rb_stage[b], c->img[slot].ptr, bytes, cudaMemcpyDeviceToDevice, c->stream));
    CK(c, cudaEventRecord(c->ev_rb_snap[b], c->stream));
    CK(c, cudaStreamWaitEvent(c->d2h_stream, c->ev_rb_snap[b], 0));
    if (rows) { rc = copy_selected_rows(c, c->img[slot], host, c->rb_stage[b], cudaMemcpyDeviceToHost, c->d2h_stream); if (rc) return rc; }
    else CK(c, cudaMemcpyAsync(host, c->rb_stage[b], bytes, cudaMemcpyDeviceToHost, c->d2h_stream));
    CK(c, cudaEventRecord(c->ev_rb_done[b], c->d2h_stream));
    c->rb_valid[b] = true;
    c->rb_cur = b;
    return F184_OK;
}

int f184_readback_wait(f184_ctx* c, uint32_t age)
{
    if (!c || age > 1) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "readback_wait: age must be 0 or 1");
    const int b = age ? c->rb_cur ^ 1 : c->rb_cur;
    if (c->rb_valid[b]) CK(c, cudaEventSynchronize(c->ev_rb_done[b]));
    return F184_OK;
}

int f184_readback(f184_ctx* c, uint32_t slot, void* host, size_t bytes)
{
    if (c && slot < F184_SLOT_COUNT && c->img[slot].ptr && bytes != c->img[slot].desc.size_bytes)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "readback: slot %u is %llu bytes, got %zu", slot, (unsigned long long)c->img[slot].desc.size_bytes, bytes);
    int rc = f184_readback_async(c, slot, host, bytes);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    if (c->d2h_stream) CK(c, cudaStreamSynchronize(c->d2h_stream));
    return F184_OK;
}

// ---- Vulkan interop: exported VkDeviceMemory / VkSemaphore file descriptors ----------------------
int f184_import_external_memory_fd(f184_ctx* c, uint32_t slot, int fd, uint64_t alloc_size, uint64_t offset, const f184_image_desc* layout)
{
    if (!c || slot >= F184_SLOT_COUNT || fd < 0 || !layout) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "import_external_memory_fd: bad argument");
    f184_image_desc want;
    default_desc(c->cfg, slot, &want);
    if (layout->format != want.format || layout->width != want.width || layout->height != want.height || layout->depth != want.depth)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "import_external_memory_fd: layout does not match slot %u", slot);
    if (offset + want.size_bytes > alloc_size) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "import_external_memory_fd: allocation too small");
    cudaExternalMemoryHandleDesc hd{};
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    hd.size = alloc_size;
    cudaExternalMemory_t ext = nullptr;
    CK(c, cudaImportExternalMemory(&ext, &hd));
    cudaExternalMemoryBufferDesc bd{};
    bd.offset = offset;
    bd.size = want.size_bytes;
    void* ptr = nullptr;
    cudaError_t e = cudaExternalMemoryGetMappedBuffer(&ptr, ext, &bd);
    if (e != cudaSuccess) { cudaDestroyExternalMemory(ext); return f184_fail(c, F184_ERR_CUDA, "cudaExternalMemoryGetMappedBuffer: %s", cudaGetErrorString(e)); }
    DevImage& im = c->img[slot];
    if (im.owned && im.ptr)
    {
        cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->vox_stream); cudaStreamSynchronize(c->build_stream);
        for (int i = 0; i < 2; i++) { if (im.buf[i]) cudaFree(im.buf[i]); im.buf[i] = nullptr; }
        im.upload_pending = 0;
    }
    im.ptr = ptr; im.owned = false; im.ext = ext; im.desc = want; im.desc.device_ptr = ptr;
    return F184_OK;
}

int f184_import_semaphores_fd(f184_ctx* c, int wait_fd, int signal_fd)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    auto imp = [&](int fd, cudaExternalSemaphore_t* out) -> cudaError_t {
        cudaExternalSemaphoreHandleDesc d{};
        d.type = cudaExternalSemaphoreHandleTypeOpaqueFd;
        d.handle.fd = fd;
        return cudaImportExternalSemaphore(out, &d);
    };
    if (wait_fd >= 0) CK(c, imp(wait_fd, &c->sem_wait));
    if (signal_fd >= 0) CK(c, imp(signal_fd, &c->sem_signal));
    return F184_OK;
}

int f184_frame_begin(f184_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    if (c->sem_wait)
    {
        cudaExternalSemaphoreWaitParams p{};
        CK(c, cudaWaitExternalSemaphoresAsync(&c->sem_wait, &p, 1, c->stream));
        return f184_pass_wrote(c);        // the internal streams of the frame pipeline start behind the wait too
    }
    return F184_OK;
}
int f184_frame_end(f184_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    int rc = f184_join_internal(c);
    if (rc) return rc;
    if (c->sem_signal)
    {
        cudaExternalSemaphoreSignalParams p{};
        CK(c, cudaSignalExternalSemaphoresAsync(&c->sem_signal, &p, 1, c->stream));
    }
    return F184_OK;
}

int f184_set_stream(f184_ctx* c, void* s)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->vox_stream);
    cudaStreamSynchronize(c->build_stream);
    if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
    c->vox_pending = c->build_pending = false;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return F184_OK;
}
int f184_sync(f184_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    int rc = f184_join_internal(c);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    if (c->d2h_stream) CK(c, cudaStreamSynchronize(c->d2h_stream));
    return c->cfg.nranks > 1 ? f184_check_device_errors(c) : F184_OK;
}

// ---- passes (dispatch on mode) ------------------------------------------------------------------
// every material a triangle names has a table entry, and every texture a material names has been uploaded: the kernels index
// mats[tri_material] and texs[mat.tex] unguarded
static int check_tables(f184_ctx* c)
{
    if (c->max_material >= c->mat_host.size())
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "voxelize: triangle material %u has no f184_material_set entry (%zu materials)", c->max_material, c->mat_host.size());
    for (size_t m = 0; m < c->mat_host.size(); m++)
    {
        const MatDev& md = c->mat_host[m];
        if (md.use_textures && md.tex >= 0 && ((size_t)md.tex >= c->tex_host.size() || !c->tex_host[md.tex].base))
            return f184_fail(c, F184_ERR_NOT_READY, "voxelize: material %zu uses texture %d, which has not been uploaded (f184_texture_upload)", m, md.tex);
    }
    return F184_OK;
}

int f184_voxelize(f184_ctx* c, const f184_view_constants* cam)
{
    if (!c || !cam) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "voxelize: null argument");
    if (!c->n_tris) return f184_fail(c, F184_ERR_NOT_READY, "voxelize: no scene uploaded");
    CK(c, cudaSetDevice(c->cfg.device));
    int rc = check_tables(c);
    if (rc) return rc;
    rc = f184_sync_tables(c);
    if (rc) return rc;
    if (c->cfg.mode == F184_MODE_REFERENCE) return f184_voxelize_r(c, cam);
    rc = f184_voxelize_accumulate(c, cam);
    if (rc) return rc;
    return f184_normalise(c);
}
// Accumulation of one frame.  Pipelined: on vox_stream, beside the build and the trace of the frames before it.  It may start once
// this rank has re-zeroed its accumulators (ev_normalised) and, on one NVLink box, once EVERY rank has (the peers add into this
// rank's accumulators and this rank into theirs): that is what the last f184_peer_barrier of the previous frame — the one between
// mips and gather, reached by each rank behind its normalise — certifies (ev_barrier).
int f184_voxelize_accumulate(f184_ctx* c, const f184_view_constants* cam)
{
    if (!c || !cam) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "voxelize_accumulate: null argument");
    if (c->cfg.mode != F184_MODE_NORTHSTAR) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "voxelize_accumulate is a north-star stage");
    if (!c->n_tris) return f184_fail(c, F184_ERR_NOT_READY, "voxelize_accumulate: no scene uploaded");
    CK(c, cudaSetDevice(c->cfg.device));
    int rc = check_tables(c);
    if (rc) return rc;
    rc = f184_sync_tables(c);
    if (rc) return rc;
    if ((rc = f184_prepare_frame(c))) return rc;
    F184Section sec;
    rc = f184_enter(c, F184_SID_VOX, &sec);
    if (rc) return rc;
    if (sec.switched)
    {
        if (c->normalised_valid) CK(c, cudaStreamWaitEvent(c->stream, c->ev_normalised, 0));
        if (c->barrier_recorded) CK(c, cudaStreamWaitEvent(c->stream, c->ev_barrier, 0));
    }
    return f184_leave(c, sec, f184_voxelize_accumulate_n(c, cam));
}
int f184_normalise(f184_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    if (c->cfg.mode != F184_MODE_NORTHSTAR) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "normalise is a north-star stage");
    if (!c->brick_prev) return f184_fail(c, F184_ERR_NOT_READY, "normalise: call f184_voxelize_accumulate first");
    CK(c, cudaSetDevice(c->cfg.device));
    F184Section sec;
    int rc = f184_enter(c, F184_SID_BUILD, &sec);
    if (rc) return rc;
    rc = f184_build_wait_vox(c);
    if (rc == F184_OK) rc = f184_normalise_n(c);
    if (rc == F184_OK && sec.switched)
    {
        cudaError_t e = cudaEventRecord(c->ev_normalised, c->stream);
        if (e != cudaSuccess) rc = f184_fail(c, F184_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
        c->normalised_valid = true;
    }
    return f184_leave(c, sec, rc);
}
// Static / dynamic split: keep what the accumulators hold now (the static geometry, accumulated by the caller) instead of normalising it.
int f184_static_cache_capture(f184_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    if (c->cfg.mode != F184_MODE_NORTHSTAR) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "static_cache_capture is a north-star stage");
    if (!c->brick_prev) return f184_fail(c, F184_ERR_NOT_READY, "static_cache_capture: call f184_voxelize_accumulate (static triangles) first");
    CK(c, cudaSetDevice(c->cfg.device));
    F184Section sec;
    int rc = f184_enter(c, F184_SID_BUILD, &sec);
    if (rc) return rc;
    rc = f184_build_wait_vox(c);
    if (rc == F184_OK) rc = f184_static_cache_capture_n(c);
    if (rc == F184_OK && sec.switched)
    {   // the accumulators are clear again: the next accumulation may start (what f184_normalise signals)
        cudaError_t e = cudaEventRecord(c->ev_normalised, c->stream);
        if (e != cudaSuccess) rc = f184_fail(c, F184_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
        c->normalised_valid = true;
    }
    rc = f184_leave(c, sec, rc);
    return rc ? rc : f184_sync(c);
}
int f184_static_cache_clear(f184_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    CK(c, cudaSetDevice(c->cfg.device));
    int rc = f184_sync(c);
    return rc ? rc : f184_static_cache_clear_n(c);
}
int f184_gather_volume_view(f184_ctx* c, const f184_trace_constants* view)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    if (c->cfg.mode != F184_MODE_NORTHSTAR) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "gather_volume is a north-star stage");
    CK(c, cudaSetDevice(c->cfg.device));
    int rc = F184_OK;
    // the gather looks at the G-buffer (which levels / bricks do this rank's cones sample?): caller-owned images are written by the
    // caller's own work on the pass stream, the build stream starts behind it
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_NORMALS, F184_SLOT_MATERIAL})
        if (c->img[s].ptr && (!c->img[s].owned || c->img[s].ext)) { if ((rc = f184_pass_wrote(c))) return rc; break; }
    F184Section sec;
    rc = f184_enter(c, F184_SID_BUILD, &sec);
    if (rc) return rc;
    return f184_leave(c, sec, f184_gather_n(c, view));
}
int f184_gather_volume(f184_ctx* c) { return f184_gather_volume_view(c, nullptr); }

// ---- CUDA IPC: share a context buffer with the other ranks of the box ----------------------------------------
int f184_ipc_export(f184_ctx* c, uint32_t buffer, f184_ipc_handle* out)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(f184_ipc_handle), "handle size");
    if (!c || buffer >= F184_IPC_COUNT || !out) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "ipc_export: bad argument");
    CK(c, cudaSetDevice(c->cfg.device));
    void* p = nullptr;
    int rc = f184_ipc_buffer_ptr(c, buffer, &p);
    if (rc) return rc;
    rc = f184_join_internal(c);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));      // the allocation's zero-fill has landed before a peer can see it
    cudaIpcMemHandle_t h;
    CK(c, cudaIpcGetMemHandle(&h, p));
    memcpy(out->opaque, &h, sizeof(h));
    return F184_OK;
}
int f184_ipc_import(f184_ctx* c, uint32_t peer_rank, uint32_t buffer, const f184_ipc_handle* in)
{
    if (!c || buffer >= F184_IPC_COUNT || !in || peer_rank >= 8 || peer_rank >= c->cfg.nranks)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "ipc_import: bad argument");
    if (peer_rank == c->cfg.rank) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "ipc_import: own rank");
    CK(c, cudaSetDevice(c->cfg.device));
    cudaIpcMemHandle_t h;
    memcpy(&h, in->opaque, sizeof(h));
    void* p = nullptr;
    CK(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[peer_rank].buf[buffer] = p;
    c->peer[peer_rank].imported[buffer] = true;
    return F184_OK;
}

int f184_inject(f184_ctx* c, const f184_sun* sun, const f184_extended_matrices* m)
{
    if (!c || !sun || !m) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "inject: null argument");
    if (c->cfg.mode != F184_MODE_NORTHSTAR) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "inject is a north-star stage");
    CK(c, cudaSetDevice(c->cfg.device));
    int rc = F184_OK;
    // a caller-owned shadow map is written by the caller's own work on the pass stream: the build stream starts behind it
    if (c->img[F184_SLOT_SHADOW].ptr && (!c->img[F184_SLOT_SHADOW].owned || c->img[F184_SLOT_SHADOW].ext) && (rc = f184_pass_wrote(c))) return rc;
    F184Section sec;
    rc = f184_enter(c, F184_SID_BUILD, &sec);
    if (rc) return rc;
    rc = f184_build_wait_vox(c);
    if (rc == F184_OK) rc = f184_inject_n(c, sun, m);
    return f184_leave(c, sec, rc);
}
int f184_build_mips(f184_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    if (c->cfg.mode != F184_MODE_NORTHSTAR) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "mips are a north-star stage");
    CK(c, cudaSetDevice(c->cfg.device));
    F184Section sec;
    int rc = f184_enter(c, F184_SID_BUILD, &sec);
    if (rc) return rc;
    rc = f184_build_wait_vox(c);
    if (rc == F184_OK) rc = f184_mips_n(c);
    return f184_leave(c, sec, rc);
}
int f184_trace_indirect(f184_ctx* c, const f184_trace_constants* k)
{
    if (!c || !k) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "trace_indirect: null argument");
    CK(c, cudaSetDevice(c->cfg.device));
    return c->cfg.mode == F184_MODE_REFERENCE ? f184_trace_r(c, k) : f184_trace_n(c, k);
}
int f184_trace_views(f184_ctx* c, const f184_trace_constants* constants, uint32_t view_height, uint32_t first, uint32_t count)
{
    if (!c || !constants) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "trace_views: null argument");
    if (c->cfg.mode != F184_MODE_NORTHSTAR) return f184_fail(c, F184_ERR_UNIMPLEMENTED, "trace_views: north-star mode only");
    CK(c, cudaSetDevice(c->cfg.device));
    return f184_trace_views_n(c, constants, view_height, first, count);
}
int f184_gtao(f184_ctx* c, const f184_view_constants* view)
{
    if (!c || !view) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "gtao: null argument");
    CK(c, cudaSetDevice(c->cfg.device));
    // mode R: every pixel pinned to the shader text; mode N: north_star's 1e-2 tolerance, on the hardware's own units
    const bool fast = c->cfg.mode == F184_MODE_NORTHSTAR && !(c->cfg.flags & F184_FLAG_EXACT_SECONDARY);
    return fast ? f184_gtao_fast_impl(c, view) : f184_gtao_impl(c, view);
}
int f184_blur_indirect(f184_ctx* c, const f184_engine_miscs* miscs)
{
    if (!c || !miscs) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "blur_indirect: null argument");
    CK(c, cudaSetDevice(c->cfg.device));
    const bool fast = c->cfg.mode == F184_MODE_NORTHSTAR && !(c->cfg.flags & F184_FLAG_EXACT_SECONDARY);
    return fast ? f184_blur_fast_impl(c, miscs) : f184_blur_impl(c, miscs);
}
int f184_lighting_deferred(f184_ctx* c, const f184_view_constants* view, const f184_extended_matrices* m, const f184_light_list* point,
                           const f184_light_list* directional)
{
    if (!c || !view || !m) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "lighting_deferred: null argument");
    CK(c, cudaSetDevice(c->cfg.device));
    return f184_lighting_impl(c, view, m, point, directional);
}
int f184_composite(f184_ctx* c, const f184_trace_constants* k)
{
    if (!c || !k) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "composite: null argument");
    CK(c, cudaSetDevice(c->cfg.device));
    return f184_composite_impl(c, k);
}
int f184_copy_taa_to_history(f184_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    int rc = f184_ensure_image(c, F184_SLOT_TAA_OUT); if (rc) return rc;
    rc = f184_ensure_image(c, F184_SLOT_TAA_HISTORY); if (rc) return rc;
    CK(c, cudaMemcpyAsync(c->img[F184_SLOT_TAA_HISTORY].ptr, c->img[F184_SLOT_TAA_OUT].ptr, c->img[F184_SLOT_TAA_OUT].desc.size_bytes,
                          cudaMemcpyDeviceToDevice, c->stream));
    return F184_OK;
}
int f184_copy_indirect_to_history(f184_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    int rc = f184_ensure_image(c, F184_SLOT_INDIRECT_OUT); if (rc) return rc;
    rc = f184_ensure_image(c, F184_SLOT_INDIRECT_HISTORY); if (rc) return rc;
    CK(c, cudaMemcpyAsync(c->img[F184_SLOT_INDIRECT_HISTORY].ptr, c->img[F184_SLOT_INDIRECT_OUT].ptr,
                          c->img[F184_SLOT_INDIRECT_OUT].desc.size_bytes, cudaMemcpyDeviceToDevice, c->stream));
    return F184_OK;
}
int f184_bind_rands(f184_ctx* c, const float* r, size_t n)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    c->rands = r; c->n_rands = n;
    return F184_OK;
}
int f184_set_triangle_range(f184_ctx* c, uint32_t first, uint32_t count)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    c->tri_first = first; c->tri_count = count;
    return F184_OK;
}
int f184_set_triangle_chunks(f184_ctx* c, const uint32_t* chunk_ids, uint32_t n_chunks)
{
    if (!c || (n_chunks && !chunk_ids)) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "set_triangle_chunks: bad argument");
    CK(c, cudaSetDevice(c->cfg.device));
    CK(c, cudaStreamSynchronize(c->stream));
    if (c->vox_stream) CK(c, cudaStreamSynchronize(c->vox_stream));       // the voxelizer in flight reads the list replaced below
    if (c->chunk_list) { cudaFree(c->chunk_list); c->chunk_list = nullptr; }
    c->n_chunks = 0;
    if (!n_chunks) return F184_OK;
    CK(c, cudaMalloc(&c->chunk_list, 4ull * n_chunks));
    CK(c, cudaMemcpy(c->chunk_list, chunk_ids, 4ull * n_chunks, cudaMemcpyHostToDevice));
    c->n_chunks = n_chunks;
    return F184_OK;
}
int f184_set_trace_rows(f184_ctx* c, uint32_t y0, uint32_t y1)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    c->row0 = y0; c->row1 = y1;
    return F184_OK;
}

int f184_set_trace_tiles(f184_ctx* c, uint32_t first, uint32_t stride)
{
    if (!c || stride == 0 || first >= stride) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "set_trace_tiles: need first < stride");
    c->tile_first = first; c->tile_stride = stride;
    return F184_OK;
}

int f184_stage_time_ms(f184_ctx* c, uint32_t stage, float* ms)
{
    if (!c || stage >= F184_STAGE_COUNT || !ms) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "stage_time_ms: bad argument");
    if (!c->ev_valid[stage]) { *ms = 0.f; return F184_OK; }
    int rc = f184_join_internal(c);
    if (rc) return rc;
    CK(c, cudaEventSynchronize(c->ev[stage][1]));
    CK(c, cudaEventElapsedTime(ms, c->ev[stage][0], c->ev[stage][1]));
    return F184_OK;
}
int f184_stage_time_reset(f184_ctx* c, uint32_t accumulate)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    int rc = f184_join_internal(c);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    c->ev_accumulate = accumulate != 0;
    for (int s = 0; s < F184_STAGE_COUNT; s++)
    {
        c->ev_runs[s] = 0;
        c->ev_valid[s] = false;
        c->ev[s][0] = c->ev_pool[s][0];
        c->ev[s][1] = c->ev_pool[s][1];
    }
    return F184_OK;
}
int f184_stage_time_total(f184_ctx* c, uint32_t stage, float* ms_sum, uint32_t* runs)
{
    if (!c || stage >= F184_STAGE_COUNT || !ms_sum || !runs) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "stage_time_total: bad argument");
    *ms_sum = 0.f; *runs = 0;
    if (!c->ev_accumulate) return f184_fail(c, F184_ERR_NOT_READY, "stage_time_total: call f184_stage_time_reset(ctx, 1) first");
    int rc = f184_join_internal(c);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    for (uint32_t r = 0; r < c->ev_runs[stage]; r++)
    {
        float ms = 0.f;
        CK(c, cudaEventElapsedTime(&ms, c->ev_pool[stage][2 * r], c->ev_pool[stage][2 * r + 1]));
        *ms_sum += ms;
    }
    *runs = c->ev_runs[stage];
    return F184_OK;
}
int f184_counter_get(f184_ctx* c, uint32_t which, uint64_t* v)
{
    if (!c || which >= F184_COUNTER_COUNT || !v) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "counter_get: bad argument");
    if (which == F184_COUNTER_KERNEL_LAUNCHES) { *v = c->launches; return F184_OK; }
    unsigned long long t = 0;
    int rc = f184_join_internal(c);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaMemcpy(&t, c->counters_dev + which, sizeof(t), cudaMemcpyDeviceToHost));
    *v = t;
    if (which == F184_COUNTER_FRAGMENTS && c->cache_slot) *v += c->cache_fragments;       // the cached static geometry's share
    return c->cfg.nranks > 1 ? f184_check_device_errors(c) : F184_OK;
}

}  // extern "C"
