// f184_internal.h — context and helpers shared by the libf184 translation units (not installed).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/f184.h"

#define F184_MAX_TEX_LEVELS 12

// Device-side descriptors -------------------------------------------------------------------------
struct TexDev
{
    const uint8_t* base;                    // RGBA8, all levels back to back
    uint32_t w, h, nlevels;
    uint32_t off[F184_MAX_TEX_LEVELS];      // byte offset of each level
};
struct MatDev
{
    float factor[4];
    int32_t tex;
    uint32_t use_textures;
};
// mat4 in upload order (column-major): m[c*4+r]
struct M4 { float m[16]; };

struct DevImage
{
    void* ptr = nullptr;                      // what passes read / write (for an uploaded slot: the most recent upload's buffer)
    bool owned = false;
    f184_image_desc desc{};
    cudaExternalMemory_t ext = nullptr;
    // f184_upload_image on a context-owned slot: two device buffers used alternately, copies on the context's copy stream, so
    // the upload of frame f+1 overlaps the kernels of frame f (which keep reading the other buffer)
    void* buf[2] = {nullptr, nullptr};
    int cur = 0;
    uint32_t upload_pending = 0;              // bit s: stream s (F184_SID_*) has not joined ev_up[cur] yet
    uint32_t release_valid[2] = {0, 0};       // bit s: ev_release[i][s] has been recorded
    cudaEvent_t ev_up[2] = {nullptr, nullptr};       // copy stream: upload into buf[i] finished
    cudaEvent_t ev_release[2][3] = {};        // stream s: everything that could read buf[i] has been enqueued before this
};

// The three streams a north-star context enqueues on (DESIGN.md "Frame pipeline"); mode R and F184_FLAG_NO_OVERLAP use the pass stream only.
enum { F184_SID_PASS = 0, F184_SID_VOX = 1, F184_SID_BUILD = 2, F184_SID_COUNT = 3 };

// Texture-side storage of one volume (what the cone tracer samples).  Two sets alternate so that frame f+1 is built
// (inject, mips, gather on the build stream) while frame f is still being traced on the pass stream.
struct VolumeSet
{
    cudaArray_t rad_array = nullptr;          // level-0 radiance as a 3D array for hardware filtering
    // levels >= 1: ONE mipmapped 3D array holds all six directions (the "atlas"): level l has extent (n, n, 12 n) and
    // direction d lives in z = [2 d n, 2 d n + n); the n slices behind each slab stay zero, so a trilinear footprint that
    // leaves a slab reads the same zeros border addressing would give.  One texture handle for every fetch of the cone
    // tracer keeps the handle warp-uniform (six per-direction handles made ptxas wrap every TEX in a waterfall loop).
    cudaMipmappedArray_t dir_atlas = nullptr;
    cudaTextureObject_t rad_tex = 0, dir_tex = 0, dir_tex_lin = 0;   // dir_tex: nearest mip level; dir_tex_lin: linear between levels (Appendix-B spec)
    cudaSurfaceObject_t rad_surf = 0;
    cudaSurfaceObject_t dir_surf[12] = {};    // one per atlas level
    cudaEvent_t ev_built = nullptr;           // build stream: the set is complete (last writer of the frame)
    cudaEvent_t ev_traced = nullptr;          // pass stream: the last trace that reads the set has been enqueued before this
    bool built_valid = false, traced_valid = false;
};

struct MipLevelInfo
{
    uint32_t n;               // edge of this level
    uint64_t offset_texels;   // offset of direction 0 inside F184_SLOT_MIPS (directions are consecutive)
};

struct f184_ctx
{
    f184_config cfg{};
    std::string err;
    cudaStream_t stream = nullptr, own_stream = nullptr, copy_stream = nullptr;
    // frame overlap (one GPU, north-star mode): voxelize + normalise of frame f+1 run on vox_stream while the cone trace of
    // frame f still occupies the pass stream (atomics/ALU-bound against texture-pipe-bound: they share the SMs well).
    //   vox_stream waits ev_consumed = the last pass-stream work that reads what voxelize/normalise overwrite
    //                                  (inject, mips, read-backs of the volume slots)
    //   pass stream waits ev_vox_done before the first call that reads their outputs (f184_join_vox)
    // Frame pipeline (north-star mode; DESIGN.md "Frame pipeline").  Three streams, one frame each in the steady state:
    //   vox_stream    f184_voxelize_accumulate of frame f+2          (atomics / integer ALU; NVLink peer atomics on one box)
    //   build_stream  [barrier] normalise, inject, mips, [barrier, gather], tail of frame f+1   (HBM; NVLink peer loads)
    //   pass stream   cone trace of frame f                          (texture pipe)
    // ordered by events only where data flows:
    //   ev_vox_done     vox -> build    the frame's fragments are in the accumulators (this rank's share)
    //   ev_normalised   build -> vox    this rank has re-zeroed its accumulators (normalise = next frame's clear)
    //   ev_barrier      build -> vox    one box: EVERY rank has normalised (recorded behind each f184_peer_barrier)
    //   VolumeSet::ev_built / ev_traced   build <-> pass, per texture set
    //   ev_pass_point   pass -> vox/build   something the internal streams read was written on the pass stream (pass_dirty)
    cudaStream_t vox_stream = nullptr, build_stream = nullptr;
    cudaEvent_t ev_vox_done = nullptr, ev_normalised = nullptr, ev_build_tail = nullptr, ev_pass_point = nullptr;
    bool vox_pending = false;                 // vox_stream has work the pass stream has not joined
    bool build_pending = false;               // build_stream has work the pass stream has not joined (ev_build_tail)
    bool vox_to_build = false;                // ev_vox_done not yet joined by the build stream
    bool normalised_valid = false;
    uint32_t pass_dirty = 0;                  // bit s: stream s must wait for ev_pass_point before its next kernel
    int cur_sid = F184_SID_PASS;              // which of the three `stream` currently aliases
    cudaEvent_t ev_barrier = nullptr;         // recorded behind every f184_peer_barrier (accumulate of the next frame waits for it)
    bool barrier_recorded = false;
    // probe batches (f184_trace_views): independent views round-robin over a few streams, joined back into the pass stream
    cudaStream_t view_streams[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_view_done[4] = {nullptr, nullptr, nullptr, nullptr}, ev_view_fork = nullptr;
    // asynchronous read-backs (f184_readback_async): device-side snapshot on the pass stream, PCIe copy on d2h_stream
    cudaStream_t d2h_stream = nullptr;
    void* rb_stage[2] = {nullptr, nullptr};
    size_t rb_cap[2] = {0, 0};
    cudaEvent_t ev_rb_snap[2] = {nullptr, nullptr}, ev_rb_done[2] = {nullptr, nullptr};
    bool rb_valid[2] = {false, false};
    int rb_cur = 0;
    // scene (device)
    float *pos = nullptr, *nrm = nullptr, *uv = nullptr;
    M4* model_mats = nullptr;
    uint32_t* idx = nullptr;
    uint16_t *tri_mat = nullptr, *tri_model = nullptr;
    uint32_t n_verts = 0, n_tris = 0, n_models = 0;
    uint32_t max_material = 0;               // largest tri_material of the uploaded scene (checked against the tables at voxelize)
    std::vector<M4> model_mats_host;
    // textures / materials
    std::vector<TexDev> tex_host;
    std::vector<uint8_t*> tex_alloc;
    std::vector<MatDev> mat_host;
    TexDev* tex_dev = nullptr;
    MatDev* mat_dev = nullptr;
    bool tables_dirty = true;
    uint32_t tex_dev_cap = 0, mat_dev_cap = 0;
    // images
    DevImage img[F184_SLOT_COUNT];
    // mode R: ordered-store keys, (draw order + 1) << 32 | texel
    unsigned long long* vox_keys = nullptr;
    void* r_queue = nullptr;                  // mode R voxelizer: [16 B header: count] + one set-up record per large triangle
    uint32_t r_queue_cap = 0;
    // mode N
    std::vector<MipLevelInfo> mip_levels;     // index 0 = level 1
    VolumeSet vs[2];
    int n_sets = 0;                           // 0 = not allocated; 1 under F184_FLAG_NO_OVERLAP / mode R, else 2
    int build_set = 0, trace_set = 0;         // the set the next inject/mips/gather writes; the set the next trace samples
    bool volume_open = false;                 // a writer of the build set has been enqueued since the last publish
    // per brick: bit 0 = touched by the previous voxelize (the linear volumes hold it), bit 1 + s = texture set s holds it.
    // A brick is listed (normalise, inject, mips, gather) when it is touched now or any of those bits says a stale copy
    // must be overwritten — this is what clears bricks that became empty, in the linear volumes and in BOTH texture sets.
    uint32_t* brick_prev = nullptr;
    uint32_t* brick_list = nullptr;           // bricks processed this frame (touched now or last frame)
    uint32_t* vox_queue = nullptr;            // voxelizer pass-2 queue: uint2 (triangle, first task) per large triangle
    uint32_t vox_queue_cap = 0;
    M4* vm_dev = nullptr;                     // View * Model per model matrix
    uint32_t vm_cap = 0;
    void* gtao_phi_table = nullptr;           // (cos, sin) of the 64 GTAO slice angles (gtao.cu)
    void* lights_dev = nullptr;               // 2 x 100 lights (lighting.cu)
    bool gamma_ready = false;
    float* gamma_table = nullptr;             // pow(a/255, 2.2), a = 0..255 (mode_n_inject.cu)
    bool defer_normalise = false;             // multi-GPU: f184_voxelize stops after accumulation
    uint32_t n_mip_levels = 0;                // levels >= 1
    void* tma_maps = nullptr;                 // host array of CUtensorMap (mode_n_mips.cu)
    // one NVLink box: peer-mapped buffers (index = rank; own entry points at own memory)
    struct Peer
    {
        void* buf[F184_IPC_COUNT] = {};
        bool imported[F184_IPC_COUNT] = {};
    } peer[8];
    // fragment queues (one NVLink box), all in THIS rank's memory: frag_queue = nranks regions of frag_cap 16-byte records, region p =
    // the fragments this rank rasterised for rank p, which p reads over NVLink behind the barrier; frag_cursor = the append cursors
    // (= record counts, what p reads first); frag_counts: unused scratch
    // Each sender's region is split into F184_FRAG_SUBQUEUES sub-queues with a cursor of their own (a warp uses the one its index
    // selects): one cursor per destination made every warp of the GPU hammer the same address with returning atomics.  Each cursor
    // sits in a 128-byte line of its own (F184_FRAG_CURSOR_STRIDE words apart): with the 32 cursors of a destination in ONE line, one L2
    // slice served every append of the GPU — at 2 GPUs (one destination) the voxelizer took 3.4 ms for half the triangles of C4.
    uint4* frag_queue = nullptr;
    uint32_t* frag_counts = nullptr;
    uint32_t* frag_cursor = nullptr;
    uint32_t frag_cap = 0;
    // static / dynamic split (f184_static_cache_capture): the accumulators of the bricks the static geometry touches, kept sparse
    // (16 KB per brick), added to the frame's accumulators inside normalise.  cache_slot: per brick, its place in the cache or -1.
    float4* cacheC = nullptr;
    float4* cacheN = nullptr;
    int32_t* cache_slot = nullptr;
    uint32_t* cache_occ = nullptr;            // occupied voxels per cached brick (F184_COUNTER_OCCUPIED stays what a full voxelization counts)
    uint32_t n_cached = 0;
    uint64_t cache_fragments = 0;             // fragments the cached geometry made (added to F184_COUNTER_FRAGMENTS)
    float cache_cam[32] = {0};                // ViewMat | ProjMat of the voxel camera the cache was captured with
    float last_vox_cam[32] = {0};             // ... and of the last accumulation
    bool frag_pending = false;                // an accumulation has run since the last normalise: the peers' queues hold fragments for this rank
    bool frag_sent_applied = false;           // a peer barrier has followed the last accumulation: the next one starts the queues over
    uint32_t* export_buf = nullptr;           // the own bricks' export arrays (k_mips_bricks): level 0 | level 1 | levels 2, 3 + brick index; 1024 words of room per brick
    uint32_t* sync_flags = nullptr;           // [8] barrier epochs written by the peers + [8] scratch
    uint32_t barrier_epoch = 0;
    // device-side state words (F184_DEV_*): sticky error bits and the level-0 bookkeeping of the peer gather
    uint32_t* dev_state = nullptr;
    uint32_t* need1 = nullptr;                // bit per brick: level 1 of the brick is sampled by a cone of this rank's rows (k_need_bricks)
    uint32_t* l1_nonzero[2] = {nullptr, nullptr};   // per texture set, bit per brick: this rank's copy of a foreign brick's level 1 is not zero
    bool prepared = false;                    // the first f184_voxelize_accumulate has allocated everything a frame touches
    float voxel_h = 0.0f;                     // voxel size under the voxel camera of the last f184_voxelize (level-0 test of the gather)
    bool inject_in_volume = false;            // f184_inject has written the open build set (its level 0 is current)
    // sharding
    uint32_t tri_first = 0, tri_count = 0xffffffffu;
    uint32_t* chunk_list = nullptr;           // device: 128-triangle chunks this rank voxelizes (nullptr = every triangle of the range)
    uint32_t n_chunks = 0;
    uint32_t row0 = 0, row1 = 0xffffffffu;
    uint32_t tile_first = 0, tile_stride = 1;
    const float* rands = nullptr;
    size_t n_rands = 0;
    // measurement
    unsigned long long* counters_dev = nullptr;   // F184_COUNTER_COUNT
    uint64_t launches = 0;
    cudaEvent_t ev[F184_STAGE_COUNT][2] = {};     // event pair of the stage's LAST run (aliases the pool while accumulating)
    bool ev_valid[F184_STAGE_COUNT] = {};
    // accumulation over a timed region (f184_stage_time_reset / f184_stage_time_total): one event pair per run
    bool ev_accumulate = false;
    std::vector<cudaEvent_t> ev_pool[F184_STAGE_COUNT];   // 2 events per run, created on demand, reused after reset
    uint32_t ev_runs[F184_STAGE_COUNT] = {};
    // interop
    cudaExternalSemaphore_t sem_wait = nullptr, sem_signal = nullptr;
};

#define F184_FRAG_SUBQUEUES 128
#define F184_FRAG_CURSOR_STRIDE 32        /* words between two cursors: a 128-byte line each */

// dev_state words
enum { F184_DEV_ERROR = 0,       // sticky error bits (F184_DEVERR_*), reported by the next synchronous call
       F184_DEV_NEED_L0 = 1,     // this frame: some cone of this rank's rows samples level 0 (k_need_level0)
       F184_DEV_L0_FULL = 2,     // [2 + s]: level 0 of texture set s holds every rank's bricks (no gather skipped it since the last clear)
       F184_DEV_WORDS = 8 };
enum { F184_DEVERR_BARRIER_TIMEOUT = 1u, F184_DEVERR_LEVEL0_MISSING = 2u };

// Error plumbing ----------------------------------------------------------------------------------
int f184_fail(f184_ctx* c, int code, const char* fmt, ...);
#define CK(c, call)                                                                                         \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            return f184_fail((c), e__ == cudaErrorMemoryAllocation ? F184_ERR_OUT_OF_MEMORY : F184_ERR_CUDA, \
                             "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));           \
    } while (0)
#define CK_LAUNCH(c)                                                                                        \
    do {                                                                                                    \
        (c)->launches++;                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                                               \
        if (e__ != cudaSuccess)                                                                             \
            return f184_fail((c), F184_ERR_CUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

int f184_ensure_image(f184_ctx* c, int slot);
// ---- frame pipeline plumbing (f184_api.cu) ----
bool f184_pipelined(const f184_ctx* c);   // north-star mode without F184_FLAG_NO_OVERLAP: the three-stream pipeline is on
// Run the rest of an entry point on internal stream `sid`: c->stream aliases it until f184_leave.  No-op when the pipeline is off
// or the caller is already inside an internal section.  Applies the pending pass->internal dependency (pass_dirty).
struct F184Section { cudaStream_t saved; int saved_sid; bool switched; };
int f184_enter(f184_ctx* c, int sid, F184Section* s);
int f184_leave(f184_ctx* c, const F184Section& s, int rc);   // records the stream's tail event, restores c->stream; returns rc
int f184_build_wait_vox(f184_ctx* c);     // build stream: wait for the accumulation in flight on vox_stream
int f184_join_internal(f184_ctx* c);      // pass stream waits for everything in flight on vox_stream / build_stream
int f184_pass_wrote(f184_ctx* c);         // pass stream wrote something the internal streams read: they wait for this point
int f184_volume_begin_write(f184_ctx* c); // before the first writer of the build set: wait for the trace that still samples it
int f184_volume_publish(f184_ctx* c);     // behind the last writer: the build set becomes the trace set, the other one the build set
int f184_volume_acquire(f184_ctx* c, VolumeSet** out);   // pass stream: the newest complete set, ordered behind its build
int f184_volume_release(f184_ctx* c, VolumeSet* v);      // pass stream: a reader of the set has been enqueued
int f184_check_device_errors(f184_ctx* c);               // sticky device error word -> F184_ERR_* (synchronous callers only)
int f184_sync_tables(f184_ctx* c);
int f184_prepare_frame(f184_ctx* c);                     // first-use allocations of a north-star frame, before anything is enqueued
// Per-frame clears are KERNELS, not cudaMemsetAsync: a memset travels through the copy-engine queue, which all streams of a process
// share — one that waits (in stream order) behind a peer barrier would hold up every other stream's memsets behind it.
int f184_zero_counters(f184_ctx* c, uint32_t mask);                       // zero counters_dev[i] for every bit i of mask, on c->stream
int f184_fill_async(f184_ctx* c, void* ptr, uint32_t value32, size_t bytes, cudaStream_t stream);   // bytes: multiple of 4
int f184_stage_begin(f184_ctx* c, int stage);
int f184_stage_end(f184_ctx* c, int stage);
template <class T> static inline T* img_ptr(f184_ctx* c, int slot) { return reinterpret_cast<T*>(c->img[slot].ptr); }

static inline M4 host_matmul(const M4& A, const M4& B)
{
    // same association as the GLSL `A * B` then `M * v`: C[:,j] = A * B[:,j], each row ((a0*x + a1*y) + a2*z) + a3*w
    M4 C;
    for (int j = 0; j < 4; j++)
        for (int r = 0; r < 4; r++)
        {
            volatile float t0 = A.m[0 + r] * B.m[4 * j + 0];
            volatile float t1 = A.m[4 + r] * B.m[4 * j + 1];
            volatile float t2 = A.m[8 + r] * B.m[4 * j + 2];
            volatile float t3 = A.m[12 + r] * B.m[4 * j + 3];
            volatile float s = t0 + t1;
            s = s + t2;
            s = s + t3;
            C.m[4 * j + r] = s;
        }
    return C;
}

// pass implementations (one per translation unit)
int f184_voxelize_r(f184_ctx* c, const f184_view_constants* cam);
int f184_trace_r(f184_ctx* c, const f184_trace_constants* k);
int f184_gtao_impl(f184_ctx* c, const f184_view_constants* view);
int f184_blur_impl(f184_ctx* c, const f184_engine_miscs* miscs);
int f184_gtao_fast_impl(f184_ctx* c, const f184_view_constants* view);     // secondary_fast.cu: north-star contract (1e-2), hardware units
int f184_blur_fast_impl(f184_ctx* c, const f184_engine_miscs* miscs);
int f184_composite_impl(f184_ctx* c, const f184_trace_constants* k);
int f184_lighting_impl(f184_ctx* c, const f184_view_constants* view, const f184_extended_matrices* m, const f184_light_list* point,
                       const f184_light_list* directional);
int f184_voxelize_n(f184_ctx* c, const f184_view_constants* cam);
int f184_inject_n(f184_ctx* c, const f184_sun* sun, const f184_extended_matrices* m);
int f184_mips_n(f184_ctx* c);
int f184_trace_n(f184_ctx* c, const f184_trace_constants* k);
int f184_trace_views_n(f184_ctx* c, const f184_trace_constants* ks, uint32_t view_h, uint32_t first, uint32_t count);
#define F184_VIEW_STREAMS 4
int f184_mode_n_release(f184_ctx* c);
int f184_mode_n_alloc(f184_ctx* c);
int f184_normalise_n(f184_ctx* c);
int f184_static_cache_capture_n(f184_ctx* c);
int f184_static_cache_clear_n(f184_ctx* c);
int f184_voxelizer_scratch_n(f184_ctx* c);
int f184_voxelize_accumulate_n(f184_ctx* c, const f184_view_constants* cam);
int f184_gather_n(f184_ctx* c, const f184_trace_constants* view);
int f184_gather_init_n(f184_ctx* c);
int f184_mips_init_n(f184_ctx* c);
int f184_trace_init_n(f184_ctx* c);
int f184_ipc_buffer_ptr(f184_ctx* c, uint32_t buffer, void** out);
M4 f184_invert_m4(const M4& A);
// world size of one voxel along voxel-x under the voxel camera (Proj * View), the `h` of the cone tracer
static inline float f184_voxel_h(const float* proj, const float* view, uint32_t N)
{
    M4 vp, vv;
    memcpy(vp.m, proj, 64);
    memcpy(vv.m, view, 64);
    const M4 v2w = f184_invert_m4(host_matmul(vp, vv));
    const float s = 2.0f / (float)N;
    const float ax = v2w.m[0] * s, ay = v2w.m[1] * s, az = v2w.m[2] * s;
    return sqrtf((ax * ax + ay * ay) + az * az);
}
float f184_exposure(const f184_ctx* c, const f184_sun* sun);
// rows/tiles selection of the trace passes: y = y0 + (tile0 + blockIdx.y * stride) * 8 + ...; returns grid.y (0 = nothing to do)
static inline uint32_t f184_trace_tiles(const f184_ctx* c, uint32_t H, uint32_t* y0, uint32_t* y1, uint32_t* tile0, uint32_t* stride)
{
    *y0 = c->row0 < H ? c->row0 : H;
    *y1 = c->row1 < H ? c->row1 : H;
    *stride = c->tile_stride ? c->tile_stride : 1;
    if (*y1 <= *y0) return 0;
    const uint32_t tiles = (*y1 - *y0 + 7) / 8, base = *y0 / 8;
    *tile0 = (c->tile_first % *stride + *stride - base % *stride) % *stride;
    if (*tile0 >= tiles) return 0;
    return (tiles - *tile0 + *stride - 1) / *stride;
}
