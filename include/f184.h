/* f184.h — C-ABI of libf184: the voxel global-illumination hot path of tobyc11/Final184 on B200.
 *
 * The reference has no plugin/FFI boundary: the path is a run of C++ member calls inside
 * CMegaPipeline::Render() (Foreground/Renderer/MegaPipeline.cpp:148-342).  This header is the boundary
 * a maintainer would cut there — every entry point names the reference call site it replaces — so
 * that the rest of the Vulkan renderer stays as it is (INTEGRATION.md shows the patch).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, POD structs only; no exceptions cross the boundary.
 *   - every call returns 0 (F184_OK) or a negative f184_status; f184_last_error() gives the text.
 *   - one host thread per context.  Pass calls enqueue on the context's CUDA stream and return
 *     immediately; only calls documented as synchronous wait for the device.
 *   - matrices are 16 floats in the order the reference uploads them (tc::Matrix4 transposed before
 *     upload, SceneView.cpp:28-30) = column-major, what GLSL `mat4` reads.
 *   - images are pitch-linear device memory (x fastest); 3D volumes are x, then y, then z.
 *   - there is no CPU fallback: without a CUDA device f184_create fails with F184_ERR_NO_DEVICE.
 *
 * Two contracts live behind the same calls:
 *   F184_MODE_REFERENCE  what the shipped shaders compute: centre-sample voxelization with a last-
 *                        writer-wins RG16UI store, 4x2 stochastic 60-step ray marches, GTAO, blur.
 *   F184_MODE_NORTHSTAR  the design BASELINE.json names: conservative voxelization into sum+count
 *                        accumulators, normalise, light injection, six-direction mips, cone tracing.
 */
#ifndef F184_H
#define F184_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F184_ABI_VERSION 1

typedef struct f184_ctx f184_ctx;

typedef enum f184_status {
    F184_OK = 0,
    F184_ERR_INVALID_ARGUMENT = -1,
    F184_ERR_NO_DEVICE = -2,
    F184_ERR_CUDA = -3,
    F184_ERR_OUT_OF_MEMORY = -4,
    F184_ERR_NOT_READY = -5,       /* a required image/scene has not been provided */
    F184_ERR_UNIMPLEMENTED = -6,
    F184_ERR_PEER_TIMEOUT = -7     /* one NVLink box: a rank did not reach f184_peer_barrier in time (reported by the next synchronous call) */
} f184_status;

typedef enum f184_mode {
    F184_MODE_REFERENCE = 0,
    F184_MODE_NORTHSTAR = 1
} f184_mode;

/* Image slots: the images the reference binds by name through CMaterial::setImageView
 * (MegaPipeline.cpp:227-229, 254-258, 272-273, 280-281) and CVoxelizeRenderer's VoxelDS
 * (VoxelizeRenderer.cpp:98), plus the volumes the north-star stages add. */
typedef enum f184_slot {
    F184_SLOT_DEPTH = 0,            /* GBufferDepth, depth aspect as R32_SFLOAT            W x H  */
    F184_SLOT_NORMALS = 1,          /* GBuffer1, R16G16B16A16_UNORM view-space n*0.5+0.5   W x H  */
    F184_SLOT_ALBEDO = 2,           /* GBuffer0, R8G8B8A8_UNORM (bound by GTAO, unused)    W x H  */
    F184_SLOT_MATERIAL = 3,         /* GBuffer2, R8G8B8A8_UNORM (0, roughness, metallic,0) W x H  */
    F184_SLOT_SHADOW = 4,           /* ShadowDepth as R32_SFLOAT                            S x S  */
    F184_SLOT_VOXELS = 5,           /* VoxelImage, R16G16_UINT (albedo565, normal565)       N^3    */
    F184_SLOT_INDIRECT_OUT = 6,     /* indirectImage, R16G16B16A16_SFLOAT (rgb, -viewZ)     W x H  */
    F184_SLOT_INDIRECT_HISTORY = 7, /* indirectTemporalImage, R16G16B16A16_SFLOAT           W x H  */
    F184_SLOT_AO_RAW = 8,           /* gtao_visibility target, R16G16B16A16_SFLOAT          W x H  */
    F184_SLOT_AO_OUT = 9,           /* gtao_blur target, R16G16B16A16_SFLOAT                W x H  */
    F184_SLOT_INDIRECT_BLUR_X = 10, /* indirect_blurX target, R16G16B16A16_SFLOAT           W x H  */
    F184_SLOT_INDIRECT_FINAL = 11,  /* indirect_blurY target, R16G16B16A16_SFLOAT           W x H  */
    /* north-star volumes (no reference image behind them) */
    F184_SLOT_ACCUM_COLOR = 12,     /* float4 (sum r, sum g, sum b, count), brick-major     N^3    */
    F184_SLOT_ACCUM_NORMAL = 13,    /* float4 (sum nx, sum ny, sum nz, 0), brick-major      N^3    */
    F184_SLOT_VOX_ALBEDO = 14,      /* R8G8B8A8_UNORM mean albedo, a = 255*occupied         N^3    */
    F184_SLOT_VOX_NORMAL = 15,      /* R8G8B8A8_SNORM normalised mean normal                N^3    */
    F184_SLOT_RADIANCE = 16,        /* R8G8B8A8_UNORM premultiplied radiance/exposure, mip 0 N^3   */
    F184_SLOT_MIPS = 17,            /* all levels >= 1, six directions each, one allocation        */
    F184_SLOT_BRICK_FLAGS = 18,     /* u32 per 8^3 brick: touched this frame                (N/8)^3 */
    /* the consumer next to the voxel-GI section */
    F184_SLOT_LIGHTING = 19,        /* lightingImage (lighting_deferred target), R16G16B16A16_SFLOAT  W x H */
    F184_SLOT_TAA_HISTORY = 20,     /* taaImageB: last frame's anti-aliased colour (rgb, -viewZ), RGBA16F  W x H */
    F184_SLOT_TAA_OUT = 21,         /* taaImageA: gtao_color target 1, RGBA16F                             W x H */
    F184_SLOT_COLOR_OUT = 22,       /* gtao_color target 0 (the reference renders it into the swap-chain image; here RGBA16F) */
    F184_SLOT_COUNT = 23
} f184_slot;

typedef enum f184_format {
    F184_FMT_UNDEFINED = 0,
    F184_FMT_R32_SFLOAT = 1,
    F184_FMT_R16G16B16A16_UNORM = 2,
    F184_FMT_R8G8B8A8_UNORM = 3,
    F184_FMT_R16G16B16A16_SFLOAT = 4,
    F184_FMT_R16G16_UINT = 5,
    F184_FMT_R32G32B32A32_SFLOAT = 6,
    F184_FMT_R8G8B8A8_SNORM = 7,
    F184_FMT_R32_UINT = 8
} f184_format;

typedef struct f184_image_desc {
    void* device_ptr;        /* CUDA device pointer (or mapped external memory); NULL = let the context own it */
    uint32_t format;         /* f184_format */
    uint32_t width, height, depth;
    uint32_t row_pitch_bytes;   /* 0 = tightly packed */
    uint64_t size_bytes;     /* filled by f184_image_info */
} f184_image_desc;

/* Creation parameters.  Replaces the literals the reference hard-codes: volume 128
 * (MegaPipeline.cpp:475-476, 486-487; indirect.frag:133,143,146), window 1280x720 (App/Game.cpp:29-31),
 * shadow 2048 (MegaPipeline.cpp:392-394; indirect.frag:164), step 0.2 x 60 (indirect.frag:109,138). */
typedef struct f184_config {
    uint32_t struct_size;    /* sizeof(f184_config), for ABI growth */
    int32_t device;          /* CUDA ordinal */
    uint32_t mode;           /* f184_mode */
    uint32_t grid_n;         /* voxel grid edge; power of two, 32..1024 */
    uint32_t width, height;  /* G-buffer / output resolution */
    uint32_t shadow_res;     /* shadow map edge */
    uint32_t march_steps;    /* 60 */
    float step_size;         /* 0.2 world units */
    /* north-star cone parameters */
    float cone_max_distance; /* 32 m */
    float radiance_exposure; /* RGBA8 radiance = radiance / exposure, clamped; 0 = auto: largest sun luminance component */
    /* sharding over one NVLink box; rank/nranks = 0/1 for a single GPU */
    uint32_t rank, nranks;
    uint32_t flags;          /* f184_flags */
} f184_config;

typedef enum f184_flags {
    F184_FLAG_NONE = 0,
    F184_FLAG_EXTERNAL_RANDS = 1,  /* mode R trace: rands come from a bound buffer (march-only parity) */
    F184_FLAG_NO_TMA = 2,          /* mode N mips: dense, plain one-thread-per-texel kernel for every level (cross-check) */
    F184_FLAG_GATHER_LINEAR = 8,   /* multi-GPU: f184_gather_volume also fills the linear RADIANCE / MIPS slots (tests) */
    F184_FLAG_DENSE_MIPS = 4,      /* mode N mips: dense chain, TMA-staged tiles for the large levels (default is the sparse
                                      brick-list path for levels 1-3 + one fused launch for the rest) */
    F184_FLAG_NO_OVERLAP = 16,     /* mode N: every pass on the pass stream, one texture set.  Default: the frame pipeline — the accumulation of
                                      frame f+2 (internal stream 1), normalise / inject / mips [/ barriers / gather] of frame f+1 (internal stream 2)
                                      and the cone trace of frame f (pass stream) run side by side, ordered by events where data flows, the
                                      texture-side volume double-buffered.  Every call that reads a stage's outputs orders itself after it:
                                      results are identical */
    F184_FLAG_SPEC_APPENDIX_B = 32,/* mode N cone tracer as SURVEY.md Appendix B.5 writes it: mip-LINEAR sampling (level 0 / level 1 blend below
                                      lod 1, the hardware's linear mip filter above) and half-diameter steps.  Default is the amended spec of
                                      DESIGN.md B.5: nearest mip level, one sample per voxel of the sampled level.  bench.py reports the image
                                      difference between the two as `spec_delta` */
    F184_FLAG_EXACT_SECONDARY = 64 /* mode N: f184_gtao / f184_blur_indirect use the reference-faithful kernels (bit-exact against the shader text,
                                      IEEE divisions, pinned transcendentals) instead of the fast ones north_star's 1e-2 tolerance allows */
} f184_flags;

/* CViewConstants, Foreground/SceneGraph/SceneView.h:8-14 = GlobalConstants, Shader/EngineCommon.h:7-13. 208 B. */
typedef struct f184_view_constants {
    float CameraPos[4];
    float ViewMat[16];
    float ProjMat[16];
    float InvProj[16];
} f184_view_constants;

/* ExtendedMatricesConstants, MegaPipeline.cpp:132-139 = ExtendedMatrices, indirect.frag:18-24. 320 B. */
typedef struct f184_extended_matrices {
    float InvModelView[16];
    float ShadowView[16];
    float ShadowProj[16];
    float VoxelView[16];
    float VoxelProj[16];
} f184_extended_matrices;

/* PreviousProjections, MegaPipeline.h:16-20 = prevProj, indirect.frag:31-34. 128 B. */
typedef struct f184_prev_proj {
    float PrevProjection[16];
    float PrevModelView[16];
} f184_prev_proj;

/* PerLightConstants, MegaPipeline.cpp:95-99 = Sun, indirect.frag:26-29 (std140: vec3 @0, vec3 @16). 32 B.
 * `position` holds the light's forward DIRECTION for a directional light (MegaPipeline.cpp:110,120). */
typedef struct f184_sun {
    float luminance[3];
    float _pad0;
    float position[3];
    float _pad1;
} f184_sun;

/* LightLists, MegaPipeline.cpp:101-105 = pointLights / directionalLights, aggregateLights.frag:36-44 (std140: 100 x 32 B, then
 * int numLights; the C++ struct is alignas(16)). 3216 B.  `position` of a directional light is its forward direction. */
typedef struct f184_light_list {
    f184_sun lights[100];
    int32_t numLights;
    int32_t _pad[3];
} f184_light_list;

/* EngineCommonMiscs, MegaPipeline.cpp:141-146 = indirect.frag:36-40. 16 B. */
typedef struct f184_engine_miscs {
    float resolution[2];
    uint32_t frameCount;
    float frameTime;
} f184_engine_miscs;

/* Everything lighting_indirect binds with setStruct (MegaPipeline.cpp:259-266). */
typedef struct f184_trace_constants {
    f184_view_constants view;
    f184_extended_matrices ext;
    f184_prev_proj prev;
    f184_sun sun;
    f184_engine_miscs miscs;
    uint32_t reset_history;   /* 1 on the first frame / after resize (MegaPipeline.cpp:197-204, 92) */
    uint32_t _pad[3];
} f184_trace_constants;

/* Flat scene: replaces the per-primitive vertex/index buffers of glTFSceneImporter.cpp:210-326 and the
 * per-primitive ModelMat of VoxelizeRenderer.cpp:100-106. */
typedef struct f184_scene_desc {
    const float* positions;       /* n_verts x 3, object space */
    const float* normals;         /* n_verts x 3 */
    const float* uvs;             /* n_verts x 2 */
    const uint32_t* indices;      /* n_tris x 3 */
    const uint16_t* tri_material; /* n_tris */
    const uint16_t* tri_model;    /* n_tris: index into model_mats */
    const float* model_mats;      /* n_models x 16, upload order */
    uint32_t n_verts, n_tris, n_models;
} f184_scene_desc;

typedef enum f184_stage_id {
    F184_STAGE_CLEAR = 0,
    F184_STAGE_VOXELIZE = 1,
    F184_STAGE_NORMALISE = 2,
    F184_STAGE_INJECT = 3,
    F184_STAGE_MIPS = 4,
    F184_STAGE_TRACE = 5,
    F184_STAGE_GTAO = 6,
    F184_STAGE_BLUR = 7,
    F184_STAGE_EXCHANGE = 8,       /* multi-GPU: gather of the other ranks' bricks over NVLink */
    F184_STAGE_LIGHTING = 9,       /* f184_lighting_deferred */
    F184_STAGE_COMPOSITE = 10,     /* f184_composite */
    F184_STAGE_BARRIER = 11,       /* multi-GPU: f184_peer_barrier (flag exchange + the wait for the slowest rank) */
    F184_STAGE_APPLY = 12,         /* multi-GPU: the owner fetches and applies the fragments the other ranks rasterised for it (head of f184_normalise) */
    F184_STAGE_NEED = 13,          /* multi-GPU: which levels / bricks do this rank's cones sample? (head of f184_gather_volume) */
    F184_STAGE_TAIL = 14,          /* multi-GPU: the dense levels below level 3, rebuilt locally behind the gather */
    F184_STAGE_COUNT = 15
} f184_stage_id;

typedef enum f184_counter_id {
    F184_COUNTER_FRAGMENTS = 0,     /* voxel fragments stored/accumulated by the last f184_voxelize */
    F184_COUNTER_MARCH_STEPS = 1,   /* cone-samples (march iterations) of the last f184_trace_indirect */
    F184_COUNTER_OCCUPIED = 2,      /* occupied voxels after the last normalise */
    F184_COUNTER_KERNEL_LAUNCHES = 3, /* kernels launched by this context since creation */
    F184_COUNTER_BRICKS = 4,        /* touched 8^3 bricks in the last f184_voxelize (mode N) */
    F184_COUNTER_GATHER_BYTES = 5,  /* bytes the last f184_gather_volume fetched from the other ranks over NVLink */
    F184_COUNTER_COUNT = 6
} f184_counter_id;

/* ---- lifetime: CMegaPipeline ctor + CreateVoxelizePass/CreateScreenPass, MegaPipeline.cpp:27-60, 470-590 */
int f184_abi_version(void);
int f184_create(const f184_config* config, f184_ctx** out_ctx);
void f184_destroy(f184_ctx* ctx);
const char* f184_last_error(const f184_ctx* ctx);   /* ctx may be NULL: error of a failed f184_create */

/* ---- scene: CVoxelizeRenderer::PreparePrimitiveResources (VoxelizeRenderer.cpp:32-67),
 *      CBasicMaterial::Bind (BasicMaterial.cpp:9-39), image upload + mip generation (DeviceVk.cpp:317-484).
 *      Synchronous. */
int f184_scene_upload(f184_ctx* ctx, const f184_scene_desc* scene);
int f184_texture_upload(f184_ctx* ctx, uint32_t tex_id, const uint8_t* rgba8, uint32_t width, uint32_t height);
int f184_material_set(f184_ctx* ctx, uint32_t material_id, const float base_color_factor[4],
                      int32_t base_color_tex, uint32_t use_textures);
/* mip level of an uploaded texture, as built on the device (testing; synchronous) */
int f184_texture_readback(f184_ctx* ctx, uint32_t tex_id, uint32_t level, uint8_t* rgba8, size_t bytes);

/* ---- images: CMaterial::setImageView / BindImageView.  Context-owned by default; bind to share
 *      torch tensors or imported Vulkan memory. */
int f184_bind_image(f184_ctx* ctx, uint32_t slot, const f184_image_desc* desc);
int f184_image_info(f184_ctx* ctx, uint32_t slot, f184_image_desc* out_desc);
int f184_upload_image(f184_ctx* ctx, uint32_t slot, const void* host, size_t bytes);      /* async H2D on the stream */
int f184_readback(f184_ctx* ctx, uint32_t slot, void* host, size_t bytes);                /* synchronous D2H */
/* Asynchronous D2H.  Images up to 256 MiB are snapshotted on the pass stream (device-to-device, two alternating staging
 * buffers) and travel to `pinned_host` on the context's read-back stream, so the next frame's passes do not wait for PCIe;
 * larger slots are copied on the pass stream itself.  The data is on the host after f184_sync, or after
 * f184_readback_wait(ctx, age): age 0 = the most recent asynchronous read-back, 1 = the one before it (a consumer that
 * keeps one frame in flight submits frame f+1, then waits for the image of frame f). */
int f184_readback_async(f184_ctx* ctx, uint32_t slot, void* pinned_host, size_t bytes);
/* Sharded frames (one process per GPU): the same two transfers restricted to the rows this rank traces (f184_set_trace_rows /
 * f184_set_trace_tiles).  `host` addresses the WHOLE W x H image; only the selected rows are read / written, the other rows of
 * the device image (upload) or of the host buffer (read-back) are left as they are.  A rank of N then moves 1/N of the
 * G-buffer bytes over its PCIe link instead of all of them.  Slots that are not W x H images are transferred whole. */
int f184_upload_image_rows(f184_ctx* ctx, uint32_t slot, const void* host, size_t bytes);
int f184_readback_async_rows(f184_ctx* ctx, uint32_t slot, void* pinned_host, size_t bytes);
int f184_readback_wait(f184_ctx* ctx, uint32_t age);

/* Vulkan interop: import an exported VkDeviceMemory (opaque fd) as a slot, and
 * the section wait/signal semaphores (RHI/Private/Vulkan/CommandListVk.h:17-24). */
int f184_import_external_memory_fd(f184_ctx* ctx, uint32_t slot, int fd, uint64_t alloc_size,
                                   uint64_t offset, const f184_image_desc* layout);
int f184_import_semaphores_fd(f184_ctx* ctx, int wait_fd, int signal_fd);
int f184_frame_begin(f184_ctx* ctx);   /* waits the imported semaphore on the stream (no-op without one) */
int f184_frame_end(f184_ctx* ctx);     /* signals the imported semaphore on the stream (no-op without one) */

/* ---- stream plumbing */
int f184_set_stream(f184_ctx* ctx, void* cuda_stream);   /* NULL = context's own stream */
int f184_sync(f184_ctx* ctx);

/* ---- passes */
/* ClearImage(VoxelImage) MegaPipeline.cpp:196 + the voxelization pass :218-223
 * (CVoxelizeRenderer::RenderList, VoxelizeRenderer.cpp:19-30, 81-127; shaders Pipelang/Internal/main.lua:60-75,
 * 83-144, 179-206, 242-275).  Mode R fills F184_SLOT_VOXELS; mode N fills the accumulators and normalises
 * into VOX_ALBEDO / VOX_NORMAL. */
int f184_voxelize(f184_ctx* ctx, const f184_view_constants* voxel_cam);
/* Mode N only: hoists the first-bounce lighting of indirect.frag:157-169 into the volume. */
int f184_inject(f184_ctx* ctx, const f184_sun* sun, const f184_extended_matrices* matrices);
/* Mode N only: six-direction anisotropic mip chain. */
int f184_build_mips(f184_ctx* ctx);
/* lighting_indirect pass, MegaPipeline.cpp:252-268 (Shader/Lighting/indirect.frag). */
int f184_trace_indirect(f184_ctx* ctx, const f184_trace_constants* constants);
/* Probe batch (BASELINE.json configs[4]: many views traced against one volume; the reference has one camera,
 * App/MainBehaviour.cpp:26-32).  The context's W x H images hold H / view_height views stacked top to bottom: view v is
 * rows [v*view_height, (v+1)*view_height) of DEPTH / NORMALS / MATERIAL / INDIRECT_OUT / INDIRECT_HISTORY and is traced
 * exactly as f184_trace_indirect would trace a W x view_height image with constants[v] (own camera, own history rows).
 * Views [first, first+count) are traced; `constants` is indexed by v.  North-star mode only.  Across GPUs a batch is
 * partitioned by whole views (each rank's context holds its own views); no exchange is involved. */
int f184_trace_views(f184_ctx* ctx, const f184_trace_constants* constants, uint32_t view_height, uint32_t first, uint32_t count);
/* gtao_visibility + gtao_blur, MegaPipeline.cpp:225-239 (Shader/GTAO/gtao.frag, blur.frag). */
int f184_gtao(f184_ctx* ctx, const f184_view_constants* view);
/* indirect_blurX + indirect_blurY, MegaPipeline.cpp:270-284 (Shader/Lighting/bilateralBlur.inc). */
int f184_blur_indirect(f184_ctx* ctx, const f184_engine_miscs* miscs);
/* lighting_deferred pass, MegaPipeline.cpp:286-300 (Shader/Lighting/aggregateLights.frag): Cook-Torrance direct lighting of every
 * point and directional light over the G-buffer, directional lights through 12 x 4 bicubic-weighted shadow taps.  Reads DEPTH,
 * NORMALS, MATERIAL, SHADOW; writes F184_SLOT_LIGHTING.  Either list may be NULL (= empty). */
int f184_lighting_deferred(f184_ctx* ctx, const f184_view_constants* view, const f184_extended_matrices* matrices,
                           const f184_light_list* point_lights, const f184_light_list* directional_lights);
/* gtao_color pass, MegaPipeline.cpp:302-319 (Shader/GTAO/color.frag): albedo^2.2 * (ao * indirect + lighting), sky scattering or
 * volumetric light, tonemap, temporal AA and motion blur.  Reads ALBEDO, AO_OUT, DEPTH, LIGHTING, SHADOW, INDIRECT_FINAL and
 * TAA_HISTORY; writes COLOR_OUT and TAA_OUT.  constants->reset_history = 1 clears TAA_HISTORY first (:197-201). */
int f184_composite(f184_ctx* ctx, const f184_trace_constants* constants);
/* CopyImage(taaImageA -> taaImageB), MegaPipeline.cpp:207-210. */
int f184_copy_taa_to_history(f184_ctx* ctx);
/* CopyImage(indirectImage -> indirectTemporalImage), MegaPipeline.cpp:211-214. */
int f184_copy_indirect_to_history(f184_ctx* ctx);

/* mode R trace with F184_FLAG_EXTERNAL_RANDS: 8 x (u, v) floats per pixel, seed-major (parity aid) */
int f184_bind_rands(f184_ctx* ctx, const float* device_rands, size_t count);

/* ---- sharding.  A rank voxelizes triangles [first, first+count) (or a chunk list, below), owns the 8^3 bricks
 * (bx, by, bz) with (bx + by + bz) % nranks == rank — diagonals, so every axis-aligned wall is dealt evenly — and traces rows
 * [y0, y1) / the tile rows below.  Defaults are derived from rank/nranks. */
int f184_set_triangle_range(f184_ctx* ctx, uint32_t first, uint32_t count);
/* Mode N: the triangles this rank voxelizes as a list of 128-triangle chunks (chunk c = triangles [128 c, 128 c + 128)), on top of which
 * the range above still filters.  Chunks let a cost model deal the scene over the ranks in small pieces (largest first, to the least loaded
 * rank): every rank gets the same mix of wall-sized and sub-voxel triangles, which a contiguous cut cannot give (Sponza's primitives run
 * from 5 to 27,796 triangles).  `chunk_ids` is a host array, copied; NULL / 0 returns to "every triangle of the range".  Synchronous. */
#define F184_TRIANGLE_CHUNK 128
int f184_set_triangle_chunks(f184_ctx* ctx, const uint32_t* chunk_ids, uint32_t n_chunks);
int f184_set_trace_rows(f184_ctx* ctx, uint32_t y0, uint32_t y1);
/* Of the rows selected above, trace only the 8-row tile rows t with t % stride == first (t counted from the top of the
 * image; y0 must be a multiple of 8 when stride > 1).  Interleaving tile rows over the ranks balances the trace where
 * contiguous bands do not (sky at the top, floor at the bottom).  Default: first 0, stride 1. */
int f184_set_trace_tiles(f184_ctx* ctx, uint32_t first, uint32_t stride);

/* ---- one NVLink box, one process per GPU (DESIGN.md "Multi-GPU").  The reference is single-GPU
 * (RHI/Private/Vulkan/DeviceVk.cpp:301-304); this is the north-star schedule:
 *   f184_voxelize_accumulate  rank r rasterises ITS triangles; a fragment whose 8^3 brick this rank owns is reduced into the own
 *                             accumulators (red.global.add.v4.f32), any other fragment becomes a 16-byte record in a LOCAL queue
 *                             for the brick's owner (warp-aggregated appends); behind the barrier the owner reads its queues out
 *                             of the senders' memory with coalesced 16-byte loads over NVLink and applies them with local
 *                             reductions, at the head of f184_normalise.  (Small remote writes — 16-byte reductions or stores
 *                             alike — were measured at 4.6 G packets/s per GPU: two GPUs voxelized slower than one.)  The
 *                             reduce-scatter is an all-to-all of fragments; a full queue falls back to the remote reduction
 *                             (system scope)
 *   f184_peer_barrier         device-side flag barrier over peer memory (no host round trip, no NCCL launch)
 *   f184_normalise / f184_inject / f184_build_mips   owner works on its own bricks only; build_mips also writes them into the
 *                             export arrays (level 0 | level 1 | "coarse" = levels 2, 3 + brick index, contiguous over bricks)
 *   f184_peer_barrier
 *   f184_gather_volume[_view] every rank pulls the other ranks' finished bricks over NVLink with bulk copies (cp.async.bulk into
 *                             shared memory: one 16 KB copy moves the coarse blocks of 64 bricks; level 1 and level 0 follow per
 *                             brick, only where this rank's cones sample them) and stores them into its own texture storage,
 *                             then finishes the small levels locally
 *   f184_trace_indirect       the 8-row tile rows t of the screen with t % nranks == rank
 * Buffers are shared between processes with CUDA IPC handles, exchanged by the caller (torch.distributed). */
typedef enum f184_ipc_buffer {
    F184_IPC_ACCUM_COLOR = 0,
    F184_IPC_ACCUM_NORMAL = 1,
    F184_IPC_BRICK_FLAGS = 2,
    F184_IPC_EXPORT = 3,        /* the export arrays written by f184_build_mips (1024 words of room per own brick) */
    F184_IPC_COUNTERS = 4,
    F184_IPC_BRICK_LIST = 5,    /* (not read by peers any more: a brick's index travels in its coarse block) */
    F184_IPC_SYNC = 6,          /* barrier flags */
    F184_IPC_FRAG_QUEUE = 7,    /* fragment records this rank has for the other ranks (one region per destination; the destination reads it) */
    F184_IPC_FRAG_COUNTS = 8,   /* how many records each region holds (this rank's append cursors) */
    F184_IPC_COUNT = 9
} f184_ipc_buffer;
typedef struct f184_ipc_handle { uint8_t opaque[64]; } f184_ipc_handle;
int f184_ipc_export(f184_ctx* ctx, uint32_t buffer, f184_ipc_handle* out_handle);
int f184_ipc_import(f184_ctx* ctx, uint32_t peer_rank, uint32_t buffer, const f184_ipc_handle* handle);
int f184_voxelize_accumulate(f184_ctx* ctx, const f184_view_constants* voxel_cam);   /* f184_voxelize = this + f184_normalise */
int f184_normalise(f184_ctx* ctx);
int f184_peer_barrier(f184_ctx* ctx);
int f184_gather_volume(f184_ctx* ctx);
/* The same, told which view the following f184_trace_indirect will trace (its constants): level 1 of a brick — 1.5 KB of its 1.7 KB —
 * then travels only if a cone of this rank's rows samples it there, i.e. only around the surfaces this rank's pixels show (cones leave
 * level 1 within ~3 voxels of their origin); elsewhere this rank keeps zeros.  Without a view (NULL, or f184_gather_volume) every
 * listed brick travels whole.  The G-buffer bound at the time of the call must be the one the trace will read. */
int f184_gather_volume_view(f184_ctx* ctx, const f184_trace_constants* view);

/* ---- static / dynamic split (mode N).  The reference re-voxelizes the whole scene every frame
 * (MegaPipeline.cpp:196, 218-223) although Sponza never moves.  Usage:
 *     select the STATIC triangles (f184_set_triangle_range / f184_set_triangle_chunks); f184_voxelize_accumulate(voxel_cam);
 *     [several ranks: f184_peer_barrier;]  f184_static_cache_capture(ctx);  [several ranks: f184_peer_barrier again — the capture pulls
 *     the fragments the other ranks hold for this rank, and nobody may start over its queues before everybody has]
 *     every frame: select the DYNAMIC triangles (possibly none: count 0) and run the frame as usual.
 * Capture moves the accumulators of every brick the static geometry touched (this rank's bricks) into a sparse cache, 16 KB per brick,
 * and leaves the accumulators clear; f184_normalise then adds a brick's cached sums to the frame's.  The sums are integer-valued, so every
 * volume, texture level and counter is bit-identical to voxelizing all triangles every frame — without the static share of voxelize
 * (and, on several ranks, of the fragment exchange), and without normalise reading static bricks no dynamic fragment touched.
 * The voxel camera must stay the one the cache was captured with (F184_ERR_INVALID_ARGUMENT otherwise).  Both calls are synchronous. */
int f184_static_cache_capture(f184_ctx* ctx);
int f184_static_cache_clear(f184_ctx* ctx);

/* ---- measurement */
int f184_stage_time_ms(f184_ctx* ctx, uint32_t stage, float* out_ms);   /* last run of the stage; synchronous */
int f184_counter_get(f184_ctx* ctx, uint32_t which, uint64_t* out_value);    /* synchronous */
/* Timing over a region of many frames without a sync inside it: after f184_stage_time_reset(ctx, 1) every
 * run of a stage records its own CUDA event pair on the stream; f184_stage_time_total sums them (synchronous).
 * f184_stage_time_reset(ctx, 0) returns to "last run only". */
int f184_stage_time_reset(f184_ctx* ctx, uint32_t accumulate);
int f184_stage_time_total(f184_ctx* ctx, uint32_t stage, float* out_ms_sum, uint32_t* out_runs);

/* ---- test hook: evaluate csrc/f184_detmath.h on the device (op: 0 sin, 1 cos, 2 log, 3 log2, 4 exp2,
 * 5 pow(x,y), 6 f32->f16->f32, 7 float->int, 8 float->uint, 9 f32->f16 bits, 10 f16 bits->f32; 7-10 return raw bits);
 * host pointers; synchronous */
int f184_debug_detmath(f184_ctx* ctx, uint32_t op, const float* x, const float* y, float* out, size_t n);

/* ---- measurement aid: device peaks MEASURED_PEAKS.json does not hold.  which = 0: trilinear RGBA8 3D texture fetches / s
 * (bound of the cone tracer); which = 1: scattered 16-byte red.global.add.v4.f32 / s over 1 GiB (the voxelizer's
 * accumulation path).  Synchronous; allocates and frees its own scratch. */
int f184_microbench(f184_ctx* ctx, uint32_t which, double* out_per_second);
/* peer reads over NVLink from rank `peer`'s export buffer (imported with f184_ipc_import): mode 0 = bulk copies (cp.async.bulk) of
 * `copy_bytes` issued by one lane per CTA into a ring of `depth` slots, 1 = issued by 32 lanes per CTA, 2 = per-lane 16-byte loads.
 * The numbers behind the gather's design (DESIGN.md "Gather").  Synchronous. */
int f184_microbench_peer(f184_ctx* ctx, uint32_t peer, uint32_t mode, uint32_t copy_bytes, uint32_t depth, uint32_t ctas, uint64_t total_bytes, double* out_gbs);

/* ---- test hook: one level of the texture-side storage the cone tracer samples (dir < 0: the level-0 radiance 3D array, copied out;
 * dir 0..5: level `level`+1 of that direction, as the texture units return it at every texel centre — the tracer's own fetch path);
 * host pointer; synchronous */
int f184_debug_read_array(f184_ctx* ctx, int32_t dir, uint32_t level, void* host, size_t bytes);

/* ---- test hooks: several contexts of ONE process as the ranks of a box ("loopback ranks": cudaIpcOpenMemHandle refuses handles of
 * the own process).  f184_debug_get_ipc_ptr returns the device pointer f184_ipc_export would share; f184_debug_set_peer installs such a
 * pointer as rank `peer_rank`'s buffer in place of f184_ipc_import.  The test-suite drives the multi-rank schedule on a single GPU
 * with them, and a barrier whose peer never arrives (F184_ERR_PEER_TIMEOUT). */
int f184_debug_get_ipc_ptr(f184_ctx* ctx, uint32_t buffer, void** out_device_ptr);
int f184_debug_set_peer(f184_ctx* ctx, uint32_t peer_rank, uint32_t buffer, void* device_ptr);

#ifdef __cplusplus
}
#endif
#endif /* F184_H */
