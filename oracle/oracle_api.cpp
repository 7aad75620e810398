// oracle_api.cpp — TEST INFRASTRUCTURE (see oracle_common.h).  Context, scene, texture and image
// plumbing of the CPU oracle; mirrors include/f184.h with an f184o_ prefix.
#include <chrono>
#include <cstdio>

#include <algorithm>

#include <omp.h>

#include "oracle_common.h"

using namespace orc;

namespace orc {

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
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
    uint32_t W = c.width, H = c.height, S = c.shadow_res, N = c.grid_n;
    auto set = [&](uint32_t f, uint32_t w, uint32_t h, uint32_t dep) {
        d->format = f; d->width = w; d->height = h; d->depth = dep;
        d->row_pitch_bytes = w * fmt_bytes(f);
        d->size_bytes = (uint64_t)w * h * dep * fmt_bytes(f);
    };
    switch (slot)
    {
    case F184_SLOT_DEPTH: set(F184_FMT_R32_SFLOAT, W, H, 1); break;
    case F184_SLOT_NORMALS: set(F184_FMT_R16G16B16A16_UNORM, W, H, 1); break;
    case F184_SLOT_ALBEDO: set(F184_FMT_R8G8B8A8_UNORM, W, H, 1); break;
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

int ensure_image(f184o_ctx* c, int slot)
{
    Image& im = c->img[slot];
    if (im.ptr) return F184_OK;
    f184_image_desc d;
    if (!default_desc(c->cfg, slot, &d)) return F184_ERR_INVALID_ARGUMENT;
    im.own.assign(d.size_bytes, 0);
    im.ptr = im.own.data();
    im.desc = d;
    im.desc.device_ptr = im.ptr;
    return F184_OK;
}

}  // namespace orc

static std::string g_create_err;

extern "C" {

int f184o_abi_version(void) { return F184_ABI_VERSION; }

int f184o_create(const f184_config* config, f184o_ctx** out)
{
    if (!config || !out || config->struct_size != sizeof(f184_config)) { g_create_err = "bad config"; return F184_ERR_INVALID_ARGUMENT; }
    uint32_t n = config->grid_n;
    if (n < 8 || n > 1024 || (n & (n - 1))) { g_create_err = "grid_n must be a power of two in [8,1024]"; return F184_ERR_INVALID_ARGUMENT; }
    auto* c = new f184o_ctx();
    c->cfg = *config;
    if (c->cfg.march_steps == 0) c->cfg.march_steps = 60;
    if (c->cfg.step_size == 0.f) c->cfg.step_size = 0.2f;
    if (c->cfg.shadow_res == 0) c->cfg.shadow_res = 2048;
    if (c->cfg.cone_max_distance == 0.f) c->cfg.cone_max_distance = 32.f;
    if (c->cfg.nranks == 0) c->cfg.nranks = 1;
    *out = c;
    return F184_OK;
}
void f184o_destroy(f184o_ctx* c) { delete c; }
const char* f184o_last_error(const f184o_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int f184o_scene_upload(f184o_ctx* c, const f184_scene_desc* s)
{
    if (!c || !s) return F184_ERR_INVALID_ARGUMENT;
    c->n_verts = s->n_verts; c->n_tris = s->n_tris; c->n_models = s->n_models;
    c->pos.assign(s->positions, s->positions + 3ull * s->n_verts);
    c->nrm.assign(s->normals, s->normals + 3ull * s->n_verts);
    c->uv.assign(s->uvs, s->uvs + 2ull * s->n_verts);
    c->idx.assign(s->indices, s->indices + 3ull * s->n_tris);
    c->tri_mat.assign(s->tri_material, s->tri_material + s->n_tris);
    c->tri_model.assign(s->tri_model, s->tri_model + s->n_tris);
    c->model_mats.assign(s->model_mats, s->model_mats + 16ull * s->n_models);
    for (uint32_t i = 0; i < 3 * s->n_tris; i++)
        if (c->idx[i] >= s->n_verts) { c->err = "index out of range"; return F184_ERR_INVALID_ARGUMENT; }
    return F184_OK;
}

// Mip chain: each level is a linear 2:1 blit of the previous one (RHI/Private/Vulkan/DeviceVk.cpp:437-465),
// i.e. the mean of a 2x2 block, rounded to nearest-even on conversion back to UNORM8.
int f184o_texture_upload(f184o_ctx* c, uint32_t id, const uint8_t* rgba, uint32_t w, uint32_t h)
{
    if (!c || !rgba || !w || !h || (w & (w - 1)) || (h & (h - 1))) return F184_ERR_INVALID_ARGUMENT;
    if (c->textures.size() <= id) c->textures.resize(id + 1);
    Texture& t = c->textures[id];
    t.w = w; t.h = h; t.levels.clear();
    t.levels.emplace_back(rgba, rgba + 4ull * w * h);
    uint32_t nl = 1;
    for (uint32_t m = (w < h ? w : h); m > 1; m >>= 1) nl++;     // 1 + floor(log2(min(w,h))), DeviceVk.cpp:504-509
    uint32_t sw = w, sh = h;
    for (uint32_t l = 1; l < nl; l++)
    {
        uint32_t dw = sw > 1 ? sw / 2 : 1, dh = sh > 1 ? sh / 2 : 1;
        const std::vector<uint8_t>& src = t.levels[l - 1];
        std::vector<uint8_t> dst(4ull * dw * dh);
        for (uint32_t y = 0; y < dh; y++)
            for (uint32_t x = 0; x < dw; x++)
                for (int ch = 0; ch < 4; ch++)
                {
                    uint32_t s = src[4 * ((2 * y) * sw + 2 * x) + ch] + src[4 * ((2 * y) * sw + 2 * x + 1) + ch] +
                                 src[4 * ((2 * y + 1) * sw + 2 * x) + ch] + src[4 * ((2 * y + 1) * sw + 2 * x + 1) + ch];
                    uint32_t q = s >> 2, r = s & 3;
                    if (r > 2 || (r == 2 && (q & 1))) q++;
                    dst[4 * (y * dw + x) + ch] = (uint8_t)q;
                }
        t.levels.push_back(std::move(dst));
        sw = dw; sh = dh;
    }
    return F184_OK;
}

int f184o_texture_readback(f184o_ctx* c, uint32_t id, uint32_t level, uint8_t* out, size_t bytes)
{
    if (!c || id >= c->textures.size() || level >= c->textures[id].levels.size()) return F184_ERR_INVALID_ARGUMENT;
    const auto& l = c->textures[id].levels[level];
    if (bytes != l.size()) return F184_ERR_INVALID_ARGUMENT;
    memcpy(out, l.data(), bytes);
    return F184_OK;
}

int f184o_material_set(f184o_ctx* c, uint32_t id, const float factor[4], int32_t tex, uint32_t use_textures)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    if (c->materials.size() <= id) c->materials.resize(id + 1);
    memcpy(c->materials[id].factor, factor, 16);
    c->materials[id].tex = tex;
    c->materials[id].use_textures = use_textures;
    return F184_OK;
}

int f184o_bind_image(f184o_ctx* c, uint32_t slot, const f184_image_desc* d)
{
    if (!c || slot >= F184_SLOT_COUNT || !d) return F184_ERR_INVALID_ARGUMENT;
    Image& im = c->img[slot];
    im.own.clear();
    im.ptr = d->device_ptr;
    im.desc = *d;
    return F184_OK;
}
int f184o_image_info(f184o_ctx* c, uint32_t slot, f184_image_desc* out)
{
    if (!c || slot >= F184_SLOT_COUNT || !out) return F184_ERR_INVALID_ARGUMENT;
    int r = ensure_image(c, slot);
    if (r) return r;
    *out = c->img[slot].desc;
    return F184_OK;
}
int f184o_upload_image(f184o_ctx* c, uint32_t slot, const void* host, size_t bytes)
{
    if (!c || slot >= F184_SLOT_COUNT) return F184_ERR_INVALID_ARGUMENT;
    int r = ensure_image(c, slot);
    if (r) return r;
    if (bytes != c->img[slot].desc.size_bytes) { c->err = "size mismatch"; return F184_ERR_INVALID_ARGUMENT; }
    memcpy(c->img[slot].ptr, host, bytes);
    return F184_OK;
}
// the rows this rank traces (set_trace_rows / set_trace_tiles) of a W x H image; other shapes whole
static int copy_selected_rows(f184o_ctx* c, uint32_t slot, void* dst, const void* src, size_t bytes)
{
    if (!c || slot >= F184_SLOT_COUNT) return F184_ERR_INVALID_ARGUMENT;
    int r = ensure_image(c, slot);
    if (r) return r;
    const f184_image_desc& d = c->img[slot].desc;
    if (bytes != d.size_bytes) { c->err = "size mismatch"; return F184_ERR_INVALID_ARGUMENT; }
    if (!dst) dst = c->img[slot].ptr;
    if (!src) src = c->img[slot].ptr;
    if (d.height != c->cfg.height || d.depth != 1) { memcpy(dst, src, bytes); return F184_OK; }
    const size_t rb = d.row_pitch_bytes;
    for (uint32_t y = c->row0; y < std::min(c->row1, d.height); y++)
        if ((y >> 3) % c->tile_stride == c->tile_first) memcpy((char*)dst + y * rb, (const char*)src + y * rb, rb);
    return F184_OK;
}
int f184o_upload_image_rows(f184o_ctx* c, uint32_t slot, const void* host, size_t bytes) { return copy_selected_rows(c, slot, nullptr, host, bytes); }
int f184o_readback_async_rows(f184o_ctx* c, uint32_t slot, void* host, size_t bytes) { return copy_selected_rows(c, slot, host, nullptr, bytes); }
int f184o_readback(f184o_ctx* c, uint32_t slot, void* host, size_t bytes)
{
    if (!c || slot >= F184_SLOT_COUNT) return F184_ERR_INVALID_ARGUMENT;
    int r = ensure_image(c, slot);
    if (r) return r;
    if (bytes != c->img[slot].desc.size_bytes) { c->err = "size mismatch"; return F184_ERR_INVALID_ARGUMENT; }
    memcpy(host, c->img[slot].ptr, bytes);
    return F184_OK;
}
int f184o_readback_async(f184o_ctx* c, uint32_t slot, void* host, size_t bytes) { return f184o_readback(c, slot, host, bytes); }
int f184o_set_stream(f184o_ctx*, void*) { return F184_OK; }
int f184o_readback_wait(f184o_ctx*, uint32_t) { return F184_OK; }
int f184o_sync(f184o_ctx*) { return F184_OK; }
int f184o_frame_begin(f184o_ctx*) { return F184_OK; }
int f184o_frame_end(f184o_ctx*) { return F184_OK; }
int f184o_bind_rands(f184o_ctx* c, const float* r, size_t n) { c->rands = r; c->n_rands = n; return F184_OK; }
// test hook: the mode R voxel pass runs these programmable stages (the reference's own shader text, compiled into
// oracle/_ref/libf184_refshaders.so) instead of the restated ones; NULL restores the restatement
int f184o_debug_set_voxel_stage_hooks(f184o_ctx* c, void* gs, void* ps)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    c->gs_hook = (f184o_voxel_gs_hook)gs; c->ps_hook = (f184o_voxel_ps_hook)ps;
    return F184_OK;
}
int f184o_set_triangle_range(f184o_ctx* c, uint32_t first, uint32_t count) { c->tri_first = first; c->tri_count = count; return F184_OK; }
int f184o_set_triangle_chunks(f184o_ctx* c, const uint32_t* chunk_ids, uint32_t n_chunks)
{
    if (!c || (n_chunks && !chunk_ids)) return F184_ERR_INVALID_ARGUMENT;
    c->chunk_mask.clear();
    if (!n_chunks) return F184_OK;
    uint32_t mx = 0;
    for (uint32_t i = 0; i < n_chunks; i++) mx = std::max(mx, chunk_ids[i]);
    c->chunk_mask.assign((size_t)mx + 1, 0);
    for (uint32_t i = 0; i < n_chunks; i++) c->chunk_mask[chunk_ids[i]] = 1;
    return F184_OK;
}
int f184o_set_trace_rows(f184o_ctx* c, uint32_t y0, uint32_t y1) { c->row0 = y0; c->row1 = y1; return F184_OK; }
int f184o_set_trace_tiles(f184o_ctx* c, uint32_t first, uint32_t stride)
{
    if (!c || stride == 0 || first >= stride) return F184_ERR_INVALID_ARGUMENT;
    c->tile_first = first; c->tile_stride = stride;
    return F184_OK;
}
int f184o_stage_time_ms(f184o_ctx* c, uint32_t stage, float* ms)
{
    if (!c || stage >= F184_STAGE_COUNT || !ms) return F184_ERR_INVALID_ARGUMENT;
    *ms = c->stage_ms[stage];
    return F184_OK;
}
int f184o_counter_get(f184o_ctx* c, uint32_t which, uint64_t* v)
{
    if (!c || which >= F184_COUNTER_COUNT || !v) return F184_ERR_INVALID_ARGUMENT;
    *v = c->counters[which];
    if (which == F184_COUNTER_FRAGMENTS && c->cache_held) *v += c->cache_fragments;
    return F184_OK;
}
int f184o_copy_taa_to_history(f184o_ctx* c)
{
    int r = ensure_image(c, F184_SLOT_TAA_OUT); if (r) return r;
    r = ensure_image(c, F184_SLOT_TAA_HISTORY); if (r) return r;
    memcpy(c->img[F184_SLOT_TAA_HISTORY].ptr, c->img[F184_SLOT_TAA_OUT].ptr, c->img[F184_SLOT_TAA_OUT].desc.size_bytes);
    return F184_OK;
}
int f184o_copy_indirect_to_history(f184o_ctx* c)
{
    int r = ensure_image(c, F184_SLOT_INDIRECT_OUT); if (r) return r;
    r = ensure_image(c, F184_SLOT_INDIRECT_HISTORY); if (r) return r;
    memcpy(c->img[F184_SLOT_INDIRECT_HISTORY].ptr, c->img[F184_SLOT_INDIRECT_OUT].ptr, c->img[F184_SLOT_INDIRECT_OUT].desc.size_bytes);
    return F184_OK;
}

// mode dispatch (implemented in oracle_mode_r.cpp / oracle_mode_n.cpp)
int orc_voxelize_r(f184o_ctx*, const f184_view_constants*);
int orc_trace_r(f184o_ctx*, const f184_trace_constants*);
int orc_voxelize_n(f184o_ctx*, const f184_view_constants*);
int orc_voxelize_accumulate_n(f184o_ctx*, const f184_view_constants*);
int orc_normalise_n(f184o_ctx*);
int orc_inject_n(f184o_ctx*, const f184_sun*, const f184_extended_matrices*);
int orc_mips_n(f184o_ctx*);
int orc_trace_n(f184o_ctx*, const f184_trace_constants*);
int orc_trace_views_n(f184o_ctx*, const f184_trace_constants*, uint32_t, uint32_t, uint32_t);

// same table checks as libf184 (f184_api.cu check_tables): the voxelizers index materials[tri_mat] and textures[mat.tex] unguarded
static int check_tables(f184o_ctx* c)
{
    uint32_t max_mat = 0;
    for (uint16_t m : c->tri_mat) if (m > max_mat) max_mat = m;
    if (max_mat >= c->materials.size()) { c->err = "voxelize: a triangle's material has no material_set entry"; return F184_ERR_INVALID_ARGUMENT; }
    for (const auto& m : c->materials)
        if (m.use_textures && m.tex >= 0 && ((size_t)m.tex >= c->textures.size() || c->textures[m.tex].levels.empty()))
        { c->err = "voxelize: a material uses a texture that has not been uploaded"; return F184_ERR_NOT_READY; }
    return F184_OK;
}
int f184o_voxelize(f184o_ctx* c, const f184_view_constants* cam)
{
    if (!c || !cam) return F184_ERR_INVALID_ARGUMENT;
    if (!c->n_tris) { c->err = "no scene"; return F184_ERR_NOT_READY; }
    if (int rc = check_tables(c)) return rc;
    return c->cfg.mode == F184_MODE_REFERENCE ? orc_voxelize_r(c, cam) : orc_voxelize_n(c, cam);
}
// threads the oracle's parallel loops actually run on (bench.py prints it beside the CPU baseline); set > 0 first asks for that many
int f184o_omp_threads(int set)
{
    if (set > 0) omp_set_num_threads(set);
    int n = 1;
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
    return n;
}
int f184o_voxelize_accumulate(f184o_ctx* c, const f184_view_constants* cam)
{
    if (!c || !cam || c->cfg.mode != F184_MODE_NORTHSTAR) return F184_ERR_INVALID_ARGUMENT;
    if (int rc = check_tables(c)) return rc;
    return orc_voxelize_accumulate_n(c, cam);
}
int f184o_normalise(f184o_ctx* c)
{
    if (!c || c->cfg.mode != F184_MODE_NORTHSTAR) return F184_ERR_INVALID_ARGUMENT;
    return orc_normalise_n(c);
}
extern "C" int orc_static_cache_capture_n(f184o_ctx*);
int f184o_static_cache_capture(f184o_ctx* c)
{
    if (!c || c->cfg.mode != F184_MODE_NORTHSTAR) return F184_ERR_INVALID_ARGUMENT;
    return orc_static_cache_capture_n(c);
}
int f184o_static_cache_clear(f184o_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    c->cacheC.clear(); c->cacheC.shrink_to_fit(); c->cacheN.clear(); c->cacheN.shrink_to_fit();
    c->cache_held = false; c->cache_fragments = 0;
    return F184_OK;
}
int f184o_inject(f184o_ctx* c, const f184_sun* sun, const f184_extended_matrices* m)
{
    if (!c || !sun || !m) return F184_ERR_INVALID_ARGUMENT;
    if (c->cfg.mode != F184_MODE_NORTHSTAR) { c->err = "inject is a north-star stage"; return F184_ERR_INVALID_ARGUMENT; }
    return orc_inject_n(c, sun, m);
}
int f184o_build_mips(f184o_ctx* c)
{
    if (!c) return F184_ERR_INVALID_ARGUMENT;
    if (c->cfg.mode != F184_MODE_NORTHSTAR) { c->err = "mips are a north-star stage"; return F184_ERR_INVALID_ARGUMENT; }
    return orc_mips_n(c);
}
int f184o_trace_indirect(f184o_ctx* c, const f184_trace_constants* k)
{
    if (!c || !k) return F184_ERR_INVALID_ARGUMENT;
    return c->cfg.mode == F184_MODE_REFERENCE ? orc_trace_r(c, k) : orc_trace_n(c, k);
}
int f184o_trace_views(f184o_ctx* c, const f184_trace_constants* ks, uint32_t view_height, uint32_t first, uint32_t count)
{
    if (!c || !ks) return F184_ERR_INVALID_ARGUMENT;
    if (c->cfg.mode != F184_MODE_NORTHSTAR) { c->err = "trace_views: north-star mode only"; return F184_ERR_UNIMPLEMENTED; }
    return orc_trace_views_n(c, ks, view_height, first, count);
}

}  // extern "C"

// ---- test hook: evaluate f184_detmath.h on the host (tests/test_detmath.py) ----------------------
extern "C" int f184o_debug_detmath(uint32_t op, const float* x, const float* y, float* out, size_t n)
{
    for (size_t i = 0; i < n; i++)
    {
        switch (op)
        {
        case 0: out[i] = dm_sin(x[i]); break;
        case 1: out[i] = dm_cos(x[i]); break;
        case 2: out[i] = dm_log(x[i]); break;
        case 3: out[i] = dm_log2(x[i]); break;
        case 4: out[i] = dm_exp2(x[i]); break;
        case 5: out[i] = dm_pow(x[i], y[i]); break;
        case 6: out[i] = dm_f16_to_f32(dm_f32_to_f16(x[i])); break;
        case 7: out[i] = dm_u2f((uint32_t)dm_f2i(x[i])); break;                    // conversions: results returned as raw bits
        case 8: out[i] = dm_u2f(dm_f2uint(x[i])); break;
        case 9: out[i] = dm_u2f((uint32_t)dm_f32_to_f16(x[i])); break;
        case 10: out[i] = dm_f16_to_f32((uint16_t)(dm_f2u(x[i]) & 0xffffu)); break;
        default: return F184_ERR_INVALID_ARGUMENT;
        }
    }
    return F184_OK;
}
