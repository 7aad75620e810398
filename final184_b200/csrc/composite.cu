// composite.cu — the composite pass that consumes the voxel-GI section's outputs.
//
// Replaces gtao_color (Foreground/Renderer/MegaPipeline.cpp:302-319; Shader/GTAO/color.frag): albedo^2.2 * (ao * indirect +
// lighting), single-scattering sky or 16-step volumetric light, tonemap, temporal AA (-> TAA_OUT) and 8-tap motion blur
// (-> COLOR_OUT).  The per-pixel arithmetic lives in f184_composite.h (shared with the CPU oracle, pinned bit for bit to the
// reference's shader text); this file is the launch.  Inputs are one read each of six full-screen images (28 B/pixel) plus
// 9 bilinear history taps and 16 shadow taps with screen-space locality; the pixel is ALU-bound (dm_pow x 6, 16 normalisations).
#include "f184_device.cuh"
#include "f184_composite.h"

namespace {

__global__ void __launch_bounds__(128) k_composite(const f184_composite_in I, uint16_t* __restrict__ out_color, uint16_t* __restrict__ out_taa)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= I.W || y >= I.H) return;
    uint16_t c[4], t[4];
    f184_composite_pixel(I, x, y, c, t);
    const size_t o = (size_t)y * I.W + x;
    reinterpret_cast<ushort4*>(out_color)[o] = make_ushort4(c[0], c[1], c[2], c[3]);
    reinterpret_cast<ushort4*>(out_taa)[o] = make_ushort4(t[0], t[1], t[2], t[3]);
}

}  // namespace

int f184_composite_impl(f184_ctx* c, const f184_trace_constants* k)
{
    for (int s : {F184_SLOT_ALBEDO, F184_SLOT_AO_OUT, F184_SLOT_DEPTH, F184_SLOT_LIGHTING, F184_SLOT_SHADOW, F184_SLOT_INDIRECT_FINAL,
                  F184_SLOT_TAA_HISTORY, F184_SLOT_TAA_OUT, F184_SLOT_COLOR_OUT})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    f184_composite_in I{};
    memcpy(I.InvProj.m, k->view.InvProj, 64); memcpy(I.InvModelView.m, k->ext.InvModelView, 64);
    memcpy(I.ShadowView.m, k->ext.ShadowView, 64); memcpy(I.ShadowProj.m, k->ext.ShadowProj, 64);
    memcpy(I.prevModelView.m, k->prev.PrevModelView, 64); memcpy(I.prevProjection.m, k->prev.PrevProjection, 64);
    memcpy(I.sun_luminance, k->sun.luminance, 12); memcpy(I.sun_position, k->sun.position, 12);
    I.albedo = img_ptr<uint8_t>(c, F184_SLOT_ALBEDO); I.ao = img_ptr<uint16_t>(c, F184_SLOT_AO_OUT);
    I.lighting = img_ptr<uint16_t>(c, F184_SLOT_LIGHTING); I.indirect = img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_FINAL);
    I.taa = img_ptr<uint16_t>(c, F184_SLOT_TAA_HISTORY); I.depth = img_ptr<float>(c, F184_SLOT_DEPTH); I.shadow = img_ptr<float>(c, F184_SLOT_SHADOW);
    I.W = c->cfg.width; I.H = c->cfg.height; I.S = c->cfg.shadow_res;
    int rc = f184_stage_begin(c, F184_STAGE_COMPOSITE);
    if (rc) return rc;
    if (k->reset_history)       // first frame: taaImageA / taaImageB are cleared, MegaPipeline.cpp:197-201
        CK(c, cudaMemsetAsync(c->img[F184_SLOT_TAA_HISTORY].ptr, 0, c->img[F184_SLOT_TAA_HISTORY].desc.size_bytes, c->stream));
    k_composite<<<dim3((I.W + 15) / 16, (I.H + 7) / 8), 128, 0, c->stream>>>(I, img_ptr<uint16_t>(c, F184_SLOT_COLOR_OUT), img_ptr<uint16_t>(c, F184_SLOT_TAA_OUT));
    CK_LAUNCH(c);
    return f184_stage_end(c, F184_STAGE_COMPOSITE);
}
