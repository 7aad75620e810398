// blur.cu — separable cross-bilateral blur of the indirect buffer (the tail of the trace).
//
// Replaces indirect_blurX + indirect_blurY (Foreground/Renderer/MegaPipeline.cpp:270-284):
// Shader/Lighting/bilateralBlur.inc, first along y, then along x: the pass the reference names indirect_blurX steps
// vertically (blurX.frag:5 defines DIR(x) as vec2(0.0, x)) and indirect_blurY horizontally (blurY.frag:5); bilateral
// weights do not commute, so the shipped order is kept.  13 taps per direction at
// pixel offsets 0, +-2, +-4, +-6, +-8, +-11, +-15 (`invres = 2/resolution`; the outer taps' half-texel
// offsets land on texel centres too), weight exp2(-r^2/32 - ((z0 - z)*512)^2), clamp-to-edge sampler.
#include "f184_device.cuh"

namespace {

struct BlurParams { const uint16_t* src; const float* depth; uint16_t* dst; int W, H, dirx, diry; };

__global__ void __launch_bounds__(128) k_blur(const BlurParams B)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= B.W || y >= B.H) return;
    const float BlurFalloff = 1.0f / (2.0f * 4.0f * 4.0f);
    auto cl = [](int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); };
    float tc0, tc1, tc2, tw = 1.0f, cz;
    {
        const ushort4 p = __ldg(reinterpret_cast<const ushort4*>(B.src) + (size_t)y * B.W + x);
        tc0 = dm_f16_to_f32(p.x) * 1.0f; tc1 = dm_f16_to_f32(p.y) * 1.0f; tc2 = dm_f16_to_f32(p.z) * 1.0f;
        cz = __ldg(B.depth + (size_t)y * B.W + x);
    }
    auto acc = [&](int off, float r) {
        const int xi = cl(x + off * B.dirx, B.W), yi = cl(y + off * B.diry, B.H);
        const ushort4 p = __ldg(reinterpret_cast<const ushort4*>(B.src) + (size_t)yi * B.W + xi);
        const float z = __ldg(B.depth + (size_t)yi * B.W + xi);
        const float dz = (cz - z) * 512.0f;
        const float w = dm_exp2(-r * r * BlurFalloff - dz * dz);
        tc0 += dm_f16_to_f32(p.x) * w; tc1 += dm_f16_to_f32(p.y) * w; tc2 += dm_f16_to_f32(p.z) * w;
        tw += w;
    };
    float i = 1.0f;
    for (; i <= 4.0f; i += 1.0f) { acc((int)(2.0f * i), i); acc(-(int)(2.0f * i), i); }
    for (; i <= 8.0f; i += 2.0f) { acc((int)(2.0f * (i + 0.5f)), i); acc(-(int)(2.0f * (0.5f + i)), i); }
    reinterpret_cast<ushort4*>(B.dst)[(size_t)y * B.W + x] =
        make_ushort4(dm_f32_to_f16(tc0 / tw), dm_f32_to_f16(tc1 / tw), dm_f32_to_f16(tc2 / tw), 0);
}

}  // namespace

int f184_blur_impl(f184_ctx* c, const f184_engine_miscs*)
{
    for (int s : {F184_SLOT_DEPTH, F184_SLOT_INDIRECT_OUT, F184_SLOT_INDIRECT_BLUR_X, F184_SLOT_INDIRECT_FINAL})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    const int W = (int)c->cfg.width, H = (int)c->cfg.height;
    int rc = f184_stage_begin(c, F184_STAGE_BLUR);
    if (rc) return rc;
    dim3 grid((W + 15) / 16, (H + 7) / 8);
    BlurParams bx{img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_OUT), img_ptr<float>(c, F184_SLOT_DEPTH), img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_BLUR_X), W, H, 0, 1};
    k_blur<<<grid, 128, 0, c->stream>>>(bx);
    CK_LAUNCH(c);
    BlurParams by{img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_BLUR_X), img_ptr<float>(c, F184_SLOT_DEPTH), img_ptr<uint16_t>(c, F184_SLOT_INDIRECT_FINAL), W, H, 1, 0};
    k_blur<<<grid, 128, 0, c->stream>>>(by);
    CK_LAUNCH(c);
    return f184_stage_end(c, F184_STAGE_BLUR);
}
