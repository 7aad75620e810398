// f184_composite.h — the composite pass (gtao_color: Shader/GTAO/color.frag, MegaPipeline.cpp:302-319) as ONE per-pixel
// function compiled for both sides: the CUDA kernel (composite.cu) and the CPU oracle (oracle_mode_r.cpp) include this
// header, the way both include f184_detmath.h.  What holds it to the reference is not the oracle but the reference's own
// shader text compiled by g++ (oracle/_ref/libf184_refshaders.so -> tests/golden/refshader_*.npz): every operation below is
// in the order that text evaluates it, nothing is contracted (-fmad=false / -ffp-contract=off).
//
// color.frag: albedo^2.2 * (ao * indirect + lighting); sky pixels (|wpos| > 256) get the single-scattering atmosphere
// (:76-167), the rest 16 steps of volumetric light through the shadow map (:169-206); ACES-style tonemap + gamma (:45-56);
// temporal AA against taaBuffer (:237-254) -> outTAA; 8-tap motion blur along the reprojection velocity (:256-264) -> outColor.
#pragma once
#include <stdint.h>

#include "f184_detmath.h"

struct cm4 { float m[16]; };        // upload order: m[c*4 + r]
struct cv2 { float x, y; };
struct cv3 { float x, y, z; };
struct cv4 { float x, y, z, w; };

struct f184_composite_in
{
    cm4 InvProj, InvModelView, ShadowView, ShadowProj, prevModelView, prevProjection;
    float sun_luminance[3], sun_position[3];
    const uint8_t* albedo;       // RGBA8
    const uint16_t* ao;          // RGBA16F, .r
    const uint16_t* lighting;    // RGBA16F
    const uint16_t* indirect;    // RGBA16F
    const uint16_t* taa;         // RGBA16F history (rgb, -viewZ)
    const float* depth;
    const float* shadow;
    uint32_t W, H, S;
};

DM_HD cv4 cm_mul(const cm4& M, cv4 v)
{
    cv4 r;
    r.x = ((M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z) + M.m[12] * v.w;
    r.y = ((M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z) + M.m[13] * v.w;
    r.z = ((M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z) + M.m[14] * v.w;
    r.w = ((M.m[3] * v.x + M.m[7] * v.y) + M.m[11] * v.z) + M.m[15] * v.w;
    return r;
}
DM_HD float cm_dot(cv3 a, cv3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
DM_HD float cm_length(cv3 a) { return DM_SQRTF(cm_dot(a, a)); }
DM_HD cv3 cm_normalize(cv3 a) { const float l = cm_length(a); return cv3{a.x / l, a.y / l, a.z / l}; }
// PINNED (glsl_shim.h): exp(x) = exp2(x * log2 e)
DM_HD float cm_exp(float x) { return dm_exp2(x * 1.4426950408889634f); }

// ---- atmosphere, color.frag:76-167 ----------------------------------------------------------------------------------
#define CM_R0 6370e3f
#define CM_RA 6425e3f
#define CM_HR 10e3f
#define CM_HM 2.7e3f

DM_HD void cm_densities(cv3 pos, float* des_x, float* des_y)
{
    // C = (0, -R0, 0): pos - C
    const cv3 d = {pos.x - 0.0f, pos.y - (-CM_R0), pos.z - 0.0f};
    const float h = cm_length(d) - CM_R0;
    float dx = cm_exp(-h / CM_HR);
    dx += cm_exp(-dm_max(0.0f, (h - 35e3f)) / 5e3f) * cm_exp(-dm_max(0.0f, (35e3f - h)) / 15e3f) * 0.2f;
    *des_x = dx;
    *des_y = cm_exp(-h / CM_HM);
}
DM_HD float cm_escape(cv3 p, cv3 d, float R)
{
    const cv3 v = {p.x - 0.0f, p.y - (-CM_R0), p.z - 0.0f};
    const float b = cm_dot(v, d);
    const float c = cm_dot(v, v) - R * R;
    const float det2 = b * b - c;
    if (det2 < 0.0f) return -1.0f;
    const float det = DM_SQRTF(det2);
    const float t1 = -b - det, t2 = -b + det;
    return (t1 >= 0.0f) ? t1 : t2;
}
DM_HD cv3 cm_scatter(cv3 o, cv3 d, cv3 Ds, float l)
{
    const float g = 0.76f;
    const float g2 = g * g;
    const float bR[3] = {5.8e-6f, 13.5e-6f, 33.1e-6f};
    const float bM = 31e-6f;
    if (d.y < 0.0f) d.y = 0.0016f / (-d.y + 0.04f) - 0.04f;
    const float L = dm_min(l, cm_escape(o, d, CM_RA));
    const float mu = cm_dot(d, Ds);
    const float opmu2 = 1.0f + mu * mu;
    const float phaseR = 0.0596831f * opmu2;
    float phaseM = 0.1193662f * (1.0f - g2) * opmu2;
    phaseM /= ((2.0f + g2) * dm_pow(1.0f + g2 - 2.0f * g * mu, 1.5f));
    float depth_x = 0.0f, depth_y = 0.0f;
    float R[3] = {0.0f, 0.0f, 0.0f}, M[3] = {0.0f, 0.0f, 0.0f};
    const float u0 = -(L - 100.0f) / (1.0f - dm_exp2((float)5));
    const float dither = 0.0f;
    for (int i = 0; i < 5; ++i)
    {
        const float dl = u0 * dm_exp2((float)i - dither);
        const float ll = -u0 * (1.0f - dm_exp2((float)i - dither + 1.0f));
        const cv3 p = {o.x + d.x * ll, o.y + d.y * ll, o.z + d.z * ll};
        float des_x, des_y;
        cm_densities(p, &des_x, &des_y);
        des_x *= dl; des_y *= dl;
        depth_x += des_x; depth_y += des_y;
        const float Ls = cm_escape(p, Ds, CM_RA);
        if (Ls > 0.0f)
        {
            float in_x = 0.0f, in_y = 0.0f;
            for (int j = 0; j < 3; ++j)
            {
                const float ls = (float)j / (float)3 * Ls;
                const cv3 ps = {p.x + Ds.x * ls, p.y + Ds.y * ls, p.z + Ds.z * ls};
                float ix, iy;
                cm_densities(ps, &ix, &iy);
                in_x += ix; in_y += iy;
            }
            in_x *= Ls / (float)3; in_y *= Ls / (float)3;
            in_x += depth_x; in_y += depth_y;
            for (int c = 0; c < 3; c++)
            {
                const float A = cm_exp(-(bR[c] * in_x + bM * in_y));
                R[c] += A * des_x;
                M[c] += A * des_y;
            }
        }
        else
            return cv3{0.0f, 0.0f, 0.0f};
    }
    float col[3];
    for (int c = 0; c < 3; c++) col[c] = dm_max(0.0f, 20.0f * (R[c] * bR[c] * phaseR + M[c] * bM * phaseM));
    return cv3{col[0], col[1], col[2]};
}

// ---- fetches ----------------------------------------------------------------------------------------------------------
DM_HD float cm_depth_fetch(const f184_composite_in& I, float u, float v)
{
    const int dx = dm_f2i(u * (float)I.W), dy = dm_f2i(v * (float)I.H);
    return (dx >= 0 && dy >= 0 && dx < (int)I.W && dy < (int)I.H) ? I.depth[(size_t)dy * I.W + dx] : 0.0f;
}
DM_HD int cm_clampi(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }
// texture(taaBuffer, uv): bilinear, clamp to edge (GlobalLinearSamplerClamped, MegaPipeline.cpp:303), fp32 weights
DM_HD cv4 cm_taa_bilinear(const f184_composite_in& I, float u, float v)
{
    const float x = u * (float)I.W - 0.5f, y = v * (float)I.H - 0.5f;
    const float x0f = DM_FLOORF(x), y0f = DM_FLOORF(y);
    const float fx = x - x0f, fy = y - y0f;
    const int xi = dm_f2i(x0f), yi = dm_f2i(y0f);
    const int xa = cm_clampi(xi, (int)I.W), xb = cm_clampi(xi + 1, (int)I.W), ya = cm_clampi(yi, (int)I.H), yb = cm_clampi(yi + 1, (int)I.H);
    const uint16_t* a = I.taa + 4 * ((size_t)ya * I.W + xa);
    const uint16_t* b = I.taa + 4 * ((size_t)ya * I.W + xb);
    const uint16_t* c = I.taa + 4 * ((size_t)yb * I.W + xa);
    const uint16_t* d = I.taa + 4 * ((size_t)yb * I.W + xb);
    float r[4];
    for (int ch = 0; ch < 4; ch++)
        r[ch] = (dm_f16_to_f32(a[ch]) * (1.0f - fx) + dm_f16_to_f32(b[ch]) * fx) * (1.0f - fy) +
                (dm_f16_to_f32(c[ch]) * (1.0f - fx) + dm_f16_to_f32(d[ch]) * fx) * fy;
    return cv4{r[0], r[1], r[2], r[3]};
}
// rand21, Shader/math.inc:9-11
DM_HD float cm_rand21(float x, float y) { return dm_fract(dm_sin(x * 12.9898f + y * 78.233f) * 43758.5453f); }

// ---- one pixel ----------------------------------------------------------------------------------------------------------
DM_HD void f184_composite_pixel(const f184_composite_in& I, uint32_t x, uint32_t y, uint16_t* out_color, uint16_t* out_taa)
{
    const size_t o = (size_t)y * I.W + x;
    const float u = ((float)x + 0.5f) / (float)I.W, v = ((float)y + 0.5f) / (float)I.H;
    // PINNED: texture() at a texel centre returns that texel (albedo, ao, lighting, indirect)
    float albedo[3], color[3];
    for (int c = 0; c < 3; c++) albedo[c] = dm_pow((float)I.albedo[4 * o + c] / 255.0f, 2.2f);
    const float depth = cm_depth_fetch(I, u, v);
    const cv4 cp = cm_mul(I.InvProj, cv4{u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f});
    const cv3 cspos = {cp.x / cp.w, cp.y / cp.w, cp.z / cp.w};
    const cv4 wp4 = cm_mul(I.InvModelView, cv4{cspos.x, cspos.y, cspos.z, 1.0f});
    const cv3 wpos = {wp4.x, wp4.y, wp4.z};
    const cv4 wc4 = cm_mul(I.InvModelView, cv4{0.0f, 0.0f, 0.0f, 1.0f});
    const cv3 wpos_cam = {wc4.x, wc4.y, wc4.z};
    const float ao = dm_f16_to_f32(I.ao[4 * o]);
    for (int c = 0; c < 3; c++)
        color[c] = albedo[c] * (ao * dm_f16_to_f32(I.indirect[4 * o + c]) + dm_f16_to_f32(I.lighting[4 * o + c]));
    const cv3 sunp = {I.sun_position[0], I.sun_position[1], I.sun_position[2]};
    if (cm_length(wpos) > 256.0f)
    {   // sky
        const cv3 ns = cm_normalize(sunp);
        const cv3 s = cm_scatter(cv3{0.0f, 1e3f, 0.0f}, cm_normalize(wpos), cv3{-ns.x, -ns.y, -ns.z}, CM_RA);
        color[0] = s.x; color[1] = s.y; color[2] = s.z;
    }
    else
    {   // volumetric light, color.frag:179-206
        const cv4 sp4 = cm_mul(I.ShadowView, cv4{wpos.x, wpos.y, wpos.z, 1.0f});
        const cv4 sc4 = cm_mul(I.ShadowView, cv4{wpos_cam.x, wpos_cam.y, wpos_cam.z, 1.0f});
        const cv3 spos = {sp4.x, sp4.y, sp4.z}, spos_cam = {sc4.x, sc4.y, sc4.z};
        const float dither = cm_rand21(u, v);
        const cv3 delta = {(spos.x - spos_cam.x) / (float)16, (spos.y - spos_cam.y) / (float)16, (spos.z - spos_cam.z) / (float)16};
        cv3 sp = {spos_cam.x + delta.x * dither, spos_cam.y + delta.y * dither, spos_cam.z + delta.z * dither};
        float contribute = 0.0f;
        const float g2 = 0.76f * 0.76f;
        for (int i = 0; i < 16; i++)
        {
            sp = cv3{sp.x + delta.x, sp.y + delta.y, sp.z + delta.z};
            cv4 pr = cm_mul(I.ShadowProj, cv4{sp.x, sp.y, sp.z, 1.0f});
            pr = cv4{pr.x / pr.w, pr.y / pr.w, pr.z / pr.w, pr.w / pr.w};
            pr.x = pr.x * 0.5f + 0.5f; pr.y = pr.y * 0.5f + 0.5f;
            const int tx = dm_f2i(pr.x * (float)I.S), ty = dm_f2i(pr.y * (float)I.S);
            const float shadowZ = (tx >= 0 && ty >= 0 && tx < (int)I.S && ty < (int)I.S) ? I.shadow[(size_t)ty * I.S + tx] : 0.0f;
            const float shade = dm_step(pr.z, shadowZ);
            // miePhase(sample_pos - spos_cam), :169-176
            const float mu = cm_normalize(cv3{sp.x - spos_cam.x, sp.y - spos_cam.y, sp.z - spos_cam.z}).z;
            const float opmu2 = 1.0f + mu * mu;
            const float phaseM = 0.1193662f * (1.0f - g2) * opmu2;
            contribute += shade * phaseM;
        }
        contribute *= 0.2f / (float)16;
        for (int c = 0; c < 3; c++) color[c] = color[c] * (1.0f - contribute * 0.5f) + I.sun_luminance[c] * contribute;
    }
    // tonemap(color, 1.0), :45-56
    for (int c = 0; c < 3; c++)
    {
        float cl = color[c];
        cl *= 1.0f;
        cl = (cl * (2.51f * cl + 0.03f)) / (cl * (2.43f * cl + 0.59f) + 0.14f);
        color[c] = dm_pow(cl, 1.0f / 2.2f);
    }
    // temporal AA, :237-254
    const cv4 pc = cm_mul(I.prevModelView, cv4{wpos.x, wpos.y, wpos.z, 1.0f});
    cv4 pp = cm_mul(I.prevProjection, pc);
    pp = cv4{pp.x / pp.w, pp.y / pp.w, pp.z / pp.w, pp.w / pp.w};
    const float ru = pp.x * 0.5f + 0.5f, rv = pp.y * 0.5f + 0.5f;
    const float velx = ru - u, vely = rv - v;
    if (dm_clamp(ru, 0.0f, 1.0f) == ru && dm_clamp(rv, 0.0f, 1.0f) == rv)
    {
        float bw = 0.6f;
        const cv4 prev = cm_taa_bilinear(I, ru, rv);
        const float vlen = DM_SQRTF(velx * velx + vely * vely);
        bw *= dm_smoothstep(0.0f, 1.0f, 1.0f - fabsf(prev.w + cspos.z) * vlen * 8.0f);
        const float pv[3] = {prev.x, prev.y, prev.z};
        for (int c = 0; c < 3; c++) color[c] = dm_clamp(dm_mix(color[c], pv[c], bw), 0.0f, 16.0f);
    }
    out_taa[0] = dm_f32_to_f16(color[0]); out_taa[1] = dm_f32_to_f16(color[1]); out_taa[2] = dm_f32_to_f16(color[2]);
    out_taa[3] = dm_f32_to_f16(-cspos.z);
    // motion blur, :256-264.  gl_FragCoord.xy = pixel centre
    const float dx = velx * 0.125f, dy = vely * 0.125f;
    const float rr = cm_rand21((float)x + 0.5f, (float)y + 0.5f);
    float su = u + dx * rr, sv = v + dy * rr;
    for (int i = 0; i < 8; i++)
    {
        su += dx; sv += dy;
        const cv4 t = cm_taa_bilinear(I, su, sv);
        color[0] += t.x; color[1] += t.y; color[2] += t.z;
    }
    out_color[0] = dm_f32_to_f16(color[0] / 9.0f); out_color[1] = dm_f32_to_f16(color[1] / 9.0f); out_color[2] = dm_f32_to_f16(color[2] / 9.0f);
    out_color[3] = dm_f32_to_f16(1.0f);
}
