// f184_cone.cuh — the cone set of the north-star tracer (DESIGN.md "Mode N" B.5) and its level selection, shared by the tracer
// (mode_n_trace.cu) and by the peer gather's "does level 0 have to travel?" test (mode_n_shard.cu), so both evaluate the SAME
// arithmetic.  The reference has no cones (Shader/Lighting/indirect.frag marches fixed 0.2 m steps through one level).
#pragma once
#include "f184_device.cuh"

constexpr float kTanHalfDiffuse = 0.57735027f;          // tan(30 deg): six 60-degree diffuse cones

// specular cone about the mirror direction: tan(theta/2) = clamp(roughness^2, 0.02, 0.6)
__device__ __forceinline__ float cone_specular_tan(float rough) { return dm_clamp(rough * rough, 0.02f, 0.6f); }

// lod of a cone sample at distance t: diameter of the cone there in voxels, log2
__device__ __forceinline__ float cone_lod(float t, float tan_half, float h, float inv_h, float* diam_out)
{
    const float diam = fmaxf(h, 2.0f * t * tan_half);
    *diam_out = diam;
    return __log2f(diam * inv_h);
}

// Does the cone ever sample level 0?  Its first sample (t = 2h) is its finest — the footprint only grows — and level 0 is read
// while lod < 0.5 (nearest-level spec: L = floor(lod + 0.5) = 0) or lod < 1 (Appendix-B spec: the blend between level 0 and
// level 1).  A small margin keeps the answer conservative against a last-bit difference in h between the two callers.
__device__ __forceinline__ bool cone_samples_level0(float tan_half, float h, bool spec_b)
{
    float diam;
    const float lod = cone_lod(2.0f * h, tan_half, h, 1.0f / h, &diam);
    return lod < (spec_b ? 1.0f : 0.5f) + 1e-3f;
}
