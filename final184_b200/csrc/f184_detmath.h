// f184_detmath.h — deterministic single-precision elementary functions, identical on host and device.
//
// Why this exists: the reference's indirect pass seeds its rays with `fract(sin(x) * 43758.5453)`
// (Shader/math.inc:107-110) for |x| up to ~2e5, then feeds `log`, `sqrt`, `cos` (math.inc:189-194) and
// decodes voxel colour with `pow(c, 2.2)` (Shader/Lighting/indirect.frag:157).  GLSL leaves the
// precision of sin/cos/log/pow/exp2 to the driver, and one ulp of `sin` at |sin| ~ 1 moves the hash by
// ~2.6e-3, i.e. a different ray.  "The reference" therefore has no unique answer; parity needs ONE
// pinned definition.  This header is that definition: every function below uses only IEEE-754
// +,-,*,/ (fp32 and fp64), conversions and bit operations, all of which round identically on x86-64
// and on sm_100a.  tests/test_detmath.py bounds the error against libm (<= 2 ulp on the ranges the path
// uses), so these are honest sin/cos/log/exp2/pow, not something else with the same name.
//
// Bit-exactness contract: translation units that need host/device agreement are compiled with
// contraction off (nvcc -fmad=false, gcc -ffp-contract=off); the fp64 range reduction additionally
// uses explicit non-contracting intrinsics on the device so it is safe under any flag.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#else
#define DM_HD static inline
#endif

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif
#if defined(__CUDA_ARCH__)
#define DM_DMUL(a, b) __dmul_rn((a), (b))
#define DM_DSUB(a, b) __dsub_rn((a), (b))
#define DM_RINT(a) rint(a)
#define DM_FLOORF(a) floorf(a)
#define DM_SQRTF(a) __fsqrt_rn(a)
#else
#include <math.h>
#include <string.h>
#define DM_DMUL(a, b) ((a) * (b))
#define DM_DSUB(a, b) ((a) - (b))
#define DM_RINT(a) rint(a)
#define DM_FLOORF(a) floorf(a)
#define DM_SQRTF(a) sqrtf(a)
#endif

DM_HD uint32_t dm_f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
DM_HD float dm_u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
DM_HD bool dm_isnan(float f) { return (dm_f2u(f) & 0x7fffffffu) > 0x7f800000u; }

// ---- sin / cos ---------------------------------------------------------------------------------
// Range reduction in fp64 (two-term Cody-Waite: k * P1 is exact for |k| < 2^20), then the classic
// minimax polynomials on [-pi/4, pi/4] evaluated in fp32 Horner form with separate multiply and add.
DM_HD void dm_sincos_reduce(float x, float* r, int* q)
{
    const double TWO_OVER_PI = 0.63661977236758134308;
    const double P1 = 1.5707963267341256e+00;   // pi/2 rounded to 33 significant bits
    const double P2 = 6.0771005065061922e-11;   // pi/2 - P1
    double xd = (double)x;
    double kd = DM_RINT(DM_DMUL(xd, TWO_OVER_PI));
    double rd = DM_DSUB(DM_DSUB(xd, DM_DMUL(kd, P1)), DM_DMUL(kd, P2));
    *r = (float)rd;
    *q = (int)((long long)kd & 3);
}
DM_HD float dm_sin_poly(float r)
{
    float z = r * r;
    float p = -1.9515295891e-4f;
    p = p * z + 8.3321608736e-3f;
    p = p * z + -1.6666654611e-1f;
    return (p * z) * r + r;
}
DM_HD float dm_cos_poly(float r)
{
    float z = r * r;
    float p = 2.443315711809948e-5f;
    p = p * z + -1.388731625493765e-3f;
    p = p * z + 4.166664568298827e-2f;
    return (p * z) * z + (1.0f - 0.5f * z);
}
DM_HD float dm_sin(float x)
{
    if (dm_isnan(x) || (dm_f2u(x) & 0x7fffffffu) >= 0x4b000000u) return x - x;   // NaN/inf/|x|>=2^23 -> NaN or 0
    float r; int q;
    dm_sincos_reduce(x, &r, &q);
    float s = (q & 1) ? dm_cos_poly(r) : dm_sin_poly(r);
    return (q & 2) ? -s : s;
}
DM_HD float dm_cos(float x)
{
    if (dm_isnan(x) || (dm_f2u(x) & 0x7fffffffu) >= 0x4b000000u) return (x - x) + 1.0f;
    float r; int q;
    dm_sincos_reduce(x, &r, &q);
    float c = (q & 1) ? dm_sin_poly(r) : dm_cos_poly(r);
    return ((q + 1) & 2) ? -c : c;
}

// ---- log / log2 --------------------------------------------------------------------------------
// x = 2^e * m, m in [sqrt(1/2), sqrt(2)); log(m) = f - f^2/2 + f^3 * P(f), f = m - 1 (Cephes logf).
DM_HD float dm_log_core(float x, float* e_out)
{
    uint32_t u = dm_f2u(x);
    int e = 0;
    if (u < 0x00800000u) { x = x * 8388608.0f; u = dm_f2u(x); e = -23; }   // subnormal
    e += (int)(u >> 23) - 126;
    float m = dm_u2f((u & 0x007fffffu) | 0x3f000000u);                      // [0.5, 1)
    if (m < 0.70710678118654752440f) { e -= 1; m = m + m; }
    float f = m - 1.0f;
    float z = f * f;
    float p = 7.0376836292e-2f;
    p = p * f + -1.1514610310e-1f;
    p = p * f + 1.1676998740e-1f;
    p = p * f + -1.2420140846e-1f;
    p = p * f + 1.4249322787e-1f;
    p = p * f + -1.6668057665e-1f;
    p = p * f + 2.0000714765e-1f;
    p = p * f + -2.4999993993e-1f;
    p = p * f + 3.3333331174e-1f;
    float y = (f * z) * p;
    y = y + -0.5f * z;
    *e_out = (float)e;
    return f + y;            // log(m)
}
DM_HD float dm_log(float x)
{
    if (dm_isnan(x) || x < 0.0f) return dm_u2f(0x7fc00000u);
    if (x == 0.0f) return dm_u2f(0xff800000u);
    if (dm_f2u(x) == 0x7f800000u) return x;
    float e;
    float lm = dm_log_core(x, &e);
    return (lm + e * -2.12194440e-4f) + e * 0.693359375f;
}
DM_HD float dm_log2(float x)
{
    if (dm_isnan(x) || x < 0.0f) return dm_u2f(0x7fc00000u);
    if (x == 0.0f) return dm_u2f(0xff800000u);
    if (dm_f2u(x) == 0x7f800000u) return x;
    float e;
    float lm = dm_log_core(x, &e);
    // log2(m) = lm * log2(e), split so the leading term is exact
    return (lm * 4.4269504088896340736e-1f + lm) + e;
}

// ---- exp2 / pow --------------------------------------------------------------------------------
DM_HD float dm_exp2(float x)
{
    if (dm_isnan(x)) return x;
    if (x >= 128.0f) return dm_u2f(0x7f800000u);
    if (x < -126.0f) return 0.0f;            // flush: the path only needs weights that vanish
    float n = DM_FLOORF(x + 0.5f);
    float f = x - n;                         // [-0.5, 0.5]
    float p = 1.535336188319500e-4f;
    p = p * f + 1.339887440266574e-3f;
    p = p * f + 9.618437357674640e-3f;
    p = p * f + 5.550332471162809e-2f;
    p = p * f + 2.402264791363012e-1f;
    p = p * f + 6.931472028550421e-1f;
    p = p * f + 1.0f;
    int ni = (int)n;
    return p * dm_u2f((uint32_t)(ni + 127) << 23);
}
// pow for the path's use: x >= 0 (unpacked colours), finite y.  pow(0, y>0) = 0.
DM_HD float dm_pow(float x, float y)
{
    if (x == 0.0f) return 0.0f;
    return dm_exp2(y * dm_log2(x));
}

// ---- GLSL helpers with pinned operation order --------------------------------------------------
DM_HD float dm_fract(float x) { return x - DM_FLOORF(x); }
// min/max that return the non-NaN operand (NVIDIA GLSL behaviour)
DM_HD float dm_min(float a, float b) { return dm_isnan(a) ? b : (dm_isnan(b) ? a : (a < b ? a : b)); }
DM_HD float dm_max(float a, float b) { return dm_isnan(a) ? b : (dm_isnan(b) ? a : (a > b ? a : b)); }
DM_HD float dm_clamp(float x, float lo, float hi) { return dm_min(dm_max(x, lo), hi); }
DM_HD float dm_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
DM_HD float dm_step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
DM_HD float dm_smoothstep(float e0, float e1, float x)
{
    float t = dm_clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
// float -> int conversions as the GPU does them: NaN -> 0, saturating, truncate toward zero
DM_HD int dm_f2i(float f)
{
#if defined(__CUDA_ARCH__)
    return __float2int_rz(f);            // cvt.rzi.s32.f32: NaN -> 0, saturating — the definition below, in one instruction
#endif
    if (dm_isnan(f)) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}
DM_HD uint32_t dm_f2uint(float f)
{
#if defined(__CUDA_ARCH__)
    return __float2uint_rz(f);           // cvt.rzi.u32.f32: NaN -> 0, negative -> 0, saturating
#endif
    if (dm_isnan(f) || f <= 0.0f) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}

// ---- fp16 (RGBA16F render targets) -------------------------------------------------------------
DM_HD uint16_t dm_f32_to_f16(float f)   // round-to-nearest-even, IEEE binary16
{
#if defined(__CUDA_ARCH__)
    // cvt.rn.f16.f32 rounds exactly as below; only its NaN encoding differs (0x7fff), so NaN keeps the explicit form
    if (dm_isnan(f)) return (uint16_t)(((dm_f2u(f) >> 16) & 0x8000u) | 0x7e00u);
    return __half_as_ushort(__float2half_rn(f));
#endif
    uint32_t u = dm_f2u(f);
    uint32_t sign = (u >> 16) & 0x8000u;
    uint32_t a = u & 0x7fffffffu;
    if (a > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);              // NaN
    if (a >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);             // >= 65520 -> inf
    if (a < 0x33000001u) return (uint16_t)sign;                          // < 2^-25 (or exactly) -> 0
    int e = (int)(a >> 23) - 127;
    uint32_t m = (a & 0x007fffffu) | 0x00800000u;
    int shift;
    uint32_t he;
    if (e < -14) { shift = 13 + (-14 - e); he = 0; }                     // subnormal half
    else { shift = 13; he = (uint32_t)(e + 15); }
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    uint32_t h = (he == 0) ? q : (((he - 1u) << 10) + q);                // q carries the implicit bit
    return (uint16_t)(sign | h);
}
DM_HD float dm_f16_to_f32(uint16_t h)
{
#if defined(__CUDA_ARCH__)
    if ((h & 0x7c00u) != 0x7c00u) return __half2float(__ushort_as_half(h));      // exact for every finite half; inf/NaN keep the bit form below
#endif
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    if (e == 0)
    {
        if (m == 0) return dm_u2f(sign);
        float v = (float)m * 5.9604644775390625e-8f;                      // m * 2^-24
        return sign ? -v : v;
    }
    if (e == 31) return dm_u2f(sign | 0x7f800000u | (m << 13));
    return dm_u2f(sign | ((e + 112u) << 23) | (m << 13));
}
