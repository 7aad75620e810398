// f184_device.cuh — small device-side vector/matrix helpers with a PINNED operation order.
// Mode R kernels are compiled with -fmad=false so every expression below rounds exactly as written;
// the order mirrors the GLSL the reference compiles (mat * vec accumulates column by column).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "f184_detmath.h"
#include "f184_internal.h"

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ f3 neg3(f3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
__device__ __forceinline__ float length3(f3 a) { return __fsqrt_rn(dot3(a, a)); }
__device__ __forceinline__ f3 normalize3(f3 a) { float l = length3(a); return {a.x / l, a.y / l, a.z / l}; }
__device__ __forceinline__ f3 abs3(f3 a) { return {fabsf(a.x), fabsf(a.y), fabsf(a.z)}; }

__device__ __forceinline__ f4 mul44(const M4& M, f4 v)
{
    f4 r;
    r.x = ((M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z) + M.m[12] * v.w;
    r.y = ((M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z) + M.m[13] * v.w;
    r.z = ((M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z) + M.m[14] * v.w;
    r.w = ((M.m[3] * v.x + M.m[7] * v.y) + M.m[11] * v.z) + M.m[15] * v.w;
    return r;
}
// xyz rows only (callers that discard .w)
__device__ __forceinline__ f3 mul43(const M4& M, f3 v, float w)
{
    f3 r;
    r.x = ((M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z) + M.m[12] * w;
    r.y = ((M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z) + M.m[13] * w;
    r.z = ((M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z) + M.m[14] * w;
    return r;
}
// mat3(M) * v
__device__ __forceinline__ f3 mul33(const M4& M, f3 v)
{
    return {(M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z, (M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z,
            (M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z};
}

// z of slice `z` of direction d inside a level of the six-direction atlas (f184_internal.h: dir_atlas); n = edge of the level
__device__ __forceinline__ int atlas_z(int d, int n, int z) { return 2 * d * n + z; }

__device__ __forceinline__ int wrap_pow2(int i, int n) { return i & (n - 1); }   // n is a power of two; works for negatives

// Warp-aggregated add of a per-lane count to a global counter: one atomic per warp.
__device__ __forceinline__ void warp_count_add(unsigned long long* counter, unsigned int v)
{
    unsigned int s = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(counter, (unsigned long long)s);
}
