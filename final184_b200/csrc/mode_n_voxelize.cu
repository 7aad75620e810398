// mode_n_voxelize.cu — north-star voxelization: conservative triangle/voxel overlap into fp32 sum+count
// accumulators, then the normalise pass (F184_MODE_NORTHSTAR, DESIGN.md "Mode N" B.1-B.2).
//
// Replaces ClearImage + the voxelization pass (Foreground/Renderer/MegaPipeline.cpp:196, 218-223) with what
// BASELINE.json:north_star asks for instead of the reference's centre-sample raster + racy imageStore
// (Pipelang/Internal/main.lua:242-275): every voxel a triangle TOUCHES receives a fragment, fragments are
// averaged, and the result does not depend on scheduling.
//
// B200 design
//   * triangle-parallel, one launch: each lane sets one triangle up (transform, snap to the 1/256-voxel
//     lattice, integer normal, dominant axis); the warp then drains its 32 triangles one at a time with
//     lanes striding over the COLUMNS of the dominant-axis projection (work ~ projected area, so a wall
//     spanning 10^4 columns costs the same per column as a sub-voxel triangle), 1-4 candidate voxels per
//     column decided by an exact int64 separating-axis test.
//   * accumulation uses the sm_90+ 16-byte vector reduction red.global.add.v4.f32 (atomicAdd(float4*)):
//     two of them per fragment (colour+count, normal) instead of seven scalar atomics.  Every addend is an
//     integer (8-bit colour, 8-bit signed normal, 1) so the fp32 sums are exact for < 2^16 fragments per
//     voxel: order-independent, bit-reproducible, and a multi-GPU sum of partial volumes is exact too.
//   * accumulators are brick-major (8^3 voxels contiguous, 8 KB per brick per volume); a per-brick flag
//     records what was touched so normalise reads, converts and re-zeroes only touched bricks — the dense
//     32 B/voxel volume is never streamed (at 512^3 that alone would be 1.3 ms of HBM time).
#include <algorithm>

#include "f184_device.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 4;

struct TriN
{
    int v[3][3];             // snapped vertices (1/256 voxel)
    long long n[3];          // integer normal
    long long eu[3], ev[3];  // sign-normalised 2D edges in the projection
    long long area;
    int lo[3], hi[3];
    float u[3], vv[3];
    f3 nrm[3];
    float dudx, dvdx, dudy, dvdy;
    uint16_t mat;
    uint8_t d, pad;
};

__device__ __forceinline__ f4 bilinear_level(const TexDev& t, uint32_t level, float u, float v)
{
    uint32_t w = max(1u, t.w >> level), h = max(1u, t.h >> level);
    const uchar4* px = reinterpret_cast<const uchar4*>(t.base + t.off[level]);
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = wrap_pow2(dm_f2i(x0f), (int)w), y0 = wrap_pow2(dm_f2i(y0f), (int)h);
    int x1 = wrap_pow2(x0 + 1, (int)w), y1 = wrap_pow2(y0 + 1, (int)h);
    uchar4 t00 = __ldg(px + (size_t)y0 * w + x0), t10 = __ldg(px + (size_t)y0 * w + x1);
    uchar4 t01 = __ldg(px + (size_t)y1 * w + x0), t11 = __ldg(px + (size_t)y1 * w + x1);
    auto lerp2 = [&](unsigned char a, unsigned char b, unsigned char c, unsigned char d) {
        float fa = (float)a / 255.0f, fb = (float)b / 255.0f, fc = (float)c / 255.0f, fd = (float)d / 255.0f;
        float top = fa * (1.0f - fx) + fb * fx, bot = fc * (1.0f - fx) + fd * fx;
        return top * (1.0f - fy) + bot * fy;
    };
    return {lerp2(t00.x, t10.x, t01.x, t11.x), lerp2(t00.y, t10.y, t01.y, t11.y), lerp2(t00.z, t10.z, t01.z, t11.z),
            lerp2(t00.w, t10.w, t01.w, t11.w)};
}

__device__ __forceinline__ f4 sample_trilinear(const TexDev& t, float u, float v, float dudx, float dvdx, float dudy, float dvdy)
{
    float ax = dudx * (float)t.w, ay = dvdx * (float)t.h, bx = dudy * (float)t.w, by = dvdy * (float)t.h;
    float mx = __fsqrt_rn(ax * ax + ay * ay), my = __fsqrt_rn(bx * bx + by * by);
    float rho = mx > my ? mx : my;
    float maxlod = (float)min(4u, t.nlevels - 1u);
    float lod = 0.0f;
    if (rho > 1.0f) lod = dm_log2(rho);
    if (!(lod < maxlod)) lod = maxlod;
    float lf = floorf(lod);
    uint32_t l0 = (uint32_t)lf;
    float f = lod - lf;
    f4 c0 = bilinear_level(t, l0, u, v);
    if (f == 0.0f) return c0;
    f4 c1 = bilinear_level(t, l0 + 1, u, v);
    return {c0.x * (1.0f - f) + c1.x * f, c0.y * (1.0f - f) + c1.y * f, c0.z * (1.0f - f) + c1.z * f, c0.w * (1.0f - f) + c1.w * f};
}

__device__ __forceinline__ long long lmin3(long long a, long long b, long long c) { return min(a, min(b, c)); }
__device__ __forceinline__ long long lmax3(long long a, long long b, long long c) { return max(a, max(b, c)); }

// Exact triangle / closed-box overlap on the integer lattice (Akenine-Moller SAT, 13 axes).
__device__ bool tri_box_overlap(const TriN& s, int bx, int by, int bz)
{
    long long p[3][3];
    const int c[3] = {256 * bx + 128, 256 * by + 128, 256 * bz + 128};
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int a = 0; a < 3; a++) p[k][a] = (long long)(s.v[k][a] - c[a]);
    const long long hs = 128;
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
        if (lmin3(p[0][a], p[1][a], p[2][a]) > hs || lmax3(p[0][a], p[1][a], p[2][a]) < -hs) return false;
    }
    {
        const long long d = s.n[0] * p[0][0] + s.n[1] * p[0][1] + s.n[2] * p[0][2];
        const long long r = hs * (llabs(s.n[0]) + llabs(s.n[1]) + llabs(s.n[2]));
        if (d > r || d < -r) return false;
    }
#pragma unroll
    for (int e = 0; e < 3; e++)
    {
        const int e1 = (e + 1) % 3;
        const long long ex = p[e1][0] - p[e][0], ey = p[e1][1] - p[e][1], ez = p[e1][2] - p[e][2];
        {   // axis (0, -ez, ey)
            const long long q0 = -ez * p[0][1] + ey * p[0][2], q1 = -ez * p[1][1] + ey * p[1][2], q2 = -ez * p[2][1] + ey * p[2][2];
            const long long r = hs * (llabs(ez) + llabs(ey));
            if (lmin3(q0, q1, q2) > r || lmax3(q0, q1, q2) < -r) return false;
        }
        {   // axis (ez, 0, -ex)
            const long long q0 = ez * p[0][0] - ex * p[0][2], q1 = ez * p[1][0] - ex * p[1][2], q2 = ez * p[2][0] - ex * p[2][2];
            const long long r = hs * (llabs(ez) + llabs(ex));
            if (lmin3(q0, q1, q2) > r || lmax3(q0, q1, q2) < -r) return false;
        }
        {   // axis (-ey, ex, 0)
            const long long q0 = -ey * p[0][0] + ex * p[0][1], q1 = -ey * p[1][0] + ex * p[1][1], q2 = -ey * p[2][0] + ex * p[2][1];
            const long long r = hs * (llabs(ey) + llabs(ex));
            if (lmin3(q0, q1, q2) > r || lmax3(q0, q1, q2) < -r) return false;
        }
    }
    return true;
}

__device__ __forceinline__ int floor_div256(int a) { return a >> 8; }

__device__ __forceinline__ size_t brick_major(int x, int y, int z, int NB)
{
    const size_t brick = ((size_t)(z >> 3) * NB + (y >> 3)) * NB + (x >> 3);
    return brick * 512 + ((z & 7) << 6) + ((y & 7) << 3) + (x & 7);
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_voxelize_n(const float* __restrict__ pos, const float* __restrict__ nrm, const float* __restrict__ uv,
             const uint32_t* __restrict__ idx, const uint16_t* __restrict__ tri_mat, const uint16_t* __restrict__ tri_model,
             const M4* __restrict__ model_mats, const M4* __restrict__ vm_mats, M4 Proj, const TexDev* __restrict__ texs,
             const MatDev* __restrict__ mats, uint32_t tri_first, uint32_t tri_end, int N,
             float4* __restrict__ accC, float4* __restrict__ accN, uint32_t* __restrict__ brick_flags,
             unsigned long long* __restrict__ frag_counter)
{
    __shared__ TriN sh[WARPS_PER_BLOCK][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t t = tri_first + (blockIdx.x * WARPS_PER_BLOCK + warp) * 32 + lane;
    const float Nf = (float)N;
    const int NB = N >> 3;
    bool active = false;

    if (t < tri_end)
    {
        const uint32_t id[3] = {idx[3 * t], idx[3 * t + 1], idx[3 * t + 2]};
        const uint32_t model = tri_model[t];
        const M4& vm = vm_mats[model];
        TriN& s = sh[warp][lane];
        bool bad = false;
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            f3 q = mul43(vm, f3{pos[3 * id[i]], pos[3 * id[i] + 1], pos[3 * id[i] + 2]}, 1.0f);
            f4 g = mul44(Proj, f4{q.x, q.y, q.z, 1.0f});
            const float gx = g.x / g.w, gy = g.y / g.w, gz = g.z / g.w;
            const float vx = (gx * 0.5f + 0.5f) * Nf, vy = (gy * 0.5f + 0.5f) * Nf, vz = gz * Nf;
            if (!(vx >= -Nf && vx < 2.0f * Nf) || !(vy >= -Nf && vy < 2.0f * Nf) || !(vz >= -Nf && vz < 2.0f * Nf)) bad = true;
            s.v[i][0] = (int)rintf(vx * 256.0f); s.v[i][1] = (int)rintf(vy * 256.0f); s.v[i][2] = (int)rintf(vz * 256.0f);
        }
        if (!bad)
        {
            long long e1[3], e2[3];
#pragma unroll
            for (int a = 0; a < 3; a++) { e1[a] = (long long)s.v[1][a] - s.v[0][a]; e2[a] = (long long)s.v[2][a] - s.v[0][a]; }
            s.n[0] = e1[1] * e2[2] - e1[2] * e2[1];
            s.n[1] = e1[2] * e2[0] - e1[0] * e2[2];
            s.n[2] = e1[0] * e2[1] - e1[1] * e2[0];
            if (s.n[0] != 0 || s.n[1] != 0 || s.n[2] != 0)
            {
                const long long anx = llabs(s.n[0]), any = llabs(s.n[1]), anz = llabs(s.n[2]);
                int d;
                if (anx > any) d = (anx > anz) ? 0 : 2;
                else d = (any > anz) ? 1 : 2;
                const int ua = (d + 1) % 3, va = (d + 2) % 3;
                bool empty = false;
#pragma unroll
                for (int a = 0; a < 3; a++)
                {
                    const int mn = min(s.v[0][a], min(s.v[1][a], s.v[2][a])), mx = max(s.v[0][a], max(s.v[1][a], s.v[2][a]));
                    s.lo[a] = max(0, floor_div256(mn - 1));
                    s.hi[a] = min(N - 1, floor_div256(mx));
                    if (s.lo[a] > s.hi[a]) empty = true;
                }
                if (!empty)
                {
                    active = true;
                    long long area = s.n[d];
                    const long long sg = area < 0 ? -1 : 1;
                    area *= sg;
                    s.area = area;
#pragma unroll
                    for (int k = 0; k < 3; k++)
                    {
                        const int a = (k + 1) % 3, b = (k + 2) % 3;
                        s.eu[k] = sg * ((long long)s.v[b][ua] - s.v[a][ua]);
                        s.ev[k] = sg * ((long long)s.v[b][va] - s.v[a][va]);
                    }
                    const M4& mm = model_mats[model];
#pragma unroll
                    for (int i = 0; i < 3; i++)
                    {
                        s.u[i] = uv[2 * id[i]]; s.vv[i] = uv[2 * id[i] + 1];
                        s.nrm[i] = normalize3(mul33(mm, f3{nrm[3 * id[i]], nrm[3 * id[i] + 1], nrm[3 * id[i] + 2]}));
                    }
                    const float areaf = (float)area;
                    float dbdu[3], dbdv[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) { dbdu[k] = (float)(-s.ev[k] * 256) / areaf; dbdv[k] = (float)(s.eu[k] * 256) / areaf; }
                    s.dudx = (s.u[0] * dbdu[0] + s.u[1] * dbdu[1]) + s.u[2] * dbdu[2];
                    s.dvdx = (s.vv[0] * dbdu[0] + s.vv[1] * dbdu[1]) + s.vv[2] * dbdu[2];
                    s.dudy = (s.u[0] * dbdv[0] + s.u[1] * dbdv[1]) + s.u[2] * dbdv[2];
                    s.dvdy = (s.vv[0] * dbdv[0] + s.vv[1] * dbdv[1]) + s.vv[2] * dbdv[2];
                    s.mat = tri_mat[t];
                    s.d = (uint8_t)d;
                }
            }
        }
    }
    unsigned int pending = __ballot_sync(0xffffffffu, active);
    __syncwarp();
    unsigned int frags = 0;
    while (pending)
    {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const TriN& s = sh[warp][src];
        const int d = s.d, ua = (d + 1) % 3, va = (d + 2) % 3;
        const int bu = s.hi[ua] - s.lo[ua] + 1, bv = s.hi[va] - s.lo[va] + 1;
        const int ncols = bu * bv;
        const float areaf = (float)s.area;
        const MatDev mat = mats[s.mat];
        // plane: x_d = v0_d - (n_u (x_u - v0_u) + n_v (x_v - v0_v)) / n_d   (|n_d| is the largest component)
        const double inv_nd = 1.0 / (double)s.n[d];
        for (int col = lane; col < ncols; col += 32)
        {
            const int iu = s.lo[ua] + col % bu, iv = s.lo[va] + col / bu;
            // depth interval of the plane over the column footprint, widened by one voxel; the exact SAT decides
            double dmin = 1e300, dmax = -1e300;
#pragma unroll
            for (int cc = 0; cc < 4; cc++)
            {
                const double xu = (double)(256 * (iu + (cc & 1)) - s.v[0][ua]), xv = (double)(256 * (iv + (cc >> 1)) - s.v[0][va]);
                const double xd = (double)s.v[0][d] - ((double)s.n[ua] * xu + (double)s.n[va] * xv) * inv_nd;
                dmin = fmin(dmin, xd); dmax = fmax(dmax, xd);
            }
            int k0 = max(s.lo[d], (int)floor(dmin / 256.0) - 1), k1 = min(s.hi[d], (int)floor(dmax / 256.0) + 1);
            for (int kd = k0; kd <= k1; kd++)
            {
                int b[3];
                b[ua] = iu; b[va] = iv; b[d] = kd;
                if (!tri_box_overlap(s, b[0], b[1], b[2])) continue;
                // barycentrics of the voxel centre in the projection, clamped into the triangle
                const long long cu = 256 * iu + 128, cv = 256 * iv + 128;
                float bc[3];
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    const int a = (k + 1) % 3;
                    const long long w = s.eu[k] * (cv - s.v[a][va]) - s.ev[k] * (cu - s.v[a][ua]);
                    bc[k] = (float)w / areaf;
                    if (bc[k] < 0.0f) bc[k] = 0.0f;
                }
                const float sum = (bc[0] + bc[1]) + bc[2];
                bc[0] = bc[0] / sum; bc[1] = bc[1] / sum; bc[2] = bc[2] / sum;
                const float u = (s.u[0] * bc[0] + s.u[1] * bc[1]) + s.u[2] * bc[2];
                const float v = (s.vv[0] * bc[0] + s.vv[1] * bc[1]) + s.vv[2] * bc[2];
                f3 nn = {(s.nrm[0].x * bc[0] + s.nrm[1].x * bc[1]) + s.nrm[2].x * bc[2], (s.nrm[0].y * bc[0] + s.nrm[1].y * bc[1]) + s.nrm[2].y * bc[2],
                         (s.nrm[0].z * bc[0] + s.nrm[1].z * bc[1]) + s.nrm[2].z * bc[2]};
                f4 base;
                if (!mat.use_textures) base = {mat.factor[0], mat.factor[1], mat.factor[2], mat.factor[3]};
                else
                {
                    f4 sc = {0.f, 0.f, 0.f, 0.f};
                    if (mat.tex >= 0) sc = sample_trilinear(texs[mat.tex], u, v, s.dudx, s.dvdx, s.dudy, s.dvdy);
                    base = {sc.x * mat.factor[0], sc.y * mat.factor[1], sc.z * mat.factor[2], sc.w * mat.factor[3]};
                    if (base.w < 0.05f) continue;                       // main.lua:199
                }
                const float ax = fabsf(nn.x), ay = fabsf(nn.y), az = fabsf(nn.z);
                const float lead = (ax >= ay && ax >= az) ? nn.x : ((ay >= az) ? nn.y : nn.z);
                if (lead < 0.0f) nn = neg3(nn);
                const float r8 = floorf(dm_clamp(base.x, 0.0f, 1.0f) * 255.0f + 0.5f);
                const float g8 = floorf(dm_clamp(base.y, 0.0f, 1.0f) * 255.0f + 0.5f);
                const float b8 = floorf(dm_clamp(base.z, 0.0f, 1.0f) * 255.0f + 0.5f);
                const float nx8 = rintf(dm_clamp(nn.x, -1.0f, 1.0f) * 127.0f), ny8 = rintf(dm_clamp(nn.y, -1.0f, 1.0f) * 127.0f),
                            nz8 = rintf(dm_clamp(nn.z, -1.0f, 1.0f) * 127.0f);
                const size_t o = brick_major(b[0], b[1], b[2], NB);
                atomicAdd(accC + o, make_float4(r8, g8, b8, 1.0f));     // red.global.add.v4.f32
                atomicAdd(accN + o, make_float4(nx8, ny8, nz8, 0.0f));
                brick_flags[o >> 9] = 1u;
                frags++;
            }
        }
    }
    warp_count_add(frag_counter, frags);
}

// vm[m] = View * Model[m] (same association as mode R)
__global__ void k_view_model_n(M4 View, const M4* __restrict__ model, M4* __restrict__ vm, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 16) return;
    const uint32_t m = i / 16, j = (i % 16) / 4, r = i % 4;
    const M4& B = model[m];
    vm[m].m[4 * j + r] = ((View.m[r] * B.m[4 * j] + View.m[4 + r] * B.m[4 * j + 1]) + View.m[8 + r] * B.m[4 * j + 2]) + View.m[12 + r] * B.m[4 * j + 3];
}

// ---- B.2 normalise: touched (or previously occupied) bricks only ----------------------------------------
// One warp per brick, 16 passes of 32 voxels: 512 B coalesced accumulator reads, 32 B output runs.
// Reads the sums, writes mean albedo / unit normal as RGBA8, re-zeroes the accumulators (= next frame's
// clear) and appends the brick to the frame's brick list for the stages downstream.
__global__ void __launch_bounds__(256)
k_normalise_n(float4* __restrict__ accC, float4* __restrict__ accN, uint32_t* __restrict__ brick_flags, uint32_t* __restrict__ brick_prev,
              uchar4* __restrict__ alb, char4* __restrict__ nrm, uint32_t* __restrict__ brick_list, unsigned long long* __restrict__ counters,
              int N, uint32_t n_bricks)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int NB = N >> 3;
    unsigned int occ = 0, nbricks = 0;
    for (uint32_t base = warp_global * 32; base < n_bricks; base += n_warps * 32)
    {
        const uint32_t bi = base + lane;
        uint32_t flag = 0, prev = 0;
        if (bi < n_bricks) { flag = brick_flags[bi]; prev = brick_prev[bi]; }
        if (bi < n_bricks && (flag | prev)) { brick_flags[bi] = 0; brick_prev[bi] = flag; }
        unsigned int todo = __ballot_sync(0xffffffffu, (flag | prev) != 0);
        const unsigned int touched = __ballot_sync(0xffffffffu, flag != 0);
        nbricks += (lane == 0) ? __popc(touched) : 0;
        if (lane == 0 && todo)
        {
            const uint32_t cnt = __popc(todo);
            const uint32_t slot = (uint32_t)atomicAdd(counters + F184_COUNTER_COUNT, (unsigned long long)cnt);   // list cursor lives past the counters
            uint32_t k = 0;
            for (unsigned int m = todo; m; m &= m - 1) brick_list[slot + k++] = base + (__ffs(m) - 1);
        }
        while (todo)
        {
            const uint32_t b = base + (__ffs(todo) - 1);
            const bool is_touched = (touched >> (__ffs(todo) - 1)) & 1u;
            todo &= todo - 1;
            const int bx = (b % NB) << 3, by = ((b / NB) % NB) << 3, bz = (b / (NB * NB)) << 3;
#pragma unroll 4
            for (int pass = 0; pass < 16; pass++)
            {
                const int local = pass * 32 + lane;                 // (z&7)<<6 | (y&7)<<3 | (x&7)
                const size_t o = (size_t)b * 512 + local;
                uchar4 a8 = make_uchar4(0, 0, 0, 0);
                char4 n8 = make_char4(0, 0, 0, 0);
                if (is_touched)
                {
                    const float4 c = accC[o];
                    if (c.w > 0.0f)
                    {
                        const float4 nn = accN[o];
                        a8 = make_uchar4((unsigned char)floorf(c.x / c.w + 0.5f), (unsigned char)floorf(c.y / c.w + 0.5f),
                                         (unsigned char)floorf(c.z / c.w + 0.5f), 255);
                        const float len = __fsqrt_rn((nn.x * nn.x + nn.y * nn.y) + nn.z * nn.z);
                        if (len > 0.0f)
                            n8 = make_char4((signed char)rintf(nn.x / len * 127.0f), (signed char)rintf(nn.y / len * 127.0f),
                                            (signed char)rintf(nn.z / len * 127.0f), 0);
                        accC[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                        accN[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                        occ++;
                    }
                }
                const int x = bx + (local & 7), y = by + ((local >> 3) & 7), z = bz + (local >> 6);
                const size_t lin = ((size_t)z * N + y) * N + x;
                alb[lin] = a8;
                nrm[lin] = n8;
            }
        }
    }
    warp_count_add(counters + F184_COUNTER_OCCUPIED, occ);
    warp_count_add(counters + F184_COUNTER_BRICKS, nbricks);
}

}  // namespace

int f184_voxelize_n(f184_ctx* c, const f184_view_constants* cam)
{
    for (int s : {F184_SLOT_ACCUM_COLOR, F184_SLOT_ACCUM_NORMAL, F184_SLOT_VOX_ALBEDO, F184_SLOT_VOX_NORMAL, F184_SLOT_BRICK_FLAGS})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    const int N = (int)c->cfg.grid_n;
    const uint32_t n_bricks = (uint32_t)(N / 8) * (N / 8) * (N / 8);
    if (!c->brick_prev)
    {
        CK(c, cudaMalloc(&c->brick_prev, 4ull * n_bricks));
        CK(c, cudaMemsetAsync(c->brick_prev, 0, 4ull * n_bricks, c->stream));
        CK(c, cudaMalloc(&c->brick_list, 4ull * n_bricks));
    }
    M4 View, Proj;
    memcpy(View.m, cam->ViewMat, 64);
    memcpy(Proj.m, cam->ProjMat, 64);
    M4* vm_dev = nullptr;
    CK(c, cudaMallocAsync(&vm_dev, sizeof(M4) * c->n_models, c->stream));
    k_view_model_n<<<(c->n_models * 16 + 127) / 128, 128, 0, c->stream>>>(View, c->model_mats, vm_dev, c->n_models);
    CK_LAUNCH(c);

    const uint32_t first = c->tri_first < c->n_tris ? c->tri_first : c->n_tris;
    const uint64_t end64 = (uint64_t)first + c->tri_count;
    const uint32_t end = end64 < c->n_tris ? (uint32_t)end64 : c->n_tris;

    int rc = f184_stage_begin(c, F184_STAGE_VOXELIZE);
    if (rc) return rc;
    // counters: fragments, occupied, bricks, and the brick-list cursor (stored right after the public counters)
    CK(c, cudaMemsetAsync(c->counters_dev + F184_COUNTER_FRAGMENTS, 0, 8, c->stream));
    CK(c, cudaMemsetAsync(c->counters_dev + F184_COUNTER_OCCUPIED, 0, 8, c->stream));
    CK(c, cudaMemsetAsync(c->counters_dev + F184_COUNTER_BRICKS, 0, 8, c->stream));
    CK(c, cudaMemsetAsync(c->counters_dev + F184_COUNTER_COUNT, 0, 8, c->stream));
    if (end > first)
    {
        const uint32_t tris = end - first;
        const uint32_t blocks = (tris + WARPS_PER_BLOCK * 32 - 1) / (WARPS_PER_BLOCK * 32);
        k_voxelize_n<<<blocks, WARPS_PER_BLOCK * 32, 0, c->stream>>>(c->pos, c->nrm, c->uv, c->idx, c->tri_mat, c->tri_model, c->model_mats,
                                                                     vm_dev, Proj, c->tex_dev, c->mat_dev, first, end, N,
                                                                     img_ptr<float4>(c, F184_SLOT_ACCUM_COLOR), img_ptr<float4>(c, F184_SLOT_ACCUM_NORMAL),
                                                                     img_ptr<uint32_t>(c, F184_SLOT_BRICK_FLAGS), c->counters_dev + F184_COUNTER_FRAGMENTS);
        CK_LAUNCH(c);
    }
    rc = f184_stage_end(c, F184_STAGE_VOXELIZE);
    if (rc) return rc;
    CK(c, cudaFreeAsync(vm_dev, c->stream));
    if (c->defer_normalise) return F184_OK;      // multi-GPU: partial volumes are summed across ranks first
    return f184_normalise_n(c);
}

int f184_normalise_n(f184_ctx* c)
{
    const int N = (int)c->cfg.grid_n;
    const uint32_t n_bricks = (uint32_t)(N / 8) * (N / 8) * (N / 8);
    int rc = f184_stage_begin(c, F184_STAGE_NORMALISE);
    if (rc) return rc;
    const int blocks = (int)std::min<uint32_t>((n_bricks + 255) / 256, 148 * 8);
    k_normalise_n<<<blocks, 256, 0, c->stream>>>(img_ptr<float4>(c, F184_SLOT_ACCUM_COLOR), img_ptr<float4>(c, F184_SLOT_ACCUM_NORMAL),
                                                 img_ptr<uint32_t>(c, F184_SLOT_BRICK_FLAGS), c->brick_prev,
                                                 img_ptr<uchar4>(c, F184_SLOT_VOX_ALBEDO), img_ptr<char4>(c, F184_SLOT_VOX_NORMAL),
                                                 c->brick_list, c->counters_dev, N, n_bricks);
    CK_LAUNCH(c);
    return f184_stage_end(c, F184_STAGE_NORMALISE);
}
