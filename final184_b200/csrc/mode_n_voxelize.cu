// mode_n_voxelize.cu — north-star voxelization: conservative triangle/voxel overlap into fp32 sum+count
// accumulators, then the normalise pass (F184_MODE_NORTHSTAR, DESIGN.md "Mode N" B.1-B.2).
//
// Replaces ClearImage + the voxelization pass (Foreground/Renderer/MegaPipeline.cpp:196, 218-223) with what
// BASELINE.json:north_star asks for instead of the reference's centre-sample raster + racy imageStore
// (Pipelang/Internal/main.lua:242-275): every voxel a triangle TOUCHES receives a fragment, fragments are
// averaged, and the result does not depend on scheduling.
//
// B200 design
//   * triangle-parallel binning + rasterisation in two launches.  Pass 1 (one thread per triangle) transforms,
//     snaps to the 1/256-voxel lattice and takes the integer normal / dominant axis; triangles whose
//     dominant-axis projection covers <= 8 columns (the bulk of a detailed mesh) are finished by that thread,
//     larger ones are queued as equal 256-column tasks.  Pass 2 (persistent warps) drains the tasks with
//     lanes striding the columns, so work ~ projected area and a wall spanning 10^5 columns is spread over
//     the chip instead of serialising a warp (the first version of this kernel took 168 ms on Sponza 512^3
//     for exactly that reason).  The triangle set-up lives in registers, in the triangle's own (u, v, w)
//     frame, so the exact int64 separating-axis test runs on IMAD.WIDE with no local-memory indexing;
//     the three axes in the projection plane are tested once per column, the other ten per candidate depth.
//   * accumulation uses the sm_90+ 16-byte vector reduction red.global.add.v4.f32 (atomicAdd(float4*)):
//     two of them per fragment (colour+count, normal) instead of seven scalar atomics.  Every addend is an
//     integer (8-bit colour, 8-bit signed normal, 1) so the fp32 sums are exact for < 2^16 fragments per
//     voxel: order-independent, bit-reproducible, and a multi-GPU sum of partial volumes is exact too.
//   * accumulators are brick-major (8^3 voxels contiguous, 8 KB per brick per volume); a per-brick flag
//     records what was touched; a compaction pass turns the flags into a brick list and normalise (one warp
//     per listed brick) reads, converts and re-zeroes only those — the dense 32 B/voxel volume is never
//     streamed (at 512^3 that alone would be 1.3 ms of HBM time).
#include <algorithm>
#include <cstdlib>

#include "f184_device.cuh"

namespace {

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
    // filtered in 8-bit units: bytes enter as 0..255 floats, the result stays on that scale (see the oracle)
    auto lerp2 = [&](unsigned char a, unsigned char b, unsigned char c, unsigned char d) {
        const float fa = (float)a, fb = (float)b, fc = (float)c, fd = (float)d;
        const float top = fa * (1.0f - fx) + fb * fx, bot = fc * (1.0f - fx) + fd * fx;
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

__device__ __forceinline__ int floor_div256(int a) { return a >> 8; }
__device__ __forceinline__ int pick3(int a0, int a1, int a2, int k) { return k == 0 ? a0 : (k == 1 ? a1 : a2); }
__device__ __forceinline__ long long pick3l(long long a0, long long a1, long long a2, int k) { return k == 0 ? a0 : (k == 1 ? a1 : a2); }
__device__ __forceinline__ long long wide(int a, int b) { return (long long)a * (long long)b; }   // one IMAD.WIDE

// 16-byte vector reduction at system scope (sm_90+): the form of red.global.add.v4.f32 that is defined on peer memory
__device__ __forceinline__ void red_add_v4_sys(float4* p, float a, float b, float c, float d)
{
    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ size_t brick_major(int x, int y, int z, int NB)
{
    const size_t brick = ((size_t)(z >> 3) * NB + (y >> 3)) * NB + (x >> 3);
    return brick * 512 + ((z & 7) << 6) + ((y & 7) << 3) + (x & 7);
}

// Triangle set-up, held in REGISTERS.  Everything is stored in the triangle's own frame (u, v, w) =
// (axis d+1, axis d+2, dominant axis d): a cyclic relabelling of x, y, z, under which the 13 separating axes of
// the triangle/box test map onto themselves — so the exact integer predicate below is the oracle's
// tri_box_overlap() evaluated in other coordinates — and every index is a compile-time constant.
struct alignas(16) TriS
{
    long long n[3];          // integer normal (u, v, w components); n[2] is the largest in magnitude
    long long area;          // |n[2]|
    int v[3][3];             // snapped vertices (1/256 voxel), [vertex][u,v,w]
    int sg;                  // sign of n[2]
    int lo[3], hi[3];        // candidate voxel range per axis (u, v, w), clipped to the grid
    float u[3], vv[3];
    f3 nrm[3];
    float dudx, dvdx, dudy, dvdy;
    uint32_t mat;
    int d;
    int pad[3];
};
static_assert(sizeof(TriS) == 192, "TriS is spilled to the pass-2 queue as 12 x 16 bytes");

struct VoxArgs
{
    const float *pos, *nrm, *uv;
    const uint32_t* idx;
    const uint16_t *tri_mat, *tri_model;
    const M4 *model_mats, *vm_mats;
    M4 Proj;
    const TexDev* texs;
    const MatDev* mats;
    uint32_t tri_first, tri_end;
    const uint32_t* chunks;            // 128-triangle chunks of this rank (one per CTA of the set-up pass), or nullptr
    int N;
    // accumulators / brick flags of the rank that owns the voxel's 8^3 brick, owner = (bx + by + bz) % nranks (a diagonal
    // interleave: every axis-aligned sheet of bricks — Sponza's floor is one — is dealt evenly over the ranks): own memory,
    // or a peer's memory mapped over NVLink — then the reductions below ARE the reduce-scatter of the partial volumes.
    // One GPU: owner_mask = 0.
    float4* accC[8];
    float4* accN[8];
    uint32_t* brick_flags[8];
    uint32_t owner_mask, rank;
    // fragments of bricks another rank owns: 16-byte records appended to THIS rank's queue for that rank (local memory; the
    // owner reads it over NVLink behind the barrier); cursor[p][s] = append position in sub-queue s of the region for rank p
    // (a warp appends to the sub-queue its index selects)
    uint4* peer_queue[8];
    uint32_t* cursor;
    uint32_t sub_cap;                  // records per sub-queue
    uint32_t no_aggregate;             // A/B knob (F184_FRAG_NO_AGG=1): every lane reserves its own slot
    unsigned long long* frag_counter;
    unsigned long long* queue_state;   // (entries << 40) | tasks, one 64-bit word so both advance together
    uint2* queue;                      // per large triangle: (triangle, first task)
    uint4* queue_tris;                 // its set-up (TriS, 12 x uint4), same index
};

constexpr int SETUP_THREADS = 128;
constexpr int SMALL_COLS = 8;          // triangles up to this many projection columns are finished by their set-up thread
constexpr int TASK_COLS = 256;         // columns per warp task for the larger ones (survivor indices are bytes)
constexpr int RASTER_THREADS = 128;

// Returns false when the triangle produces nothing (outside the guard band, zero area after snapping, off-grid).
__device__ __forceinline__ bool setup_triangle(const VoxArgs& A, uint32_t t, TriS& s)
{
    const float Nf = (float)A.N;
    const uint32_t id0 = __ldg(A.idx + 3 * t), id1 = __ldg(A.idx + 3 * t + 1), id2 = __ldg(A.idx + 3 * t + 2);
    const uint32_t id[3] = {id0, id1, id2};
    const uint32_t model = __ldg(A.tri_model + t);
    const M4& vm = A.vm_mats[model];
    int vi[3][3];
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        f3 q = mul43(vm, f3{__ldg(A.pos + 3 * id[i]), __ldg(A.pos + 3 * id[i] + 1), __ldg(A.pos + 3 * id[i] + 2)}, 1.0f);
        f4 g = mul44(A.Proj, f4{q.x, q.y, q.z, 1.0f});
        const float gx = g.x / g.w, gy = g.y / g.w, gz = g.z / g.w;
        const float vx = (gx * 0.5f + 0.5f) * Nf, vy = (gy * 0.5f + 0.5f) * Nf, vz = gz * Nf;
        if (!(vx >= -Nf && vx < 2.0f * Nf) || !(vy >= -Nf && vy < 2.0f * Nf) || !(vz >= -Nf && vz < 2.0f * Nf)) bad = true;
        vi[i][0] = (int)rintf(vx * 256.0f); vi[i][1] = (int)rintf(vy * 256.0f); vi[i][2] = (int)rintf(vz * 256.0f);
    }
    if (bad) return false;
    int e1[3], e2[3];
#pragma unroll
    for (int a = 0; a < 3; a++) { e1[a] = vi[1][a] - vi[0][a]; e2[a] = vi[2][a] - vi[0][a]; }
    const long long n0 = wide(e1[1], e2[2]) - wide(e1[2], e2[1]);
    const long long n1 = wide(e1[2], e2[0]) - wide(e1[0], e2[2]);
    const long long n2 = wide(e1[0], e2[1]) - wide(e1[1], e2[0]);
    if (n0 == 0 && n1 == 0 && n2 == 0) return false;
    const long long anx = llabs(n0), any = llabs(n1), anz = llabs(n2);
    int d;
    if (anx > any) d = (anx > anz) ? 0 : 2;
    else d = (any > anz) ? 1 : 2;
    const int ua = (d + 1) % 3, va = (d + 2) % 3;
    bool empty = false;
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        s.v[k][0] = pick3(vi[k][0], vi[k][1], vi[k][2], ua);
        s.v[k][1] = pick3(vi[k][0], vi[k][1], vi[k][2], va);
        s.v[k][2] = pick3(vi[k][0], vi[k][1], vi[k][2], d);
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
        const int mn = min(s.v[0][a], min(s.v[1][a], s.v[2][a])), mx = max(s.v[0][a], max(s.v[1][a], s.v[2][a]));
        s.lo[a] = max(0, floor_div256(mn - 1));
        s.hi[a] = min(A.N - 1, floor_div256(mx));
        if (s.lo[a] > s.hi[a]) empty = true;
    }
    if (empty) return false;
    s.n[0] = pick3l(n0, n1, n2, ua); s.n[1] = pick3l(n0, n1, n2, va); s.n[2] = pick3l(n0, n1, n2, d);
    s.sg = s.n[2] < 0 ? -1 : 1;
    s.area = s.n[2] * s.sg;
    s.d = d;
    const M4& mm = A.model_mats[model];
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        s.u[i] = __ldg(A.uv + 2 * id[i]); s.vv[i] = __ldg(A.uv + 2 * id[i] + 1);
        s.nrm[i] = normalize3(mul33(mm, f3{__ldg(A.nrm + 3 * id[i]), __ldg(A.nrm + 3 * id[i] + 1), __ldg(A.nrm + 3 * id[i] + 2)}));
    }
    const float areaf = (float)s.area;
    float dbdu[3], dbdv[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        const long long eu = (long long)s.sg * (s.v[b][0] - s.v[a][0]), ev = (long long)s.sg * (s.v[b][1] - s.v[a][1]);
        dbdu[k] = (float)(-ev * 256) / areaf; dbdv[k] = (float)(eu * 256) / areaf;
    }
    s.dudx = (s.u[0] * dbdu[0] + s.u[1] * dbdu[1]) + s.u[2] * dbdu[2];
    s.dvdx = (s.vv[0] * dbdu[0] + s.vv[1] * dbdu[1]) + s.vv[2] * dbdu[2];
    s.dudy = (s.u[0] * dbdv[0] + s.u[1] * dbdv[1]) + s.u[2] * dbdv[2];
    s.dvdy = (s.vv[0] * dbdv[0] + s.vv[1] * dbdv[1]) + s.vv[2] * dbdv[2];
    s.mat = __ldg(A.tri_mat + t);
    return true;
}

// The three separating axes that live in the (u, v) projection plane, for column (iu, iv): exact, depth-independent.
__device__ __forceinline__ bool column_in_projection(const TriS& s, int iu, int iv)
{
    const int cu = 256 * iu + 128, cv = 256 * iv + 128;
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        const int eu = s.sg * (s.v[b][0] - s.v[a][0]), ev = s.sg * (s.v[b][1] - s.v[a][1]);
        const long long w = wide(eu, cv - s.v[a][1]) - wide(ev, cu - s.v[a][0]);
        const long long r = 128ll * (abs(eu) + abs(ev));
        if (w < -r || w > s.area + r) return false;
    }
    return true;
}

// One column (iu, iv) of the triangle's dominant-axis projection: exact closed-box overlap (Akenine-Moller SAT on
// the integer lattice, all 13 axes) for every candidate depth, shading and accumulation of the voxels that pass.
__device__ __forceinline__ unsigned int process_column(const VoxArgs& A, const TriS& s, const MatDev& mat, int iu, int iv)
{
    const int hs = 128;
    const int cu = 256 * iu + 128, cv = 256 * iv + 128;
    // -- the three axes that live in the (u, v) plane do not depend on the depth: test them once per column.
    //    W[k] = edge function of edge (k+1 -> k+2) at the column centre, orientation-normalised (inside >= 0);
    //    the separating-axis interval test for that edge reads  -r <= W <= area + r,  r = hs (|eu| + |ev|).
    long long Wk[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        const int eu = s.sg * (s.v[b][0] - s.v[a][0]), ev = s.sg * (s.v[b][1] - s.v[a][1]);
        const long long w = wide(eu, cv - s.v[a][1]) - wide(ev, cu - s.v[a][0]);
        const long long r = (long long)hs * (abs(eu) + abs(ev));
        if (w < -r || w > s.area + r) return 0;
        Wk[k] = w;
    }
    // -- depth candidates: the plane over the column footprint, in float, widened by a voxel; the integer tests decide
    float dmin = 3.0e38f, dmax = -3.0e38f;
    {
        const float inv_nw = 1.0f / (float)s.n[2];
        const float su = (float)s.n[0] * inv_nw, sv = (float)s.n[1] * inv_nw;
#pragma unroll
        for (int cc = 0; cc < 4; cc++)
        {
            const float xu = (float)(256 * (iu + (cc & 1)) - s.v[0][0]), xv = (float)(256 * (iv + (cc >> 1)) - s.v[0][1]);
            const float xd = (float)s.v[0][2] - (su * xu + sv * xv);
            dmin = fminf(dmin, xd); dmax = fmaxf(dmax, xd);
        }
    }
    // the plane crosses box kd iff [dmin, dmax] meets [256 kd, 256 kd + 256]: kd in [dmin/256 - 1, dmax/256]; the float estimate is
    // good to ~1e-3 voxel, the 0.02-voxel margin keeps the range a superset and the exact tests below decide
    const int k0 = max(s.lo[2], (int)ceilf(dmin * (1.0f / 256.0f) - 1.02f)), k1 = min(s.hi[2], (int)floorf(dmax * (1.0f / 256.0f) + 0.02f));
    if (k0 > k1) return 0;
    // plane: n . (v0 - c) = base - n_w * cw ; |.| <= hs (|n_u| + |n_v| + |n_w|)
    const long long plane_r = (long long)hs * (llabs(s.n[0]) + llabs(s.n[1]) + llabs(s.n[2]));
    const long long plane_uv = s.n[0] * (long long)(s.v[0][0] - cu) + s.n[1] * (long long)(s.v[0][1] - cv);
    unsigned int frags = 0;
    float bc[3];
    bool shaded = false;
    float u = 0.f, v = 0.f;
    f3 nn = {0.f, 0.f, 0.f};
    f4 base = {0.f, 0.f, 0.f, 0.f};
    for (int kd = k0; kd <= k1; kd++)
    {
        const int cw = 256 * kd + 128;
        {
            const long long dist = plane_uv + s.n[2] * (long long)(s.v[0][2] - cw);
            if (dist > plane_r || dist < -plane_r) continue;
        }
        bool sep = false;
#pragma unroll
        for (int e = 0; e < 3; e++)
        {
            const int e1 = (e + 1) % 3, o = (e + 2) % 3;
            const int ex = s.v[e1][0] - s.v[e][0], ey = s.v[e1][1] - s.v[e][1], ez = s.v[e1][2] - s.v[e][2];
            const int pe_u = s.v[e][0] - cu, pe_v = s.v[e][1] - cv, pe_w = s.v[e][2] - cw;
            const int po_u = s.v[o][0] - cu, po_v = s.v[o][1] - cv, po_w = s.v[o][2] - cw;
            {   // axis (0, -ez, ey): the edge's two end points project to the same value
                const long long qa = wide(ey, pe_w) - wide(ez, pe_v), qo = wide(ey, po_w) - wide(ez, po_v);
                const long long r = (long long)hs * (abs(ez) + abs(ey));
                if (min(qa, qo) > r || max(qa, qo) < -r) sep = true;
            }
            {   // axis (ez, 0, -ex)
                const long long qa = wide(ez, pe_u) - wide(ex, pe_w), qo = wide(ez, po_u) - wide(ex, po_w);
                const long long r = (long long)hs * (abs(ez) + abs(ex));
                if (min(qa, qo) > r || max(qa, qo) < -r) sep = true;
            }
        }
        if (sep) continue;
        if (!shaded)
        {   // attributes depend on the column only: barycentrics of the column centre, clamped into the triangle
            shaded = true;
            const float areaf = (float)s.area;
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                bc[k] = (float)Wk[k] / areaf;
                if (bc[k] < 0.0f) bc[k] = 0.0f;
            }
            const float sum = (bc[0] + bc[1]) + bc[2];
            bc[0] = bc[0] / sum; bc[1] = bc[1] / sum; bc[2] = bc[2] / sum;
            u = (s.u[0] * bc[0] + s.u[1] * bc[1]) + s.u[2] * bc[2];
            v = (s.vv[0] * bc[0] + s.vv[1] * bc[1]) + s.vv[2] * bc[2];
            nn = {(s.nrm[0].x * bc[0] + s.nrm[1].x * bc[1]) + s.nrm[2].x * bc[2], (s.nrm[0].y * bc[0] + s.nrm[1].y * bc[1]) + s.nrm[2].y * bc[2],
                  (s.nrm[0].z * bc[0] + s.nrm[1].z * bc[1]) + s.nrm[2].z * bc[2]};
            // base colour in 8-bit units (0..255)
            if (!mat.use_textures) base = {mat.factor[0] * 255.0f, mat.factor[1] * 255.0f, mat.factor[2] * 255.0f, mat.factor[3] * 255.0f};
            else
            {
                f4 sc = {0.f, 0.f, 0.f, 0.f};
                if (mat.tex >= 0) sc = sample_trilinear(A.texs[mat.tex], u, v, s.dudx, s.dvdx, s.dudy, s.dvdy);
                base = {sc.x * mat.factor[0], sc.y * mat.factor[1], sc.z * mat.factor[2], sc.w * mat.factor[3]};
                if (base.w < 12.75f) return frags;                  // main.lua:199 (alpha < 0.05 cut-out): nothing in this column
            }
            const float ax = fabsf(nn.x), ay = fabsf(nn.y), az = fabsf(nn.z);
            const float lead = (ax >= ay && ax >= az) ? nn.x : ((ay >= az) ? nn.y : nn.z);
            if (lead < 0.0f) nn = neg3(nn);
        }
        const float r8 = floorf(dm_clamp(base.x, 0.0f, 255.0f) + 0.5f);
        const float g8 = floorf(dm_clamp(base.y, 0.0f, 255.0f) + 0.5f);
        const float b8 = floorf(dm_clamp(base.z, 0.0f, 255.0f) + 0.5f);
        const float nx8 = rintf(dm_clamp(nn.x, -1.0f, 1.0f) * 127.0f), ny8 = rintf(dm_clamp(nn.y, -1.0f, 1.0f) * 127.0f),
                    nz8 = rintf(dm_clamp(nn.z, -1.0f, 1.0f) * 127.0f);
        // back to x, y, z: axis (d+1)%3 = u, (d+2)%3 = v, d = w
        const int bx = s.d == 0 ? kd : (s.d == 1 ? iv : iu);
        const int by = s.d == 0 ? iu : (s.d == 1 ? kd : iv);
        const int bz = s.d == 0 ? iv : (s.d == 1 ? iu : kd);
        const size_t o = brick_major(bx, by, bz, A.N >> 3);
        const uint32_t owner = (uint32_t)((bx >> 3) + (by >> 3) + (bz >> 3)) & A.owner_mask;
        if (owner == A.rank)
        {   // own memory: red.global.add.v4.f32 at device scope
            atomicAdd(A.accC[owner] + o, make_float4(r8, g8, b8, 1.0f));
            atomicAdd(A.accN[owner] + o, make_float4(nx8, ny8, nz8, 0.0f));
        }
        else
        {   // another rank's brick: send the fragment.  The lanes executing this together that share a destination reserve their slots
            // with ONE local atomic (opportunistic warp aggregation) and store their records side by side.
            const unsigned lane = threadIdx.x & 31u;
            const unsigned sub = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) & (F184_FRAG_SUBQUEUES - 1u);
            uint32_t slot;
            if (A.no_aggregate) slot = atomicAdd(A.cursor + (owner * F184_FRAG_SUBQUEUES + sub) * F184_FRAG_CURSOR_STRIDE, 1u);
            else
            {
                const unsigned grp = __match_any_sync(__activemask(), owner);
                const int leader = __ffs(grp) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(A.cursor + (owner * F184_FRAG_SUBQUEUES + sub) * F184_FRAG_CURSOR_STRIDE, (uint32_t)__popc(grp));
                base = __shfl_sync(grp, base, leader);
                slot = base + (uint32_t)__popc(grp & ((1u << lane) - 1u));
            }
            if (slot < A.sub_cap)
            {   // every component is an integer: 0..255 colour, -127..127 normal — a record loses nothing
                const uint32_t rgb = (uint32_t)r8 | ((uint32_t)g8 << 8) | ((uint32_t)b8 << 16);
                const uint32_t nrm = ((uint32_t)(int)nx8 & 255u) | (((uint32_t)(int)ny8 & 255u) << 8) | (((uint32_t)(int)nz8 & 255u) << 16);
                A.peer_queue[owner][(size_t)sub * A.sub_cap + slot] = make_uint4((uint32_t)o, rgb, nrm, 0u);
                frags++;
                continue;
            }
            // queue full: reduce straight into the owner's memory over NVLink.  The CUDA memory model guarantees inter-GPU atomicity
            // at SYSTEM scope only.
            red_add_v4_sys(A.accC[owner] + o, r8, g8, b8, 1.0f);
            red_add_v4_sys(A.accN[owner] + o, nx8, ny8, nz8, 0.0f);
        }
        A.brick_flags[owner][o >> 9] = 1u;
        frags++;
    }
    return frags;
}

// ---- pass 1: one thread per triangle.  Sets the triangle up; finishes it on the spot when its projection is at
// most SMALL_COLS columns (most of a detailed mesh at any practical grid), otherwise queues it as
// ceil(columns / TASK_COLS) equal tasks for pass 2, so that a wall spanning 10^5 columns is spread over the
// whole chip instead of serialising one warp.  Queue slots and task numbers are reserved by ONE 64-bit atomic per
// warp, (entries << 40) | tasks, which keeps `first task` monotone in the entry index: pass 2 finds the triangle of
// a task by binary search, no prefix-sum pass needed.  The set-up travels with the queue entry (192 B).
__global__ void __launch_bounds__(SETUP_THREADS, 4) k_voxelize_setup(const VoxArgs A)
{
    const int lane = threadIdx.x & 31;
    static_assert(SETUP_THREADS == F184_TRIANGLE_CHUNK, "one CTA of the set-up pass = one chunk");
    const uint32_t t = A.chunks ? __ldg(A.chunks + blockIdx.x) * SETUP_THREADS + threadIdx.x : A.tri_first + blockIdx.x * SETUP_THREADS + threadIdx.x;
    TriS s;
    bool active = false;
    if (t >= A.tri_first && t < A.tri_end) active = setup_triangle(A, t, s);
    unsigned int frags = 0;
    uint32_t ntasks = 0;
    if (active)
    {
        const int bu = s.hi[0] - s.lo[0] + 1, bv = s.hi[1] - s.lo[1] + 1;
        const int ncols = bu * bv;
        if (ncols <= SMALL_COLS)
        {
            const MatDev mat = A.mats[s.mat];
            for (int col = 0; col < ncols; col++) frags += process_column(A, s, mat, s.lo[0] + col % bu, s.lo[1] + col / bu);
        }
        else ntasks = (uint32_t)((ncols + TASK_COLS - 1) / TASK_COLS);
    }
    const unsigned int big = __ballot_sync(0xffffffffu, ntasks != 0);
    if (big)
    {
        uint32_t incl = ntasks;                         // inclusive warp scan of the task counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long old = 0;
        if (lane == 0) old = atomicAdd(A.queue_state, ((unsigned long long)__popc(big) << 40) | total);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (ntasks)
        {
            const uint32_t slot = (uint32_t)(old >> 40) + __popc(big & ((1u << lane) - 1u));
            A.queue[slot] = make_uint2(t, (uint32_t)(old & 0xffffffffffull) + incl - ntasks);
            uint4 w[12];
            memcpy(w, &s, sizeof(TriS));
#pragma unroll
            for (int i = 0; i < 12; i++) A.queue_tris[(size_t)slot * 12 + i] = w[i];
        }
    }
    warp_count_add(A.frag_counter, frags);
}

// ---- pass 2: persistent warps, one task (TASK_COLS columns of one triangle) at a time.  Two steps per task: every lane
// runs the cheap projection-plane test on its columns and the survivors are compacted through shared memory (ballot +
// popc); then the expensive part — candidate depths, shading, atomics — runs on survivors only, 32 at a time.  (ncu on the
// uncompacted loop: 13 of 32 lanes active on average, because the bounding box of a slanted triangle is mostly empty.)
__global__ void __launch_bounds__(RASTER_THREADS, 4) k_voxelize_raster(const VoxArgs A)
{
    __shared__ uint8_t survivors[RASTER_THREADS / 32][TASK_COLS];
    const int lane = threadIdx.x & 31;
    uint8_t* q = survivors[threadIdx.x >> 5];
    const uint32_t gwarp = (blockIdx.x * RASTER_THREADS + threadIdx.x) >> 5, nwarps = (gridDim.x * RASTER_THREADS) >> 5;
    const unsigned long long st = *A.queue_state;
    const uint32_t n_entries = (uint32_t)(st >> 40), n_tasks = (uint32_t)(st & 0xffffffffffull);
    unsigned int frags = 0;
    for (uint32_t task = gwarp; task < n_tasks; task += nwarps)
    {
        uint32_t lo = 0, hi = n_entries;                // last entry whose first task <= task
        while (hi - lo > 1)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&A.queue[mid].y) <= task) lo = mid; else hi = mid;
        }
        const uint2 e = __ldg(A.queue + lo);
        TriS s;
        {
            uint4 w[12];
#pragma unroll
            for (int i = 0; i < 12; i++) w[i] = A.queue_tris[(size_t)lo * 12 + i];
            memcpy(&s, w, sizeof(TriS));
        }
        const MatDev mat = A.mats[s.mat];
        const int bu = s.hi[0] - s.lo[0] + 1, bv = s.hi[1] - s.lo[1] + 1;
        const int ncols = bu * bv;
        const int c0 = (int)(task - e.y) * TASK_COLS, c1 = min(ncols, c0 + TASK_COLS);
        int n = 0;
#pragma unroll 1
        for (int base = c0; base < c1; base += 32)
        {
            const int col = base + lane;
            const bool in = col < c1 && column_in_projection(s, s.lo[0] + col % bu, s.lo[1] + col / bu);
            const unsigned int m = __ballot_sync(0xffffffffu, in);
            if (in) q[n + __popc(m & ((1u << lane) - 1u))] = (uint8_t)(col - c0);
            n += __popc(m);
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32)
        {
            const int col = c0 + q[i];
            frags += process_column(A, s, mat, s.lo[0] + col % bu, s.lo[1] + col / bu);
        }
        __syncwarp();
    }
    warp_count_add(A.frag_counter, frags);
}

__global__ void k_reset_cursors(uint32_t* cursor) { cursor[threadIdx.x * F184_FRAG_CURSOR_STRIDE] = 0u; }           // <<<1, 8 * F184_FRAG_SUBQUEUES>>>

// Owner side, at the head of normalise: fetch the records the other ranks hold for this rank — coalesced 16-byte loads out of the
// senders' memory over NVLink, 512 bytes per warp instruction — and apply them: two local 16-byte reductions and the brick flag per
// record, exactly what the sender would have done to its own memory.
struct ApplyArgs
{
    const uint4* peer_queue[8];        // sender s: its region for this rank
    const uint32_t* peer_cursor[8];    // sender s: its cursors for this rank (F184_FRAG_SUBQUEUES of them)
    float4 *accC, *accN;
    uint32_t* brick_flags;
    uint32_t rank, nranks, sub_cap;
};
__global__ void __launch_bounds__(256) k_apply_fragments(const ApplyArgs P)     // grid (2, nranks * F184_FRAG_SUBQUEUES): blockIdx.y = one sub-queue
{
    const uint32_t s = blockIdx.y / F184_FRAG_SUBQUEUES, sub = blockIdx.y % F184_FRAG_SUBQUEUES;
    if (s == P.rank) return;
    const uint32_t n = min(P.peer_cursor[s][sub * F184_FRAG_CURSOR_STRIDE], P.sub_cap);
    const uint4* q = P.peer_queue[s] + (size_t)sub * P.sub_cap;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride)
    {   // four peer loads in flight per thread (each is an NVLink round trip), then the local reductions
        uint4 r[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (i0 + k * stride < n) r[k] = q[i0 + k * stride];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            if (i0 + k * stride >= n) break;
            const size_t o = r[k].x;
            atomicAdd(P.accC + o, make_float4((float)(r[k].y & 255u), (float)((r[k].y >> 8) & 255u), (float)((r[k].y >> 16) & 255u), 1.0f));
            atomicAdd(P.accN + o, make_float4((float)(int8_t)(r[k].z & 255u), (float)(int8_t)((r[k].z >> 8) & 255u), (float)(int8_t)((r[k].z >> 16) & 255u), 0.0f));
            P.brick_flags[o >> 9] = 1u;
        }
    }
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

// ---- B.2 normalise, sparse: only bricks touched this frame or last frame ---------------------------------
// pass 1: one thread per brick flag -> compact list (warp-aggregated append).  Entry = brick | touched << 31.
// `prev` bits (f184_internal.h: brick_prev): bit 0 = touched by the previous voxelize, bit 1 + s = texture set s holds the brick.
// Any of them lists the brick, so whichever copy is stale gets overwritten (with zeros if the brick is empty now).
__global__ void __launch_bounds__(256)
k_brick_compact(uint32_t* __restrict__ brick_flags, uint32_t* __restrict__ brick_prev, uint32_t* __restrict__ brick_list,
                unsigned long long* __restrict__ counters, uint32_t n_own, uint32_t NB, uint32_t G, uint32_t rank, const int32_t* __restrict__ cache_slot)
{
    const int lane = threadIdx.x & 31;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;             // j-th brick of this rank
    // own bricks of row (by, bz): bx = (rank - by - bz) mod G, + G, + 2G, ...   (owner = (bx + by + bz) % G)
    const uint32_t per_row = NB / G, row = j / per_row, by = row % NB, bz = row / NB;
    const uint32_t bx = ((rank + 2u * G - (by % G) - (bz % G)) % G) + (j % per_row) * G;
    const uint32_t bi = (bz * NB + by) * NB + bx;
    // With a static cache (f184_static_cache_capture): a cached brick always has content — it is listed every frame for inject and
    // mips — but normalise has to look at it only if this frame's or the previous frame's (dynamic) fragments touched it: history
    // bit 3, valid for this frame only, says so.  (Without a cache bit 3 is never set and nothing below differs from before.)
    uint32_t flag = 0, prev = 0, cached = 0;
    if (j < n_own)
    {
        flag = brick_flags[bi];
        prev = brick_prev[bi] & ~8u;
        if (cache_slot) cached = cache_slot[bi] >= 0 ? 1u : 0u;
    }
    if (flag | prev | cached)
    {
        const uint32_t redo = (cache_slot && (flag | (prev & 1u))) ? 8u : 0u;
        brick_flags[bi] = 0;
        brick_prev[bi] = (prev & ~1u) | (flag ? 1u : 0u) | redo;
    }
    const bool content = (flag | cached) != 0;
    const unsigned int todo = __ballot_sync(0xffffffffu, (flag | prev | cached) != 0);
    const unsigned int touched = __ballot_sync(0xffffffffu, content);
    if (!todo) return;
    unsigned long long slot = 0;
    if (lane == 0)
    {
        slot = atomicAdd(counters + F184_COUNTER_COUNT, (unsigned long long)__popc(todo));   // list cursor lives past the public counters
        if (touched) atomicAdd(counters + F184_COUNTER_BRICKS, (unsigned long long)__popc(touched));
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (flag | prev | cached) brick_list[(uint32_t)slot + __popc(todo & ((1u << lane) - 1u))] = bi | (content ? 0x80000000u : 0u);
}

// pass 2: one warp per listed brick, 16 passes of 32 voxels: 512 B coalesced accumulator reads, 32 B output runs.
// Reads the sums, writes mean albedo / unit normal as RGBA8 and re-zeroes the accumulators (= next frame's clear).
__global__ void __launch_bounds__(256)
k_normalise_n(float4* __restrict__ accC, float4* __restrict__ accN, uchar4* __restrict__ alb, char4* __restrict__ nrm,
              const uint32_t* __restrict__ brick_list, unsigned long long* __restrict__ counters, int N,
              const int32_t* __restrict__ cache_slot, const float4* __restrict__ cacheC, const float4* __restrict__ cacheN,
              const uint32_t* __restrict__ cache_occ, const uint32_t* __restrict__ brick_prev)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t count = (uint32_t)counters[F184_COUNTER_COUNT];
    const int NB = N >> 3;
    unsigned int occ = 0;
    for (uint32_t i = warp_global; i < count; i += n_warps)
    {
        const uint32_t entry = __ldg(brick_list + i);
        const uint32_t b = entry & 0x7fffffffu;
        bool is_touched = (entry >> 31) != 0;                         // the accumulators hold fragments of this frame
        int slot = -1;
        if (cache_slot)
        {   // static cache: sums = this frame's accumulators (if any fragment landed: history bit 0, just set by the compaction) + the
            // cached static sums.  A cached brick no dynamic fragment touched now or last frame keeps what the volumes hold already.
            slot = cache_slot[b];
            const uint32_t hist = brick_prev[b];
            is_touched = (hist & 1u) != 0;
            if (slot >= 0 && !(hist & 8u))
            {
                if (lane == 0) occ += cache_occ[slot];
                continue;
            }
        }
        const float4* kc = slot >= 0 ? cacheC + (size_t)slot * 512 : nullptr;
        const float4* kn = slot >= 0 ? cacheN + (size_t)slot * 512 : nullptr;
        const int bx = (b % NB) << 3, by = ((b / NB) % NB) << 3, bz = (b / (NB * NB)) << 3;
        // two halves of 8 passes: all 8 colour loads, then the 8 normal loads predicated on occupancy (no branch around
        // them, so they are all in flight together), then the arithmetic and the stores
#pragma unroll 1
        for (int half = 0; half < 2; half++)
        {
            float4 cs[8], ns[8];
#pragma unroll
            for (int k = 0; k < 8; k++)
            {
                const size_t o = (size_t)b * 512 + (half * 8 + k) * 32 + lane;
                cs[k] = is_touched ? accC[o] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (kc) { const float4 q = __ldg(kc + (half * 8 + k) * 32 + lane); cs[k].x += q.x; cs[k].y += q.y; cs[k].z += q.z; cs[k].w += q.w; }
            }
#pragma unroll
            for (int k = 0; k < 8; k++)
            {
                const size_t o = (size_t)b * 512 + (half * 8 + k) * 32 + lane;
                ns[k] = (cs[k].w > 0.0f && is_touched) ? accN[o] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (kn && cs[k].w > 0.0f) { const float4 q = __ldg(kn + (half * 8 + k) * 32 + lane); ns[k].x += q.x; ns[k].y += q.y; ns[k].z += q.z; }
            }
#pragma unroll
            for (int k = 0; k < 8; k++)
            {
                const int local = (half * 8 + k) * 32 + lane;           // (z&7)<<6 | (y&7)<<3 | (x&7)
                const size_t o = (size_t)b * 512 + local;
                uchar4 a8 = make_uchar4(0, 0, 0, 0);
                char4 n8 = make_char4(0, 0, 0, 0);
                const float4 c = cs[k];
                if (c.w > 0.0f)
                {
                    const float4 nn = ns[k];
                    a8 = make_uchar4((unsigned char)floorf(c.x / c.w + 0.5f), (unsigned char)floorf(c.y / c.w + 0.5f),
                                     (unsigned char)floorf(c.z / c.w + 0.5f), 255);
                    const float len = __fsqrt_rn((nn.x * nn.x + nn.y * nn.y) + nn.z * nn.z);
                    if (len > 0.0f)
                        n8 = make_char4((signed char)rintf(nn.x / len * 127.0f), (signed char)rintf(nn.y / len * 127.0f),
                                        (signed char)rintf(nn.z / len * 127.0f), 0);
                    if (is_touched)
                    {
                        accC[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                        accN[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    occ++;
                }
                const int x = bx + (local & 7), y = by + ((local >> 3) & 7), z = bz + (local >> 6);
                const size_t lin = ((size_t)z * N + y) * N + x;
                alb[lin] = a8;
                nrm[lin] = n8;
            }
        }
    }
    warp_count_add(counters + F184_COUNTER_OCCUPIED, occ);
}

}  // namespace

// scratch sized by the scene: the pass-2 queue and the View * Model matrices (re-allocated only when the scene grows)
int f184_voxelizer_scratch_n(f184_ctx* c)
{
    if (c->vox_queue_cap < c->n_tris)
    {   // worst case every triangle is large: (8 + 192) B per triangle
        if (c->vox_queue) cudaFree(c->vox_queue);
        c->vox_queue = nullptr; c->vox_queue_cap = 0;
        CK(c, cudaMalloc(&c->vox_queue, (8ull + sizeof(TriS)) * c->n_tris));
        c->vox_queue_cap = c->n_tris;
    }
    if (c->vm_cap < c->n_models)
    {
        if (c->vm_dev) cudaFree(c->vm_dev);
        c->vm_dev = nullptr; c->vm_cap = 0;
        CK(c, cudaMalloc(&c->vm_dev, sizeof(M4) * c->n_models));
        c->vm_cap = c->n_models;
    }
    return F184_OK;
}

int f184_voxelize_accumulate_n(f184_ctx* c, const f184_view_constants* cam)
{
    memcpy(c->last_vox_cam, cam->ViewMat, 64);
    memcpy(c->last_vox_cam + 16, cam->ProjMat, 64);
    if (c->cache_slot && memcmp(c->last_vox_cam, c->cache_cam, sizeof(c->cache_cam)) != 0)
        return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "voxelize: the static cache was captured with another voxel camera (f184_static_cache_clear, or capture again)");
    for (int s : {F184_SLOT_ACCUM_COLOR, F184_SLOT_ACCUM_NORMAL, F184_SLOT_VOX_ALBEDO, F184_SLOT_VOX_NORMAL, F184_SLOT_BRICK_FLAGS})
    {
        int rc = f184_ensure_image(c, s);
        if (rc) return rc;
    }
    const int N = (int)c->cfg.grid_n;
    const uint32_t G = c->cfg.nranks;
    if (G > 1 && ((G & (G - 1)) || G > 8 || N / 8 < (int)G)) return f184_fail(c, F184_ERR_INVALID_ARGUMENT, "nranks must be 1, 2, 4 or 8 and at most grid_n / 8");
    {
        void* dummy = nullptr;
        int rc = f184_ipc_buffer_ptr(c, F184_IPC_BRICK_LIST, &dummy);      // brick_prev / brick_list
        if (rc) return rc;
    }
    {
        int rc = f184_voxelizer_scratch_n(c);
        if (rc) return rc;
    }
    VoxArgs A{};
    for (uint32_t p = 0; p < (G ? G : 1); p++)
    {
        if (p == c->cfg.rank || G <= 1)
        {
            A.accC[p] = img_ptr<float4>(c, F184_SLOT_ACCUM_COLOR); A.accN[p] = img_ptr<float4>(c, F184_SLOT_ACCUM_NORMAL);
            A.brick_flags[p] = img_ptr<uint32_t>(c, F184_SLOT_BRICK_FLAGS);
        }
        else
        {
            if (!c->peer[p].buf[F184_IPC_ACCUM_COLOR] || !c->peer[p].buf[F184_IPC_ACCUM_NORMAL] || !c->peer[p].buf[F184_IPC_BRICK_FLAGS])
                return f184_fail(c, F184_ERR_NOT_READY, "voxelize: accumulators of rank %u were not imported (f184_ipc_import)", p);
            A.accC[p] = reinterpret_cast<float4*>(c->peer[p].buf[F184_IPC_ACCUM_COLOR]);
            A.accN[p] = reinterpret_cast<float4*>(c->peer[p].buf[F184_IPC_ACCUM_NORMAL]);
            A.brick_flags[p] = reinterpret_cast<uint32_t*>(c->peer[p].buf[F184_IPC_BRICK_FLAGS]);
        }
    }
    A.owner_mask = G > 1 ? G - 1 : 0;
    A.rank = G > 1 ? c->cfg.rank : 0;
    if (G > 1)
    {
        void* dummy = nullptr;
        int rc = f184_ipc_buffer_ptr(c, F184_IPC_FRAG_QUEUE, &dummy);
        if (rc) return rc;
        for (uint32_t p = 0; p < G; p++)
            if (p != c->cfg.rank) A.peer_queue[p] = c->frag_queue + (size_t)p * c->frag_cap;      // own memory: the region for rank p
        A.cursor = c->frag_cursor;
        A.sub_cap = c->frag_cap / F184_FRAG_SUBQUEUES;
        static const bool no_agg = [] { const char* e = getenv("F184_FRAG_NO_AGG"); return e && atoi(e) != 0; }();
        A.no_aggregate = no_agg ? 1u : 0u;
    }
    c->voxel_h = f184_voxel_h(cam->ProjMat, cam->ViewMat, c->cfg.grid_n);
    M4 View;
    memcpy(View.m, cam->ViewMat, 64);
    k_view_model_n<<<(c->n_models * 16 + 127) / 128, 128, 0, c->stream>>>(View, c->model_mats, c->vm_dev, c->n_models);
    CK_LAUNCH(c);

    const uint32_t first = c->tri_first < c->n_tris ? c->tri_first : c->n_tris;
    const uint64_t end64 = (uint64_t)first + c->tri_count;
    const uint32_t end = end64 < c->n_tris ? (uint32_t)end64 : c->n_tris;

    int rc = f184_stage_begin(c, F184_STAGE_VOXELIZE);
    if (rc) return rc;
    // device words after the public counters: [COUNT] brick-list cursor, [COUNT+1] voxelizer queue state, [COUNT+2] mip tail ticket
    if ((rc = f184_zero_counters(c, (1u << F184_COUNTER_FRAGMENTS) | (1u << (F184_COUNTER_COUNT + 1))))) return rc;
    if (G > 1 && c->frag_sent_applied)
    {   // the owners have applied (or are about to apply, behind the barrier that followed) what this rank sent last time: start the
        // queues over.  Accumulations that follow each other WITHOUT a barrier in between keep appending (partial volumes add up).
        k_reset_cursors<<<1, 8 * F184_FRAG_SUBQUEUES, 0, c->stream>>>(c->frag_cursor);
        CK_LAUNCH(c);
        c->frag_sent_applied = false;
    }
    c->frag_pending = true;
    if (end > first)
    {
        A.pos = c->pos; A.nrm = c->nrm; A.uv = c->uv; A.idx = c->idx; A.tri_mat = c->tri_mat; A.tri_model = c->tri_model;
        A.model_mats = c->model_mats; A.vm_mats = c->vm_dev;
        memcpy(A.Proj.m, cam->ProjMat, 64);
        A.texs = c->tex_dev; A.mats = c->mat_dev;
        A.tri_first = first; A.tri_end = end; A.N = N;
        A.frag_counter = c->counters_dev + F184_COUNTER_FRAGMENTS;
        A.queue_state = c->counters_dev + F184_COUNTER_COUNT + 1;
        A.queue_tris = reinterpret_cast<uint4*>(c->vox_queue);                      // 192 B records first (16-byte aligned)
        A.queue = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(c->vox_queue) + sizeof(TriS) * (size_t)c->vox_queue_cap);
        const uint32_t tris = end - first;
        A.chunks = c->n_chunks ? c->chunk_list : nullptr;
        k_voxelize_setup<<<c->n_chunks ? c->n_chunks : (tris + SETUP_THREADS - 1) / SETUP_THREADS, SETUP_THREADS, 0, c->stream>>>(A);
        CK_LAUNCH(c);
        k_voxelize_raster<<<148 * 4 * 4, RASTER_THREADS, 0, c->stream>>>(A);
        CK_LAUNCH(c);
    }
    return f184_stage_end(c, F184_STAGE_VOXELIZE);
}

// The fragments the other ranks hold for this rank (their kernels finished before the barrier this call follows): pulled and reduced
// into the own accumulators.  Once per accumulation.
static int apply_pending_fragments(f184_ctx* c)
{
    const uint32_t G = c->cfg.nranks > 1 ? c->cfg.nranks : 1;
    int rc = F184_OK;
    if (G > 1 && c->frag_queue && c->frag_pending)
    {
        if ((rc = f184_stage_begin(c, F184_STAGE_APPLY))) return rc;   // the fragments the other ranks hold for this rank (their kernels finished before the barrier this call follows)
        ApplyArgs P{};
        for (uint32_t p = 0; p < G; p++)
        {
            if (p == c->cfg.rank) continue;
            if (!c->peer[p].buf[F184_IPC_FRAG_QUEUE] || !c->peer[p].buf[F184_IPC_FRAG_COUNTS])
                return f184_fail(c, F184_ERR_NOT_READY, "normalise: fragment queue of rank %u was not imported (f184_ipc_import)", p);
            P.peer_queue[p] = reinterpret_cast<const uint4*>(c->peer[p].buf[F184_IPC_FRAG_QUEUE]) + (size_t)c->cfg.rank * c->frag_cap;
            P.peer_cursor[p] = reinterpret_cast<const uint32_t*>(c->peer[p].buf[F184_IPC_FRAG_COUNTS]) + (size_t)c->cfg.rank * F184_FRAG_SUBQUEUES * F184_FRAG_CURSOR_STRIDE;
        }
        P.accC = img_ptr<float4>(c, F184_SLOT_ACCUM_COLOR); P.accN = img_ptr<float4>(c, F184_SLOT_ACCUM_NORMAL);
        P.brick_flags = img_ptr<uint32_t>(c, F184_SLOT_BRICK_FLAGS);
        P.rank = c->cfg.rank; P.nranks = G; P.sub_cap = c->frag_cap / F184_FRAG_SUBQUEUES;
        k_apply_fragments<<<dim3(2, G * F184_FRAG_SUBQUEUES), 256, 0, c->stream>>>(P);
        CK_LAUNCH(c);
        c->frag_pending = false;          // applied once: a second normalise without a new accumulation must not add them again
        if ((rc = f184_stage_end(c, F184_STAGE_APPLY))) return rc;
    }
    return F184_OK;
}

// Normalise this rank's bricks (the whole volume on one GPU).
int f184_normalise_n(f184_ctx* c)
{
    const int N = (int)c->cfg.grid_n;
    const uint32_t NB = (uint32_t)N / 8, G = c->cfg.nranks > 1 ? c->cfg.nranks : 1;
    const uint32_t n_own = (NB / G) * NB * NB;
    int rc = apply_pending_fragments(c);
    if (rc) return rc;
    rc = f184_stage_begin(c, F184_STAGE_NORMALISE);
    if (rc) return rc;
    if ((rc = f184_zero_counters(c, (1u << F184_COUNTER_OCCUPIED) | (1u << F184_COUNTER_BRICKS) | (1u << F184_COUNTER_COUNT)))) return rc;   // COUNT = the list cursor
    k_brick_compact<<<(n_own + 255) / 256, 256, 0, c->stream>>>(img_ptr<uint32_t>(c, F184_SLOT_BRICK_FLAGS), c->brick_prev, c->brick_list,
                                                              c->counters_dev, n_own, NB, G, c->cfg.rank % G, c->cache_slot);
    CK_LAUNCH(c);
    k_normalise_n<<<148 * 16, 128, 0, c->stream>>>(img_ptr<float4>(c, F184_SLOT_ACCUM_COLOR), img_ptr<float4>(c, F184_SLOT_ACCUM_NORMAL),
                                                  img_ptr<uchar4>(c, F184_SLOT_VOX_ALBEDO), img_ptr<char4>(c, F184_SLOT_VOX_NORMAL),
                                                  c->brick_list, c->counters_dev, N, c->cache_slot, c->cacheC, c->cacheN, c->cache_occ, c->brick_prev);
    CK_LAUNCH(c);
    return f184_stage_end(c, F184_STAGE_NORMALISE);
}

// ---- static / dynamic split ---------------------------------------------------------------------------
// Geometry that never moves need not be rasterised every frame.  After an accumulation of the STATIC triangles (and, on several
// ranks, the barrier that completes it) f184_static_cache_capture moves the accumulators of every touched brick of this rank into a
// sparse cache — 16 KB per brick — instead of normalising them.  From then on the caller accumulates only the DYNAMIC triangles;
// normalise adds a brick's cached sums to whatever the frame's fragments left in its accumulators.  Sums of integer-valued floats
// are exact in any order, so the volumes are bit-identical to voxelizing everything every frame.  A cached brick that no dynamic
// fragment touches (now or in the previous frame) is not read at all: the linear volumes hold its values already.
static __global__ void __launch_bounds__(256) k_cache_count(const uint32_t* __restrict__ flags, uint32_t n_bricks, unsigned int* __restrict__ out)
{
    unsigned int n = 0;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_bricks; b += gridDim.x * blockDim.x) n += flags[b] ? 1u : 0u;
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(out, n);
}
static __global__ void __launch_bounds__(256)
k_cache_capture(float4* __restrict__ accC, float4* __restrict__ accN, uint32_t* __restrict__ flags, uint32_t* __restrict__ brick_prev, uint32_t n_bricks,
                float4* __restrict__ cacheC, float4* __restrict__ cacheN, int32_t* __restrict__ cache_slot, uint32_t* __restrict__ cache_occ,
                unsigned int* __restrict__ cursor)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t b = warp_global; b < n_bricks; b += n_warps)
    {
        if (!flags[b]) continue;                                      // (warp-uniform: every lane reads the same word)
        unsigned int slot = 0;
        if (lane == 0) slot = atomicAdd(cursor, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        unsigned int occ = 0;
#pragma unroll 4
        for (int k = 0; k < 16; k++)
        {
            const size_t o = (size_t)b * 512 + k * 32 + lane;
            const float4 cc = accC[o], nn = accN[o];
            cacheC[(size_t)slot * 512 + k * 32 + lane] = cc;
            cacheN[(size_t)slot * 512 + k * 32 + lane] = nn;
            if (cc.w > 0.0f)
            {
                occ++;
                accC[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                accN[o] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int o = 16; o; o >>= 1) occ += __shfl_xor_sync(0xffffffffu, occ, o);
        if (lane == 0)
        {
            cache_slot[b] = (int32_t)slot;
            cache_occ[slot] = occ;
            flags[b] = 0;
            brick_prev[b] |= 1u;                                      // "touched by the previous voxelize": the next normalise writes the brick's static values
        }
    }
}

static void cache_free(f184_ctx* c)
{
    for (void* p : {(void*)c->cacheC, (void*)c->cacheN, (void*)c->cache_slot, (void*)c->cache_occ})
        if (p) cudaFree(p);
    c->cacheC = c->cacheN = nullptr; c->cache_slot = nullptr; c->cache_occ = nullptr;
    c->n_cached = 0; c->cache_fragments = 0;
}

int f184_static_cache_capture_n(f184_ctx* c)
{
    const uint32_t N = c->cfg.grid_n, n_bricks = (N / 8) * (N / 8) * (N / 8);
    int rc = apply_pending_fragments(c);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    cache_free(c);
    uint32_t* flags = img_ptr<uint32_t>(c, F184_SLOT_BRICK_FLAGS);
    unsigned int* counter = nullptr;
    CK(c, cudaMalloc(&counter, 2 * sizeof(unsigned int)));
    CK(c, cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned int), c->stream));
    k_cache_count<<<148 * 4, 256, 0, c->stream>>>(flags, n_bricks, counter);
    CK_LAUNCH(c);
    unsigned int n = 0;
    CK(c, cudaMemcpyAsync(&n, counter, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaMalloc(&c->cache_slot, sizeof(int32_t) * (size_t)n_bricks));
    if ((rc = f184_fill_async(c, c->cache_slot, 0xffffffffu, sizeof(int32_t) * (size_t)n_bricks, c->stream))) return rc;
    const size_t slots = n ? n : 1;
    CK(c, cudaMalloc(&c->cacheC, sizeof(float4) * 512 * slots));
    CK(c, cudaMalloc(&c->cacheN, sizeof(float4) * 512 * slots));
    CK(c, cudaMalloc(&c->cache_occ, sizeof(uint32_t) * slots));
    k_cache_capture<<<148 * 8, 256, 0, c->stream>>>(img_ptr<float4>(c, F184_SLOT_ACCUM_COLOR), img_ptr<float4>(c, F184_SLOT_ACCUM_NORMAL), flags, c->brick_prev, n_bricks,
                                                   c->cacheC, c->cacheN, c->cache_slot, c->cache_occ, counter + 1);
    CK_LAUNCH(c);
    unsigned long long frags = 0;
    CK(c, cudaMemcpyAsync(&frags, c->counters_dev + F184_COUNTER_FRAGMENTS, sizeof(frags), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(counter);
    c->n_cached = n;
    c->cache_fragments = frags;
    memcpy(c->cache_cam, c->last_vox_cam, sizeof(c->cache_cam));
    return F184_OK;
}

int f184_static_cache_clear_n(f184_ctx* c)
{
    CK(c, cudaStreamSynchronize(c->stream));
    cache_free(c);            // the bricks it held carry their history bits: the next frames list them and write what is left (zeros, or the dynamic part)
    return F184_OK;
}
