"""The CPU oracle, north-star mode, against independent pure-Python restatements on small cases:
exact-integer conservative voxelization (triangle/closed-box SAT with Python ints), the integer 2x2x2 six-direction
reduction, normalise means, and analytic properties of the cone tracer."""
from fractions import Fraction

import numpy as np
import pytest

import cpu_helpers as Hc
from final184_b200 import api as A
from final184_b200 import scene as S
from final184_b200.fixture import frame_inputs


# ---- independent conservative voxelizer: Python ints, brute force over all N^3 voxels -------------------------
def snap_vertices(tris_world, n):
    """Voxel space of the tracer (indirect.frag:141-143) via the survey's closed form of the voxel camera:
    ndc = (wx/16, -wz/16, (15-wy)/19); v = ((x*.5+.5)N, (y*.5+.5)N, z*N); snapped to 1/256 voxel (exact rationals)."""
    out = []
    for tri in tris_world:
        vs = []
        for wx, wy, wz in tri:
            wx, wy, wz = Fraction(wx).limit_denominator(1 << 20), Fraction(wy).limit_denominator(1 << 20), Fraction(wz).limit_denominator(1 << 20)
            v = ((wx / 16 / 2 + Fraction(1, 2)) * n, (-wz / 16 / 2 + Fraction(1, 2)) * n, (15 - wy) / 19 * n)
            sn = []
            for c in v:
                s = c * 256
                r = int(s + Fraction(1, 2)) if s >= 0 else -int(-s + Fraction(1, 2))
                assert abs(s - r) < Fraction(49, 100), "test vertex too close to a rounding boundary for float32"
                sn.append(r)
            vs.append(sn)
        out.append(vs)
    return out


def sat_overlap(v, box):
    p = [[v[k][a] - (256 * box[a] + 128) for a in range(3)] for k in range(3)]
    hs = 128
    for a in range(3):
        if min(p[0][a], p[1][a], p[2][a]) > hs or max(p[0][a], p[1][a], p[2][a]) < -hs:
            return False
    e1 = [p[1][a] - p[0][a] for a in range(3)]; e2 = [p[2][a] - p[0][a] for a in range(3)]
    n = [e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]]
    d = sum(n[a] * p[0][a] for a in range(3))
    if abs(d) > hs * sum(abs(c) for c in n):
        return False
    for e in range(3):
        ed = [p[(e + 1) % 3][a] - p[e][a] for a in range(3)]
        for ax in ([0, -ed[2], ed[1]], [ed[2], 0, -ed[0]], [-ed[1], ed[0], 0]):
            q = [sum(ax[a] * p[k][a] for a in range(3)) for k in range(3)]
            r = hs * sum(abs(c) for c in ax)
            if min(q) > r or max(q) < -r:
                return False
    return True


def python_voxelize(tris_world, n):
    counts = np.zeros((n, n, n), np.int64)
    for v in snap_vertices(tris_world, n):
        e1 = [v[1][a] - v[0][a] for a in range(3)]; e2 = [v[2][a] - v[0][a] for a in range(3)]
        nn = [e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]]
        if nn == [0, 0, 0]:
            continue
        lo = [max(0, (min(v[k][a] for k in range(3)) - 1) // 256) for a in range(3)]
        hi = [min(n - 1, max(v[k][a] for k in range(3)) // 256) for a in range(3)]
        for z in range(lo[2], hi[2] + 1):
            for y in range(lo[1], hi[1] + 1):
                for x in range(lo[0], hi[0] + 1):
                    if sat_overlap(v, (x, y, z)):
                        counts[z, y, x] += 1
    return counts


TRIS = [
    [(-3.3, 1.7, -2.1), (4.9, 2.3, -1.2), (0.4, 6.1, 3.7)],          # slanted, large
    [(1.0, 0.5, 1.0), (1.3, 0.55, 1.1), (1.1, 0.7, 1.4)],            # smaller than a voxel
    [(-6.2, 0.3, -5.1), (6.7, 0.3, -5.1), (-6.2, 0.3, 5.9)],         # axis-aligned floor piece
    [(2.2, 0.1, -7.3), (2.2, 9.3, -7.3), (2.25, 4.0, 6.6)],          # thin sliver, x-dominant
    [(-20.0, 3.0, 0.0), (-14.5, 3.1, 1.0), (-15.0, 5.0, -2.0)],      # partly outside the grid
]


@pytest.mark.parametrize("n", [16, 32])
def test_conservative_voxelization_equals_python_int_sat(oracle_lib, n):
    o = A.VoxelGI(grid_n=n, width=16, height=16, mode=A.MODE_NORTHSTAR, lib=oracle_lib)
    o.upload_scene(Hc.tri_scene(TRIS))
    o.voxelize(S.fixture_constants("voxel"))
    want = python_voxelize(TRIS, n)
    acc = o.readback(A.SLOT_ACCUM_COLOR)
    assert np.array_equal(acc[..., 3].astype(np.int64), want), "fragment count per voxel"
    assert o.counter(A.COUNTER_FRAGMENTS) == want.sum() and o.counter(A.COUNTER_OCCUPIED) == (want > 0).sum()
    alb = o.readback(A.SLOT_VOX_ALBEDO)
    assert np.array_equal(alb[..., 3] == 255, want > 0)
    assert (alb[want > 0][:, :3] == 255).all()                       # white texture: mean colour is exactly white
    # conservative: a superset of what centre sampling would give, and no voxel the triangle does not touch
    assert (want > 0).sum() > 0


def test_normalise_is_the_mean_and_the_normal_is_unit(oracle_lib):
    """Two coincident floors, one dark one bright: mean colour within 1/255 (north_star tolerance), unit normal."""
    tex = np.full((4, 4, 4), 255, np.uint8)
    sc = Hc.quad_scene([((-4, 0.3, -4), (8, 0, 0), (0, 0, 8), (0, 1, 0)), ((-4, 0.3, -4), (8, 0, 0), (0, 0, 8), (0, 1, 0))], tex=tex)
    sc.mat_tex = np.array([0, 0], np.int32); sc.mat_factor = np.array([[0.2, 0.4, 0.6, 1.0], [1.0, 0.8, 0.2, 1.0]], np.float32)
    sc.tri_mat = np.array([0, 0, 1, 1], np.uint16)
    o = A.VoxelGI(grid_n=32, width=16, height=16, mode=A.MODE_NORTHSTAR, lib=oracle_lib)
    o.upload_scene(sc)
    o.voxelize(S.fixture_constants("voxel"))
    alb, nrm = o.readback(A.SLOT_VOX_ALBEDO), o.readback(A.SLOT_VOX_NORMAL)
    occ = alb[..., 3] == 255
    assert occ.sum() > 64
    mean = (np.array([0.2, 0.4, 0.6]) + np.array([1.0, 0.8, 0.2])) / 2 * 255
    assert np.abs(alb[occ][:, :3].astype(np.float64) - mean).max() <= 1.0
    assert (nrm[occ][:, 1] == 127).all() and (nrm[occ][:, 0] == 0).all() and (nrm[occ][:, 2] == 0).all()


def test_partial_accumulators_add_up_exactly(oracle_lib, proc_scene):
    """The exchange step of the multi-GPU schedule: sums of integer-valued fp32 addends are order independent, so
    accumulators of disjoint triangle ranges add to the accumulators of the whole scene bit for bit."""
    cam = S.fixture_constants("voxel")
    o = A.VoxelGI(grid_n=32, width=16, height=16, mode=A.MODE_NORTHSTAR, lib=oracle_lib)
    o.upload_scene(proc_scene)
    o.voxelize(cam)
    full = [o.readback(s).copy() for s in (A.SLOT_ACCUM_COLOR, A.SLOT_ACCUM_NORMAL)]
    parts = [np.zeros_like(full[0]), np.zeros_like(full[1])]
    T = proc_scene.n_tris
    for first, count in ((0, T // 3), (T // 3, T // 3), (2 * (T // 3), T - 2 * (T // 3))):
        o.set_triangle_range(first, count); o.voxelize(cam)
        parts[0] += o.readback(A.SLOT_ACCUM_COLOR); parts[1] += o.readback(A.SLOT_ACCUM_NORMAL)
    assert np.array_equal(parts[0], full[0]) and np.array_equal(parts[1], full[1])


# ---- six-direction mips: integer reduction restated in Python -------------------------------------------------
def python_mips(level0):
    n = level0.shape[0]
    src = [level0.astype(np.int64)] * 6
    out = []
    while n > 1:
        m = n // 2
        lv = []
        for d in range(6):
            axis, neg = d >> 1, d & 1
            s = src[d]                                             # [z, y, x, c]
            ax_np = 2 - axis                                       # x is the last spatial axis
            first = np.take(s, range(neg, n, 2), axis=ax_np)
            second = np.take(s, range(1 - neg, n, 2), axis=ax_np)
            col = first * 255 + (255 - first[..., 3:4]) * second   # front over back, premultiplied, in 255^2 units
            others = [a for a in (0, 1, 2) if a != ax_np]
            for a in others:                                       # sum the 4 columns of the block
                col = col.reshape(*col.shape[:a], col.shape[a] // 2, 2, *col.shape[a + 1:]).sum(a + 1)
            lv.append(((col + 510) // 1020).astype(np.uint8))
        out.append(lv)
        src = [l.astype(np.int64) for l in lv]
        n = m
    return out


def test_mips_match_python_integer_reduction(oracle_lib):
    rng = np.random.default_rng(5)
    n = 16
    rad = rng.integers(0, 256, (n, n, n, 4), dtype=np.uint8)
    rad[rng.uniform(size=(n, n, n)) < 0.5] = 0
    rad[..., :3] = np.minimum(rad[..., :3], rad[..., 3:4])         # premultiplied
    o = A.VoxelGI(grid_n=n, width=16, height=16, mode=A.MODE_NORTHSTAR, lib=oracle_lib)
    o.upload(A.SLOT_RADIANCE, rad)
    o.build_mips()
    got = o.readback(A.SLOT_MIPS).reshape(-1, 4)
    off = 0
    for lv in python_mips(rad):
        m = lv[0].shape[0]
        for d in range(6):
            assert np.array_equal(got[off:off + m ** 3].reshape(m, m, m, 4), lv[d]), (m, d)
            off += m ** 3
    assert off == len(got)


def test_mip_direction_semantics(oracle_lib):
    """An opaque red voxel in FRONT of a green one along +x: the +x volume sees red, the -x volume sees green."""
    n = 8
    rad = np.zeros((n, n, n, 4), np.uint8)
    rad[0, 0, 0] = (255, 0, 0, 255); rad[0, 0, 1] = (0, 255, 0, 255)        # x = 0 red, x = 1 green
    o = A.VoxelGI(grid_n=n, width=16, height=16, mode=A.MODE_NORTHSTAR, lib=oracle_lib)
    o.upload(A.SLOT_RADIANCE, rad); o.build_mips()
    l1 = o.readback(A.SLOT_MIPS).reshape(-1, 4)[:6 * 64].reshape(6, 4, 4, 4, 4)
    assert tuple(l1[0, 0, 0, 0]) == (64, 0, 0, 64)      # +x: red hides green; 1 of 4 columns -> 255/4 rounded
    assert tuple(l1[1, 0, 0, 0]) == (0, 64, 0, 64)      # -x: green hides red
    assert tuple(l1[2, 0, 0, 0]) == (64, 64, 0, 128)    # +y: two separate columns, both visible


# ---- cone tracer properties -----------------------------------------------------------------------------------
def test_empty_volume_returns_the_sky_term(oracle_lib, proc_scene, cams):
    """No occluder: every cone ends with A = 0 and adds (0.7,0.8,1)*0.4 (indirect.frag:180); diffuse weights sum to 1,
    the specular cone adds F in [0.04, 1] more."""
    W, H = 64, 36
    o = A.VoxelGI(grid_n=32, width=W, height=H, mode=A.MODE_NORTHSTAR, shadow_res=64, lib=oracle_lib)
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], W, H, 64, 0, cache=False)
    for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material")):
        o.upload(slot, fi[key])
    o.upload(A.SLOT_RADIANCE, np.zeros((32, 32, 32, 4), np.uint8)); o.build_mips()
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    o.trace_indirect(k)
    img = o.readback(A.SLOT_INDIRECT_OUT).astype(np.float32)
    geo = fi["depth"] < 1.0
    sky = np.array([0.28, 0.32, 0.4], np.float32)
    ratio = img[geo][:, :3] / sky
    assert np.abs(ratio - ratio[:, :1]).max() < 5e-3            # same multiple of the sky colour in every channel
    assert ratio.min() >= 1.0 - 2e-3 and ratio.max() <= 2.0 + 2e-3
    assert (img[~geo][:, :3] == 0).all()
    assert o.counter(A.COUNTER_MARCH_STEPS) > 0


def test_trace_bands_compose(oracle_lib, proc_scene, cams):
    W, H, n = 64, 36, 32
    o = A.VoxelGI(grid_n=n, width=W, height=H, mode=A.MODE_NORTHSTAR, shadow_res=128, lib=oracle_lib)
    o.upload_scene(proc_scene)
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], W, H, 128, 0, cache=False)
    for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material"), (A.SLOT_SHADOW, "shadow")):
        o.upload(slot, fi[key])
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    o.voxelize(cams["voxel"]); o.inject(k); o.build_mips(); o.trace_indirect(k)
    full = o.readback(A.SLOT_INDIRECT_OUT).copy()
    assert full[..., :3].astype(np.float32).mean() > 1e-2
    out = np.zeros_like(full)
    for y0, y1 in ((0, 8), (8, 24), (24, H)):
        o.set_trace_rows(y0, y1); o.trace_indirect(k)
        out[y0:y1] = o.readback(A.SLOT_INDIRECT_OUT)[y0:y1]
    assert np.array_equal(out.view(np.uint16), full.view(np.uint16))
    # interleaved 8-row tiles (the multi-GPU trace split) compose as well
    o.set_trace_rows(0, 0xffffffff)
    out = np.zeros_like(full)
    for first in range(3):
        o.set_trace_tiles(first, 3); o.trace_indirect(k)
        m = (np.arange(H) // 8) % 3 == first
        out[m] = o.readback(A.SLOT_INDIRECT_OUT)[m]
    assert np.array_equal(out.view(np.uint16), full.view(np.uint16))


@pytest.mark.parametrize("spec_b", [False, True], ids=["amended", "appendix_b"])
def test_cone_trace_equals_an_independent_float64_restatement(oracle_lib, proc_scene, cams, spec_b):
    """DESIGN.md §3 B.5 written out again in plain Python / float64, straight from the prose — tangent frame, 6 + 1 cones,
    nearest-level direction-weighted trilinear sampling with zero border, front-to-back compositing, one sample per voxel of
    the level, sky term, Schlick-weighted specular, the temporal blend against an empty history — over the volumes the oracle
    built.  The oracle works in fp32 with 8-bit filter weights, so agreement is asserted per image (5e-3 rel. L2) and on the
    cone-sample count (a sample is gained or lost only where a float comparison sits on its edge).
    spec_b: the same for SURVEY.md Appendix B.5 as written (F184_FLAG_SPEC_APPENDIX_B): mip-linear sampling — a blend of the isotropic
    level 0 and the directional level 1 below lod 1, of two adjacent levels of the chain above — and half-diameter steps."""
    n, w, h, sh = 32, 24, 16, 128
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], w, h, sh, 0, cache=False)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], w, h, 0, True)
    o = A.VoxelGI(grid_n=n, width=w, height=h, mode=A.MODE_NORTHSTAR, shadow_res=sh, lib=oracle_lib, flags=A.FLAG_SPEC_APPENDIX_B if spec_b else 0)
    o.upload_scene(proc_scene)
    for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_SHADOW, "shadow"), (A.SLOT_MATERIAL, "material")):
        o.upload(slot, fi[key])
    o.voxelize(cams["voxel"]); o.inject(k); o.build_mips(); o.trace_indirect(k)
    got = o.readback(A.SLOT_INDIRECT_OUT).astype(np.float64)
    samples_oracle = o.counter(A.COUNTER_MARCH_STEPS)
    rad = o.readback(A.SLOT_RADIANCE).reshape(n, n, n, 4).astype(np.float64) / 255.0
    mips = o.readback(A.SLOT_MIPS).reshape(-1, 4).astype(np.float64) / 255.0
    levels, off, m = [], 0, n // 2
    while m >= 1:
        levels.append([mips[off + d * m ** 3: off + (d + 1) * m ** 3].reshape(m, m, m, 4) for d in range(6)])
        off += 6 * m ** 3
        m //= 2
    o.close()

    def col_major(c16):
        return np.array(list(c16), np.float64).reshape(4, 4).T            # upload order is column-major
    inv_proj, inv_mv = col_major(k.view.InvProj), col_major(k.ext.InvModelView)
    w2v = col_major(k.ext.VoxelProj) @ col_major(k.ext.VoxelView)
    hvox = np.linalg.norm((np.linalg.inv(w2v) @ np.array([2.0 / n, 0, 0, 0]))[:3])
    exposure = max(k.sun.luminance)
    cam = inv_mv[:3, 3]

    def trilinear(vol, q):
        m_ = vol.shape[0]
        p = q * m_ - 0.5
        p0 = np.floor(p)
        f = p - p0
        out = np.zeros(4)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    i = (int(p0[0]) + dx, int(p0[1]) + dy, int(p0[2]) + dz)
                    if min(i) < 0 or max(i) >= m_:
                        continue                                       # zero border
                    wgt = (f[0] if dx else 1 - f[0]) * (f[1] if dy else 1 - f[1]) * (f[2] if dz else 1 - f[2])
                    out += wgt * vol[i[2], i[1], i[0]]
        return out

    n_samples = 0

    def cone(origin, d, tan_half):
        nonlocal n_samples
        dv = (w2v[:3, :3] @ d) * np.array([0.5, 0.5, 1.0])
        du = dv / np.linalg.norm(dv)
        t, alpha, acc = 2 * hvox, 0.0, np.zeros(3)
        while alpha < 0.95 and t < 32.0:
            diam = max(hvox, 2 * t * tan_half)
            q4 = w2v @ np.append(origin + d * t, 1.0)
            q = np.array([q4[0] * 0.5 + 0.5, q4[1] * 0.5 + 0.5, q4[2]])
            if (q < 0).any() or (q > 1).any():
                break
            n_samples += 1
            lod = np.log2(diam / hvox)
            directional = lambda lv: sum(du[a] ** 2 * trilinear(lv[2 * a + (1 if du[a] < 0 else 0)], q) for a in range(3) if du[a] != 0)
            if spec_b:
                if lod < 1.0:
                    s = (1 - lod) * trilinear(rad, q) + lod * directional(levels[0])
                else:
                    lp = min(lod - 1.0, len(levels) - 1.0)
                    l0_ = int(np.floor(lp))
                    f_ = lp - l0_
                    s = directional(levels[l0_]) if f_ == 0 else (1 - f_) * directional(levels[l0_]) + f_ * directional(levels[l0_ + 1])
            else:
                level = int(np.floor(lod + 0.5))
                s = trilinear(rad, q) if level <= 0 else directional(levels[min(level, len(levels)) - 1])
            acc += (1 - alpha) * s[:3]
            alpha += (1 - alpha) * s[3]
            t += 0.5 * diam if spec_b else diam
        return acc * exposure + np.array([0.7, 0.8, 1.0]) * 0.4 * max(0.0, 1 - alpha)

    diffuse = [(0.0, 0.0, 1.0, 0.25)] + [(np.sin(np.pi / 3) * np.cos(2 * np.pi * j / 5), np.sin(np.pi / 3) * np.sin(2 * np.pi * j / 5), 0.5, 0.15) for j in range(5)]
    want = np.zeros((h, w, 4))
    for y in range(h):
        for x in range(w):
            u, v = (x + 0.5) / w, (y + 0.5) / h
            depth = float(fi["depth"][y, x])
            cp = inv_proj @ np.array([u * 2 - 1, v * 2 - 1, depth, 1.0])
            cs = cp[:3] / cp[3]
            want[y, x, 3] = -cs[2]
            if depth >= 1.0:
                continue
            wpos = (inv_mv @ np.append(cs, 1.0))[:3]
            nrm = fi["normals"][y, x, :3].astype(np.float64) / 65535.0 * 2 - 1
            z = inv_mv[:3, :3] @ (nrm / np.linalg.norm(nrm))
            hh = z.copy()
            a = np.abs(hh)
            if a[0] <= a[1] and a[0] <= a[2]: hh[0] = 1.0
            elif a[1] <= a[0] and a[1] <= a[2]: hh[1] = 1.0
            else: hh[2] = 1.0
            z = z / np.linalg.norm(z)
            ty = np.cross(hh, z); ty /= np.linalg.norm(ty)
            tx = np.cross(z, ty); tx /= np.linalg.norm(tx)
            origin = wpos + z * hvox
            ind = np.zeros(3)
            for d0, d1, d2, wgt in diffuse:
                ind += wgt * cone(origin, tx * d0 + ty * d1 + z * d2, np.tan(np.pi / 6))
            rough = float(fi["material"][y, x, 1]) / 255.0
            view = (wpos - cam) / np.linalg.norm(wpos - cam)
            refl = view - 2 * np.dot(z, view) * z
            if np.dot(refl, z) > 0:
                fres = 0.04 + 0.96 * (1 - max(-np.dot(z, view), 0.0)) ** 5
                ind += fres * cone(origin, refl, min(max(rough * rough, 0.02), 0.6))
            # temporal blend (indirect.frag:225-240) against the cleared history: the reprojection lands on the pixel itself
            tt = min(max(1 - abs(0.0 + cs[2]), 0.0), 1.0)
            bw = 0.95 * tt * tt * (3 - 2 * tt)
            want[y, x, :3] = np.clip(ind * (1 - bw), 0.0, 16.0)
    lit = want[..., :3].sum(-1) > 0
    assert lit.mean() > 0.5
    err = np.linalg.norm(got[..., :3] - want[..., :3]) / np.linalg.norm(want[..., :3])
    print(f"cone trace: oracle vs float64 restatement rel. L2 {err:.2e}, cone-samples {samples_oracle} vs {n_samples}")
    assert err < 5e-3, err
    assert np.allclose(got[..., 3], want[..., 3], rtol=2e-3, atol=1e-3)
    assert abs(n_samples - samples_oracle) <= 0.01 * samples_oracle, (n_samples, samples_oracle)


def test_injection_equals_an_independent_float64_restatement(oracle_lib, proc_scene, cams):
    """DESIGN.md §3 B.3 again in numpy float64, vectorised over the occupied voxels: world centre through the inverse voxel
    camera, decoded normal, albedo^2.2, the shadow test of indirect.frag:161-165, two-sided Lambert, RGBA8 quantisation by the
    exposure.  fp32 vs fp64 may flip a value that sits on a rounding edge (+-1 LSB) or a shadow comparison on its edge; both
    are counted and bounded."""
    n, sh = 64, 256
    fi = frame_inputs(proc_scene, cams["main"], cams["shadow"], 16, 16, sh, 0, cache=False)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], 16, 16, 0, True)
    o = A.VoxelGI(grid_n=n, width=16, height=16, mode=A.MODE_NORTHSTAR, shadow_res=sh, lib=oracle_lib)
    o.upload_scene(proc_scene)
    o.upload(A.SLOT_SHADOW, fi["shadow"])
    o.voxelize(cams["voxel"]); o.inject(k)
    alb = o.readback(A.SLOT_VOX_ALBEDO).reshape(n, n, n, 4)
    nrm = o.readback(A.SLOT_VOX_NORMAL).reshape(n, n, n, 4).astype(np.float64)
    got = o.readback(A.SLOT_RADIANCE).reshape(n, n, n, 4).astype(np.int64)
    o.close()
    cm = lambda c16: np.array(list(c16), np.float64).reshape(4, 4).T
    v2w = np.linalg.inv(cm(k.ext.VoxelProj) @ cm(k.ext.VoxelView))
    sview, sproj = cm(k.ext.ShadowView), cm(k.ext.ShadowProj)
    zz, yy, xx = np.nonzero(alb[..., 3])
    assert len(zz) > 5000 and not got[alb[..., 3] == 0].any()
    ndc = np.stack([(xx + 0.5) / n * 2 - 1, (yy + 0.5) / n * 2 - 1, (zz + 0.5) / n, np.ones(len(zz))])
    p = v2w @ ndc
    p = p[:3] / p[3]
    nv = nrm[zz, yy, xx, :3].T / 127.0
    ln = np.linalg.norm(nv, axis=0)
    nv = np.where(ln > 0, nv / np.maximum(ln, 1e-30), nv)
    col = (alb[zz, yy, xx, :3].T.astype(np.float64) / 255.0) ** 2.2
    s = sproj @ (sview @ np.vstack([p + nv * 0.06, np.ones(len(zz))]))
    s = s[:3] / s[3]
    tx, ty = np.floor((s[0] * 0.5 + 0.5) * sh).astype(int), np.floor((s[1] * 0.5 + 0.5) * sh).astype(int)
    inside = (tx >= 0) & (ty >= 0) & (tx < sh) & (ty < sh)
    shz = np.where(inside, fi["shadow"][np.clip(ty, 0, sh - 1), np.clip(tx, 0, sh - 1)], 0.0)
    margin = shz - (s[2] + 0.005)
    shade = (margin >= 0).astype(np.float64)
    sun_dir = np.array(list(k.sun.position), np.float64)
    lam = np.abs(-(sun_dir[:, None] * nv).sum(0))
    lum = np.array(list(k.sun.luminance), np.float64)
    want = np.floor(np.clip(col * lum[:, None] * lam * shade / lum.max(), 0, 1) * 255 + 0.5).astype(np.int64).T
    have = got[zz, yy, xx, :3]
    assert (got[zz, yy, xx, 3] == 255).all()
    on_edge = np.abs(margin) < 1e-5                        # the shadow comparison itself sits within fp32 noise
    diff = np.abs(have - want)
    assert (diff[~on_edge] <= 1).all(), int((diff[~on_edge] > 1).any(axis=1).sum())
    assert (diff[~on_edge] == 1).mean() < 0.02 and on_edge.mean() < 0.01
    assert have.max() > 20 and (shade == 0).any() and (shade == 1).any()       # lit and shadowed voxels both occur
