"""The view-driven gather's level-1 mask (mode_n_shard.cu: k_need_bricks) does not march the six diffuse cones: it marks the box of
bricks around the pixel's world position that must hold every fine sample they take.  A brick too few would be a silently wrong
image on a multi-GPU box, so the argument is checked numerically here: for random surface points, normals, tangent frames and a
sheared / anisotropic world->volume map, every brick the tracer's own march touches at level <= 1 (restated from mark_cone: same
loop, same footprint with its 1/16-texel margin) lies inside the box (restated from the kernel: reach = h + last fine t, per-axis
texels-per-world = row norms, slack 1/16 + 1/4 texel) — for the default tracer and for the Appendix-B one."""
import numpy as np
import pytest

TAN_DIFFUSE = 0.57735027
DIRS = np.array([[0.0, 0.0, 1.0], [0.8660254, 0.0, 0.5], [0.26761657, 0.82363910, 0.5], [-0.70062927, 0.50903696, 0.5],
                 [-0.70062927, -0.50903696, 0.5], [0.26761657, -0.82363910, 0.5]])


def fine_ts(h, tan_half, spec_b, max_dist):
    """the t of every sample a cone takes while its level is <= 1 (mark_cone / cone_fine_reach)"""
    out, t = [], 2.0 * h
    while t < max_dist:
        diam = max(h, 2.0 * t * tan_half)
        lod = np.log2(diam / h)
        if lod >= (2.0 if spec_b else 1.5) + 1e-3:
            break
        out.append(t)
        t += 0.5 * diam if spec_b else diam
    return out


def bricks_of_sample(q, n1):
    p = q * n1 - 0.5
    e = 1.0 / 16.0
    lo = np.floor(p - e).astype(int) >> 2
    hi = (np.floor(p + e).astype(int) + 1) >> 2
    return lo, hi


@pytest.mark.parametrize("spec_b", [False, True])
def test_box_of_bricks_covers_every_fine_sample_of_the_diffuse_cones(spec_b):
    rng = np.random.default_rng(7 + int(spec_b))
    N = 256
    n1 = N // 2
    for case in range(300):
        # world -> normalised volume coordinates: q = 0.5 * (A x + b) + 0.5 for x, y; A x + b for z — a rotated, anisotropic box
        R, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        scale = rng.uniform(0.02, 0.08, 3)
        A = (R * scale[None, :]).T                       # rows = volume axes
        b = rng.uniform(-0.2, 0.2, 3)
        half = np.array([0.5, 0.5, 1.0])
        to_q = lambda x: (A @ x + b) * half + np.array([0.5, 0.5, 0.0])
        h = 1.0 / (np.linalg.norm(A[0]) * 0.5 * N)       # world size of a voxel along volume-x, as f184_voxel_h
        texel_per_world = np.array([np.linalg.norm(A[i]) * half[i] * n1 for i in range(3)])
        reach = max(fine_ts(h, TAN_DIFFUSE, spec_b, 32.0), default=0.0)
        assert reach > 0.0
        wpos = np.linalg.solve(A, (rng.uniform(0.1, 0.9, 3) - np.array([0.5, 0.5, 0.0])) / half - b)
        z = rng.normal(size=3); z /= np.linalg.norm(z)
        hh = z.copy(); hh[np.argmin(np.abs(hh))] = 1.0
        ty = np.cross(hh, z); ty /= np.linalg.norm(ty)
        tx = np.cross(z, ty); tx /= np.linalg.norm(tx)
        origin = wpos + z * h
        # the box (kernel arithmetic, float32 where it rounds)
        pc = (to_q(wpos) * n1 - 0.5).astype(np.float32)
        r = np.float32((h + reach) * 1.0001)
        rad = (r * texel_per_world.astype(np.float32) + np.float32(1.0 / 16.0 + 0.25)).astype(np.float32)
        box_lo = np.floor(pc - rad).astype(int) >> 2
        box_hi = (np.floor(pc + rad).astype(int) + 1) >> 2
        for d in DIRS:
            direction = tx * d[0] + ty * d[1] + z * d[2]
            for t in fine_ts(h, TAN_DIFFUSE, spec_b, 32.0):
                lo, hi = bricks_of_sample(to_q(origin + direction * t), n1)
                assert (lo >= box_lo).all() and (hi <= box_hi).all(), (case, t, lo, hi, box_lo, box_hi)


def test_a_rough_specular_cone_needs_no_march_and_a_glossy_one_does():
    """cone_specular_tan = clamp(rough^2, 0.02, 0.6).  The kernel marches the specular cone only where its fine reach exceeds the diffuse
    one (otherwise its samples are inside the box already): never for rough materials under the default tracer; always for glossy
    ones, which stay fine for tens of voxels.  (Under Appendix B a tan of 0.6 reaches 3.20 h against the diffuse 3.155 h: marched —
    slower, never wrong.)"""
    h = 0.05
    clamp = lambda r: min(max(r * r, 0.02), 0.6)
    for spec_b in (False, True):
        reach_d = max(fine_ts(h, TAN_DIFFUSE, spec_b, 32.0))
        glossy_reach = max(fine_ts(h, clamp(0.2), spec_b, 32.0))
        assert reach_d < glossy_reach and glossy_reach > 10 * h
    assert max(fine_ts(h, clamp(1.0), False, 32.0)) <= max(fine_ts(h, TAN_DIFFUSE, False, 32.0))
