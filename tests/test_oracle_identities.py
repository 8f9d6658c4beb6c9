"""Closed-form identities of the shading arithmetic (SURVEY.md 8c list, items 1, 2, 4): independent of any reference output."""
import numpy as np

from oracle import orc


def _poly(points):
    v = np.zeros((8, 3), dtype=np.float32)
    v[: len(points)] = np.asarray(points, dtype=np.float32)
    return v


def test_wang_frame_words():
    # g_noise_random_numbers.x = wang_random_number(4 * frame), noise_table.c:26; values computed from math_utilities.h:50-57
    assert [orc.wang(4 * f) for f in range(3)] == [0xc0a9496a, 0xcc49325c, 0xb49ae7f7]


def test_noise_is_lcg_of_murmur_seed():
    seed = int(orc.lib().orc_noise_seed(3, 5, 640, 0xc0a9496a))
    draws = orc.noise(3, 5, 640, 0xc0a9496a, 4)
    s = seed
    for k in range(4):
        s = (1664525 * s + 1013904223) & 0xFFFFFFFF
        assert draws[k] == np.float32(np.float32(s) * np.float32(2.0 ** -32))


def test_hemisphere_projected_solid_angle_is_pi():
    # a huge square just above the horizon plane covers the hemisphere: PSA -> pi, LTC form factor -> 1
    z = 1e-4
    v = _poly([[-1e3, -1e3, z], [1e3, -1e3, z], [1e3, 1e3, z], [-1e3, 1e3, z]])
    poly, _ = orc.psa(4, v, 0.3, 0.7, 5)
    assert abs(poly[43] - np.pi) < 2e-3
    assert abs(orc.calculate_ltc(4, v) - 1.0) < 2e-3


def test_ltc_integral_matches_psa_by_two_code_paths():
    """pi * calculate_ltc(P) (polygon_sampling.glsl:523-530) and prepare_...(P).projected_solid_angle (:545-613) are the
    same quantity; the rational fit of integrateEdgeVec limits the agreement to ~1e-3 relative."""
    rng = np.random.default_rng(3)
    worst = 0.0
    for _ in range(200):
        c = np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(0.5, 3.0)])
        ang = np.sort(rng.uniform(0, 2 * np.pi, 3))
        tri = c + 0.4 * np.stack([np.cos(ang), np.sin(ang), 0.3 * rng.normal(size=3)], axis=1)
        # the sampler expects the winding that the shading pass establishes by flipping the frame (shading_pass.frag.glsl:300-304)
        for cand in (tri, tri[::-1]):
            v = _poly(cand)
            vc, clipped = orc.clip_polygon(3, v, 3, 4)
            assert vc == 3
            poly, _ = orc.psa(vc, clipped, 0.5, 0.5, 4)
            if poly[43] > 0:
                break
        total, ff = float(poly[43]), orc.calculate_ltc(vc, clipped)
        assert total > 0
        assert abs(sum(poly[35:35 + (3 if poly[33] > 0 else 2)]) - total) <= 1e-6 * max(total, 1e-6) + 1e-7   # sectors sum to the total
        worst = max(worst, abs(np.pi * ff - total) / total)
    assert worst < 5e-3


def test_psa_samples_lie_inside_the_polygon():
    rng = np.random.default_rng(5)
    v = _poly([[0.3, -0.2, 1.0], [0.9, 0.1, 0.8], [0.5, 0.7, 1.2]])
    vc, clipped = orc.clip_polygon(3, v, 3, 4)
    n = np.cross(clipped[1] - clipped[0], clipped[2] - clipped[0])
    for _ in range(300):
        _, d = orc.psa(vc, clipped, float(rng.random()), float(rng.random()), 4)
        assert abs(np.linalg.norm(d) - 1.0) < 1e-5 and d[2] > 0
        t = (clipped[0] @ n) / (d @ n)
        p = t * d
        for i in range(3):
            a, b = clipped[i], clipped[(i + 1) % 3]
            assert np.cross(b - a, p - a) @ n > -2e-3 * (n @ n), "sample outside the light"


def _sutherland_hodgman(points):
    out = []
    n = len(points)
    for i in range(n):
        a, b = points[i], points[(i + 1) % n]
        ina, inb = a[2] > 0, b[2] > 0
        if ina:
            out.append(a)
        if ina != inb:
            t = a[2] / (a[2] - b[2])
            out.append(a + t * (b - a))
    return out


def test_clip_table_is_exhaustive_against_sutherland_hodgman():
    """Every above-horizon mask of every vertex count 3..7: same vertex set (as a cycle) as a generic clipper."""
    for n in range(3, 8):
        ang = np.arange(n) * (2 * np.pi / n)
        ring = np.stack([np.cos(ang), np.sin(ang)], axis=1)
        for mask in range(1 << n):
            z = np.where([(mask >> i) & 1 for i in range(n)], 1.0, -1.0)
            # only masks with one contiguous run of vertices above the horizon arise from convex polygons
            runs = sum(1 for i in range(n) if z[i] > 0 and z[(i - 1) % n] <= 0)
            if runs > 1:
                continue
            pts = np.concatenate([ring + np.array([0.1, 0.05]), (z * (0.5 + 0.1 * np.arange(n)))[:, None]], axis=1)
            vc, clipped = orc.clip_polygon(n, _poly(pts), n, n + 1)
            want = _sutherland_hodgman([p for p in pts])
            assert vc == len(want), (n, mask)
            if vc == 0:
                continue
            got = clipped[:vc].astype(np.float64)
            want = np.array(want)
            start = int(np.argmin(np.linalg.norm(want - got[0], axis=1)))
            assert np.allclose(np.roll(want, -start, axis=0), got, atol=1e-5), (n, mask)
            if vc < n + 1:
                assert np.array_equal(clipped[vc], clipped[0])   # first vertex repeated at [vc]


def test_accumulation_is_a_running_mean():
    rng = np.random.default_rng(1)
    import ctypes as C
    frames = rng.random((5, 16, 4)).astype(np.float32)
    acc = np.zeros((16, 4), dtype=np.float32)
    for k in range(5):
        orc.lib().orc_accum_pass(acc.ctypes.data_as(C.c_void_p), frames[k].ctypes.data_as(C.c_void_p), C.c_uint32(k), C.c_uint64(16))
    assert np.allclose(acc, frames.mean(axis=0), rtol=1e-6)


def test_unorm_decodes_are_exactly_rounded_quotients():
    """The kernels decode x / 255 (texels) and x / 65535 (vertex attributes, LTC tables) as q0 = x * c, q0 + (x - q0 * d) * c
    with c = fl(1 / d) and fused multiply-adds (risltc_b200/csrc/common.cuh: unorm8, unorm16) instead of an IEEE division.
    Checked here against exact rational arithmetic for every input: the sequence returns the correctly rounded quotient."""
    from fractions import Fraction as F
    import math

    def fl32(fr):
        if fr == 0:
            return F(0)
        sign, a = (1 if fr > 0 else -1), abs(fr)
        e = math.floor(math.log2(float(a)))
        while F(2) ** e > a:
            e -= 1
        while F(2) ** (e + 1) <= a:
            e += 1
        ulp = F(2) ** (e - 23)
        q = a / ulp
        n = q.numerator // q.denominator
        rem = q - n
        if rem > F(1, 2) or (rem == F(1, 2) and n % 2 == 1):
            n += 1
        return sign * n * ulp

    for d in (255, 65535):
        c = fl32(F(1, d))
        assert c == F(float(np.float32(1.0) / np.float32(d)))
        for x in range(d + 1):
            q0 = fl32(F(x) * c)
            r = fl32(fl32(-q0 * d + x) * c + q0)      # two fused multiply-adds: each rounds once
            assert r == fl32(F(x, d)), (x, d)
