"""-m gpu: the device functions of the shading kernels, one at a time, against the oracle (and therefore against the golden
vectors of the compiled reference): clip, LTC integral, PSA prepare + sample, noise stream, LTC coefficients, any-hit.
Tolerance: bit-exact for integer / index work and for everything that does not call a transcendental function; where
atan / acos / sin / cos are involved <= 1e-5 relative (BASELINE.json: "LTC-table and polygon-sample setup ... <= 1e-5
relative") -- and the share of bit-identical values is logged and gated too, because kat.cu computes those functions
correctly rounded like the oracle defines them (csrc/common.cuh RL_CR_LIBM, oracle/risltc_oracle.c glsl_atan)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from tests.util import parity_log

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def fn():
    return np.load(GOLDEN / "functions.npz")


@pytest.mark.parametrize("name,V", [("ris_ltc_v3", 3), ("ris_ltc_v4", 4), ("uni_psa_biased_fast_v5", 5), ("uni_psa_v7", 7)])
def test_clip_and_ltc_integral_bit_exact(device, fn, name, V):
    polys, counts = fn[f"{name}.polygons"], fn[f"{name}.counts"]
    got_p, got_c = device.kat_clip(polys, counts, V)
    want_c = fn[f"{name}.clipped_counts"]
    assert np.array_equal(got_c, want_c)
    for i in range(len(counts)):
        n = int(want_c[i])   # the kernels close the polygon by index instead of storing vertex 0 again at slot [vc]
        assert np.array_equal(got_p[i, :n].view(np.uint32), fn[f"{name}.clipped"][i, :n].view(np.uint32)), i
    valid = ~np.isnan(fn[f"{name}.ltc_integral"])
    if valid.any():
        got = device.kat_ltc_integral(fn[f"{name}.clipped"][valid], want_c[valid])
        assert np.array_equal(got.view(np.uint32), fn[f"{name}.ltc_integral"][valid].view(np.uint32))


@pytest.mark.parametrize("name,P,fast,biased", [("ris_ltc_v3", 4, 0, 0), ("ris_ltc_v4", 5, 0, 0), ("uni_psa_biased_fast_v5", 6, 1, 1), ("uni_psa_v7", 8, 0, 0)])
def test_psa_prepare_and_sample(device, fn, name, P, fast, biased):
    keep = fn[f"{name}.clipped_counts"] > 0
    polys, counts, rnd = fn[f"{name}.clipped"][keep], fn[f"{name}.clipped_counts"][keep], fn[f"{name}.randoms"][keep]
    got_poly, got_dir = device.kat_psa(polys, counts, rnd, P, fast, biased)
    want_poly, want_dir = fn[f"{name}.psa_polygon"][keep].copy(), fn[f"{name}.psa_dir"][keep]
    for i in range(len(counts)):   # the decentral case leaves sector [vc - 1] unwritten in the reference
        if not want_poly[i, 33] > 0:
            want_poly[i, 35 + counts[i] - 1] = 0.0; got_poly[i, 35 + counts[i] - 1] = 0.0
    # vertices, ellipses, inner ellipse: no libm -> bit-exact (kahan determinants, sign-of-zero tests, sort network)
    assert np.array_equal(got_poly[:, :35].view(np.uint32), want_poly[:, :35].view(np.uint32))
    # sector areas and total: atanf (or the polynomial fast atan, which is bit-exact)
    if fast:
        assert np.array_equal(got_poly[:, 35:].view(np.uint32), want_poly[:, 35:].view(np.uint32))
    else:
        total = want_poly[:, 43:44]
        got_a, want_a = got_poly[:, 35:], want_poly[:, 35:]
        both_nan = np.isnan(got_a) & np.isnan(want_a)     # degenerate polygons (0 / 0 in a tangent) are NaN in the reference, too
        close = (np.abs(got_a - want_a) <= 1e-5 * np.nan_to_num(total) + 1e-7) | both_nan
        if not close.all():
            bad = np.argwhere(~close)[:5]
            parity_log(f"kat psa {name}: {np.count_nonzero(~close)} sector areas out of tolerance, e.g. " + "; ".join(f"[{i},{j}] got {got_a[i, j]!r} want {want_a[i, j]!r}" for i, j in bad))
        assert close.all()
        same = np.mean((got_a.view(np.uint32) == want_a.view(np.uint32)) | both_nan)
        parity_log(f"kat psa {name}: sector areas bit-identical {same:.5f}")
        assert same >= 0.999
    # sampled directions: the iteration amplifies ulps where sectors are thin; 99 % within 1e-4, all unit length and above the horizon
    ok = np.isfinite(want_dir).all(axis=1)     # degenerate polygons (zero solid angle) sample NaN in the reference, too
    assert np.array_equal(ok, np.isfinite(got_dir).all(axis=1))
    err = np.linalg.norm(got_dir[ok] - want_dir[ok], axis=1)
    same = np.mean(np.all(got_dir[ok].view(np.uint32) == want_dir[ok].view(np.uint32), axis=1))
    parity_log(f"kat psa {name}: sampled directions bit-identical {same:.5f}, within 1e-4: {np.mean(err <= 1e-4):.5f}, max error {err.max():.3e}")
    assert np.mean(err <= 1e-4) >= 0.99, np.sort(err)[-5:]
    assert same >= 0.99
    assert np.all(np.abs(np.linalg.norm(got_dir[ok], axis=1) - 1.0) < 1e-5) and np.all(got_dir[ok][:, 2] >= 0.0)


def test_noise_stream_bit_exact(device, fn):
    for k, w in enumerate(fn["noise.frame_words"]):
        out = device.kat_noise(640, 360, int(w), 8)
        for j, (px, py) in enumerate(fn["noise.pixels"]):
            assert np.array_equal(out[py, px].view(np.uint32), fn["noise.draws"][k, j].view(np.uint32))


def test_ltc_coefficients(device, fn):
    device.upload_ltc(fn["ltc.rgba16"], fn["ltc.rg16"])
    got = device.kat_ltc_coefficients(fn["ltc.inputs"], [float(x) for x in fn["ltc.constants"]])[:, :32]
    want = fn["ltc.coefficients"]
    ok = np.isfinite(want).all(axis=1)         # outgoing == normal has no tangent frame (NaN) in the reference, too
    assert np.array_equal(ok, np.isfinite(got).all(axis=1)) and ok.sum() > 100
    scale = np.maximum(np.abs(want[ok]).max(axis=1, keepdims=True), 1e-6)
    assert np.all(np.abs(got[ok] - want[ok]) <= 1e-5 * scale), np.abs(got[ok] - want[ok]).max()   # acos feeds the bilinear lookup
    same = np.mean(got[ok].view(np.uint32) == want[ok].view(np.uint32))
    parity_log(f"kat ltc coefficients: bit-identical {same:.5f}")
    assert same >= 0.999


def test_any_hit_matches_oracle_bit_for_bit(device, ltc_tables):
    from oracle import orc
    from risltc_b200 import scenes
    _, rgba, rg = ltc_tables
    scene = scenes.many_light_room(16, 60, seed=8, occluder_triangles=20000, width=64, height=36)
    osc = orc.OracleScene(scene, rgba, rg)
    device.upload_mesh(scene["mesh"])
    rng = np.random.default_rng(4)
    n = 20000
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = rng.uniform([-9, -9, 0.1], [9, 9, 5.5], (n, 3))
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d; rays[:, 3] = 1e-3; rays[:, 7] = rng.uniform(0.5, 25.0, n)
    # rays through mesh vertices and along edges: the decisions that need the exactly rounded quotients
    verts = scenes.dequantize_positions(scene["mesh"]).astype(np.float32)
    for i in range(2000):
        t = verts[rng.integers(0, len(verts))]
        if i % 2:
            t = (t + verts[rng.integers(0, len(verts))]) * np.float32(0.5)
        v = t - rays[i, 0:3]
        rays[i, 4:7] = v / np.linalg.norm(v)
    got = device.kat_any_hit(rays)
    want = np.array([osc.any_hit(r[0:3], r[4:7], float(r[3]), float(r[7])) for r in rays], dtype=np.uint32)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} of {n} any-hit decisions differ"
    # the shadow-ray kernels of the frame path: the 4-wide quantised tree (default) and the binary tree
    for kind in (4, 2):
        got = device.kat_trace(rays, kind)
        assert np.array_equal(got, want), f"trace kernel {kind}: {np.count_nonzero(got != want)} of {n} any-hit decisions differ"
    assert 0.05 < want.mean() < 0.98
    # the default kernel traverses the tree once for the two rays of a pixel (ray i and ray i + n / 2 share their origin):
    # nearly parallel pairs (two samples of one small light), unrelated pairs (other octants: traced alone), pairs of which
    # one ray is switched off (t_max below t_min), pairs that repeat the vertex / edge rays above
    half = n // 2
    pairs = rays.copy()
    pairs[half:, 0:3] = pairs[:half, 0:3]
    jitter = pairs[:half, 4:7] + rng.normal(scale=0.03, size=(half, 3)).astype(np.float32)
    coherent = rng.random(half) < 0.6
    pairs[half:, 4:7][coherent] = (jitter / np.linalg.norm(jitter, axis=1, keepdims=True))[coherent]
    pairs[half:half + 1000, 4:7] = pairs[:1000, 4:7]                       # the exact vertex / edge rays, twice in a pair
    pairs[half:half + 1000, 7] = rng.uniform(0.5, 25.0, 1000)
    off = rng.random(n) < 0.1
    pairs[off, 7] = 0.0
    want2 = np.array([osc.any_hit(r[0:3], r[4:7], float(r[3]), float(r[7])) if r[7] > 1e-3 else 0 for r in pairs], dtype=np.uint32)
    for kind in (8, 4):
        got = device.kat_trace(pairs, kind)
        assert np.array_equal(got, want2), f"trace kernel {kind} on pairs: {np.count_nonzero(got != want2)} of {n} any-hit decisions differ"
    assert 0.05 < want2.mean() < 0.98


def test_exactly_rounded_sequences_exhaustively(device):
    """inversesqrt (RSQ seed + FMA corrections) == 1 / sqrt for every float, unorm16 == x / 65535 for every 16-bit value."""
    assert device.kat_exact_math() == [0, 0, 0]
