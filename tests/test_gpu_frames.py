"""-m gpu: whole-frame parity of the CUDA path (through the C ABI) against the CPU oracle.

Two modes are checked (include/risltc_cuda.h, risltc_cuda_set_precision):
  exact -- every operation rounded as in the oracle; the only differences left are libm ulps (atanf, acosf, sinf, cosf);
  fast  -- the production mode: the persistent fused kernel whose 32-candidate loop uses fused multiply-adds and MUFU
           rsqrt / rcp; the winner's estimator stays exactly rounded.
Tolerance (BASELINE.json north_star): relative RMSE <= 1e-3 on the converged (accumulated) frame and >= 99 % of the
pixels agreeing AT MATCHED SPP (1, 4 and 16 accumulated frames alike), where a pixel agrees iff
|a - b| <= 1e-3 max(|a|, |b|) + 1e-5 on every channel of the fp32 accumulation buffer. Both modes are
gated at that tolerance or tighter (GATES below); every figure is also appended to gpurun_out/parity_log.txt, which is
committed under profiles/ after a GPU session."""
from pathlib import Path

import numpy as np
import pytest

from tests.util import constants_bytes, image_metrics, parity_log, setup_device

pytestmark = pytest.mark.gpu

# minimal per-pixel agreement and maximal relative RMSE of a CONVERGED frame (>= 16 spp) per mode. A frame of 1-4 spp is not
# converged: one reservoir decision flipped by a rounding replaces a pixel's whole sample, which moves the frame's RMSE by
# more than 1e-3 while touching a single pixel; such frames are gated at RMSE_UNCONVERGED (the agreement gate is the same).
GATES = {"exact": dict(agree=0.999, rmse=1e-3), "fast": dict(agree=0.99, rmse=1e-3)}
RMSE_UNCONVERGED = 5e-2


def check(tag, precision, got, ref, frames, counters=None, ref_rays=None, agree_floor=None):
    rmse, agree = image_metrics(got, ref)
    extra = "" if counters is None else f" rays gpu={counters['shadow_rays']} cpu={ref_rays}"
    parity_log(f"{tag} [{precision}, {frames} spp]: rel_rmse={rmse:.3e} agree={agree:.5f}{extra}")
    gate = GATES[precision]
    assert agree >= (gate["agree"] if agree_floor is None else agree_floor), f"{tag}: pixel agreement {agree:.5f}"
    assert rmse <= (gate["rmse"] if frames >= 16 else RMSE_UNCONVERGED), f"{tag}: relative RMSE {rmse:.3e}"
    return rmse, agree


def _render_both(device, scene, ltc_tables, ovar, gvar, width, height, frames, precision="exact", **ckw):
    from oracle import orc
    _, rgba, rg = ltc_tables
    osc = orc.OracleScene(scene, rgba, rg)
    cs = [orc.make_constants(scene, width, height, orc.frame_words(f)[0], ltc_res=rgba.shape[1], ltc_layers=rgba.shape[0], **ckw) for f in range(frames)]
    ref, ref_vis, ref_rays = osc.render(cs, ovar)
    setup_device(device, scene, rgba, rg, gvar, width, height, osc.records)
    device.set_precision(precision)
    device.render_frames(constants_bytes(cs))
    got = device.read_accum()
    vis = device.read_visibility()
    return ref, ref_vis, ref_rays, got, vis, device.counters()


@pytest.mark.parametrize("precision", ["exact", "fast"])
@pytest.mark.parametrize("light_sampling,technique", [("uniform", "projected_solid_angle"), ("reservoir", "ltc_cp")])
def test_quad_over_plane(device, ltc_tables, light_sampling, technique, precision):
    """BASELINE.json configs[0] at reduced size: single quad light over a diffuse plane."""
    from oracle import orc
    from risltc_b200 import api, scenes
    scene = scenes.quad_over_plane(320, 180)
    kw = dict(light_sampling=light_sampling, technique=technique, min_vertices=4, max_vertices=4)
    ref, ref_vis, ref_rays, got, vis, counters = _render_both(device, scene, ltc_tables, orc.variant(**kw), api.variant(**kw), 320, 180, 1, precision)
    assert np.array_equal(vis, ref_vis), "primary visibility must be bit-exact"
    rmse, _ = check(f"quad over plane {light_sampling}/{technique}", precision, got, ref, 1, counters, ref_rays)
    assert rmse <= 1e-3


@pytest.mark.parametrize("precision", ["exact", "fast"])
@pytest.mark.parametrize("frames", [1, 4, 16])
def test_room_default_variant(device, ltc_tables, frames, precision):
    """The default estimator (RIS over LTC integrals -> PSA + LTC MIS) on a 64-light room; 16 frames = C2's spp."""
    from oracle import orc
    from risltc_b200 import api, scenes
    scene = scenes.many_light_room(64, 50, width=320, height=180)
    ref, ref_vis, ref_rays, got, vis, counters = _render_both(device, scene, ltc_tables, orc.variant(), api.variant(), 320, 180, frames, precision)
    assert np.array_equal(vis, ref_vis), "primary visibility must be bit-exact"
    check("room 64 lights 320x180, default estimator", precision, got, ref, frames, counters, ref_rays)
    assert counters["candidates"] == 32 * counters["shaded_pixels"]


VARIANTS = [
    dict(light_sampling="uniform"),
    dict(technique="projected_solid_angle"),
    dict(technique="area_turk"),
    dict(light_sampling="uniform", technique="area_turk"),
    dict(technique="projected_solid_angle", mis="balance", sample_count=2, light_samples=2),
    dict(mis="weighted"),
    dict(mis="optimal"),
    dict(mis="power", technique="projected_solid_angle_biased", fast_atan=1),
    dict(min_vertices=4, max_vertices=4),
    dict(min_vertices=5, max_vertices=5),
    dict(technique="projected_solid_angle", min_vertices=6, max_vertices=6),
    dict(light_sampling="uniform", technique="projected_solid_angle", min_vertices=3, max_vertices=7),
    dict(min_vertices=3, max_vertices=7),
]


@pytest.mark.parametrize("precision", ["exact", "fast"])
@pytest.mark.parametrize("kw", VARIANTS, ids=lambda k: "-".join(f"{a}={b}" for a, b in k.items()))
def test_room_variants(device, ltc_tables, kw, precision):
    """The other shader variants of the comparison matrix (experiment_list.c:316-396)."""
    from oracle import orc
    from risltc_b200 import api, scenes
    lo, hi = kw.get("min_vertices", 3), kw.get("max_vertices", 3)
    scene = scenes.many_light_room(24, 30, seed=4, width=256, height=144, vertex_count=hi if lo == hi else (lo, hi))
    ref, ref_vis, ref_rays, got, vis, counters = _render_both(device, scene, ltc_tables, orc.variant(**kw), api.variant(**kw), 256, 144, 2, precision)
    assert np.array_equal(vis, ref_vis)
    check("variant " + " ".join(f"{a}={b}" for a, b in kw.items()), precision, got, ref, 2, counters, ref_rays)



@pytest.mark.parametrize("frames", [1, 4, 16])
def test_quad_lights_on_the_specialised_kernels(device, ltc_tables, frames):
    """The default estimator on QUAD lights (MIN = MAX_POLYGONAL_LIGHT_VERTEX_COUNT = 4, the shape polygonal_light.c creates
    by default, reference variant ris_ltc_v4): in FAST precision ris_ltc4_kernel + winner_kernel<.., 4> instead of the
    generic kernel. Same gates as the triangle path: visibility bit-exact, >= 99 % of the pixels at 1 / 4 / 16 spp."""
    from oracle import orc
    from risltc_b200 import api, scenes
    kw = dict(min_vertices=4, max_vertices=4)
    scene = scenes.many_light_room(64, 60, seed=7, width=320, height=180, vertex_count=4)
    ref, ref_vis, ref_rays, got, vis, counters = _render_both(device, scene, ltc_tables, orc.variant(**kw), api.variant(**kw), 320, 180, frames, "fast")
    assert np.array_equal(vis, ref_vis)
    check("room 64 quad lights 320x180, default estimator", "fast", got, ref, frames, counters, ref_rays)
    assert abs(counters["shadow_rays"] - ref_rays) <= max(8, ref_rays // 10000)
    assert counters["candidates"] == 32 * counters["shaded_pixels"]      # the counters of the specialised RIS kernel


def test_mixed_triangle_and_quad_lights_on_the_specialised_kernels(device, ltc_tables):
    """MIN = 3, MAX = 4 vertices: write_lights repeats a triangle's first vertex in the fourth slot (main.c:483-487), the RIS
    kernel integrates the degenerate quad, the winner kernel clips with the light's own vertex count."""
    from oracle import orc
    from risltc_b200 import api, scenes
    kw = dict(min_vertices=3, max_vertices=4)
    scene = scenes.many_light_room(48, 60, seed=9, width=320, height=180, vertex_count=(3, 4))
    assert sorted({len(l["vertices_plane_space"]) for l in scene["lights"]}) == [3, 4]
    ref, ref_vis, ref_rays, got, vis, counters = _render_both(device, scene, ltc_tables, orc.variant(**kw), api.variant(**kw), 320, 180, 4, "fast")
    assert np.array_equal(vis, ref_vis)
    check("room 48 lights with 3 or 4 vertices 320x180, default estimator", "fast", got, ref, 4, counters, ref_rays)
    assert counters["candidates"] == 32 * counters["shaded_pixels"]      # the counters of the specialised RIS kernel


@pytest.mark.parametrize("precision", ["exact", "fast"])
@pytest.mark.parametrize("lights", [1024, 4096])
def test_many_lights(device, ltc_tables, lights, precision):
    """BASELINE.json configs[2] / [4] light counts against the oracle: 1024 lights is C3's table (the persistent RIS kernel
    then runs 22-warp CTAs beside a 48 KB table in shared memory), 4096 lights exceeds the shared-memory table and takes
    the kernel's global-memory instantiation (shade_fast.cuh, ris_ltc3_kernel<false>); shading_pass.frag.glsl:723-761."""
    from oracle import orc
    from risltc_b200 import api, scenes
    scene = scenes.many_light_room(lights, 60, seed=8, width=320, height=180)
    ref, ref_vis, ref_rays, got, vis, counters = _render_both(device, scene, ltc_tables, orc.variant(), api.variant(), 320, 180, 4, precision)
    assert np.array_equal(vis, ref_vis)
    check(f"room {lights} lights 320x180, default estimator", precision, got, ref, 4, counters, ref_rays)
    assert counters["candidates"] == 32 * counters["shaded_pixels"]


@pytest.mark.parametrize("precision", ["fast", "exact"])
def test_c2_full_size_against_compiled_reference(device, precision):
    """BASELINE.json configs[1] at its real size (1920x1080, 64 lights, the bench's scene and 64 x 64 x 51 LTC tables), two
    accumulated frames, against the REFERENCE's own shading_pass.frag.glsl compiled for the CPU (oracle/_ref) where that
    library travelled with the snapshot, else against the oracle (which is pinned bit for bit to it)."""
    from oracle import orc, ref
    from risltc_b200 import api, ltc_fit, scenes
    W, H, frames = 1920, 1080, 2
    rgba, rg = ltc_fit.quantize_fits(ltc_fit.fit_ggx_ltc(64, 51, 64))
    scene = scenes.many_light_room(64, 200, seed=2, width=W, height=H)
    osc = orc.OracleScene(scene, rgba, rg)
    cs = [orc.make_constants(scene, W, H, orc.frame_words(f)[0]) for f in range(frames)]
    if ref.available("ris_ltc_v3"):
        r = ref.RefShading("ris_ltc_v3"); r.bind(osc)
        want, want_vis, want_rays = r.render(cs)
        checker = "compiled reference GLSL"
    else:
        want, want_vis, want_rays = osc.render(cs, orc.variant())
        checker = "oracle"
    setup_device(device, scene, rgba, rg, api.variant(), W, H, osc.records)
    device.set_precision(precision)
    device.render_frames(constants_bytes(cs))
    got, vis, counters = device.read_accum(), device.read_visibility(), device.counters()
    assert np.array_equal(vis, want_vis), "primary visibility must be bit-exact at full size"
    check(f"C2 full size 1920x1080 vs {checker}", precision, got, want, frames, counters, want_rays)


@pytest.mark.parametrize("precision", ["fast", "exact"])
def test_c4_shaped_scene(device, ltc_tables, precision):
    """BASELINE.json configs[3] in shape: 64 lights over a field of 1.1 M occluder triangles. Beyond the rasteriser's queue of
    2^20 triangles the visibility pass is the per-pixel BVH walk (api.cu), the trees are ~25 levels deep and most shadow
    rays are occluded. Visibility bit-exact, image within tolerance, ray counts close."""
    from oracle import orc
    from risltc_b200 import api, scenes
    W, H, frames = 480, 270, 2
    scene = scenes.many_light_room(64, 200, seed=2, occluder_triangles=1_100_000, width=W, height=H)
    assert scene["mesh"]["material_indices"].shape[0] > (1 << 20)
    ref, ref_vis, ref_rays, got, vis, counters = _render_both(device, scene, ltc_tables, orc.variant(), api.variant(), W, H, frames, precision)
    assert np.array_equal(vis, ref_vis), "primary visibility must be bit-exact"
    check("C4-shaped scene, 1.1 M triangles, 480x270", precision, got, ref, frames, counters, ref_rays)
    assert abs(counters["shadow_rays"] - ref_rays) <= max(16, ref_rays // 2000)


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_stripes_equal_whole_frame(device, ltc_tables, precision):
    """any image partition gives the single-device image bit for bit."""
    from oracle import orc
    from risltc_b200 import api, scenes
    _, rgba, rg = ltc_tables
    W, H = 200, 123     # neither a multiple of the tile nor of the stripe height
    scene = scenes.many_light_room(32, 20, seed=6, width=W, height=H)
    osc = orc.OracleScene(scene, rgba, rg)
    cs = constants_bytes([orc.make_constants(scene, W, H, orc.frame_words(f)[0], ltc_res=rgba.shape[1], ltc_layers=rgba.shape[0]) for f in range(2)])
    setup_device(device, scene, rgba, rg, api.variant(), W, H, osc.records)
    device.set_precision(precision)
    device.render_frames(cs)
    whole = device.read_accum()
    assembled = np.zeros_like(whole)
    for index in range(3):
        device.resize(W, H, 8, index, 3)
        device.render_frames(cs)
        rows = device.owned_row_indices()
        assembled[rows] = device.read_accum()
    device.resize(W, H, 8, 0, 1)
    assert np.array_equal(assembled.view(np.uint32), whole.view(np.uint32))


def test_kernel_alternatives_are_bit_identical(device, ltc_tables):
    """The rasteriser and the BVH walk (visibility), the 4-wide and the binary tree (shadow rays) are interchangeable:
    visibility buffer and accumulated image must not differ in a single bit, whole frame or one stripe of three."""
    from oracle import orc
    from risltc_b200 import api, scenes
    _, rgba, rg = ltc_tables
    W, H = 333, 190
    scene = scenes.many_light_room(48, 120, seed=11, occluder_triangles=6000, width=W, height=H)
    osc = orc.OracleScene(scene, rgba, rg)
    cs = constants_bytes([orc.make_constants(scene, W, H, orc.frame_words(f)[0], ltc_res=rgba.shape[1], ltc_layers=rgba.shape[0]) for f in range(3)])
    _, ref_vis, _ = osc.render([orc.make_constants(scene, W, H, orc.frame_words(0)[0], ltc_res=rgba.shape[1], ltc_layers=rgba.shape[0])], orc.variant())
    try:
        for stripes in ((8, 0, 1), (8, 1, 3)):
            results = []
            for gbuffer, shadow in (("raster", "wide"), ("bvh", "wide"), ("raster", "binary"), ("auto", "wide"), ("raster", "pairs"), ("auto", "pairs")):
                setup_device(device, scene, rgba, rg, api.variant(), W, H, osc.records, stripes)
                device.set_kernels(gbuffer, shadow)
                device.render_frames(cs)
                results.append((device.read_visibility().copy(), device.read_accum().copy()))
            for vis, img in results[1:]:
                assert np.array_equal(vis, results[0][0])
                assert np.array_equal(img.view(np.uint32), results[0][1].view(np.uint32))
            if stripes[2] == 1:
                assert np.array_equal(results[0][0], ref_vis)
            assert 0.2 < np.mean(results[0][0] != 0xFFFFFFFF) <= 1.0
        # frame overlap (two streams, two sets of per-frame buffers) must not change a bit either, with 7 frames in flight
        cs7 = constants_bytes([orc.make_constants(scene, W, H, orc.frame_words(f)[0], ltc_res=rgba.shape[1], ltc_layers=rgba.shape[0]) for f in range(7)])
        images = []
        for mode in ("off", "on"):
            setup_device(device, scene, rgba, rg, api.variant(), W, H, osc.records, (8, 1, 3))
            device.set_frame_overlap(mode)
            device.render_frames(cs7)
            images.append((device.read_visibility().copy(), device.read_accum().copy()))
        assert np.array_equal(images[0][0], images[1][0])
        assert np.array_equal(images[0][1].view(np.uint32), images[1][1].view(np.uint32))
    finally:
        device.set_kernels("auto", "pairs")
        device.set_frame_overlap("auto")
        device.resize(W, H, 8, 0, 1)


def test_device_built_acceleration_structure(device, ltc_tables):
    """The acceleration structures built by kernels (bvh_gpu.cu; the reference builds on the device, scene.c:142-406):
    invariants of both trees, every triangle record bit-identical to the host builder's, any-hit decisions equal to the
    oracle's bit for bit, and frames (visibility + accumulation) bit-identical to those rendered with the host-built tree."""
    from oracle import orc
    from risltc_b200 import api, scenes
    _, rgba, rg = ltc_tables
    W, H = 320, 180
    scene = scenes.many_light_room(48, 120, seed=5, occluder_triangles=30000, width=W, height=H)
    osc = orc.OracleScene(scene, rgba, rg)
    cs = constants_bytes([orc.make_constants(scene, W, H, orc.frame_words(f)[0], ltc_res=rgba.shape[1], ltc_layers=rgba.shape[0]) for f in range(2)])
    rng = np.random.default_rng(9)
    n = 6000
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = rng.uniform([-9, -9, 0.1], [9, 9, 5.5], (n, 3))
    rays[n // 2:, 0:3] = rays[:n // 2, 0:3]
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d; rays[:, 3] = 1e-3; rays[:, 7] = rng.uniform(0.5, 25.0, n)
    want = np.array([osc.any_hit(r[0:3], r[4:7], float(r[3]), float(r[7])) for r in rays], dtype=np.uint32)
    results = {}
    try:
        for builder in ("host", "device", "radix"):
            device.set_bvh_builder(builder)
            setup_device(device, scene, rgba, rg, api.variant(), W, H, osc.records)
            stats, report = device.bvh_stats(), device.check_scene_bvh()
            assert stats["builder"] == builder
            assert report["bad_order"] == 0 and report["bad_binary"] == 0 and report["bad_wide"] == 0, report
            parity_log(f"acceleration structure, {builder} builder, {scene['mesh']['material_indices'].shape[0]} triangles: "
                       f"build {stats['build_ms']:.1f} ms, {report['binary_nodes']} binary / {report['wide_nodes']} 4-wide nodes, depth {stats['binary_depth']} / {stats['wide_depth']}, "
                       f"{report['children_per_node']:.2f} children per 4-wide node")
            for kind in (8, 4, 2):
                assert np.array_equal(device.kat_trace(rays, kind), want), f"{builder} builder, trace kernel {kind}"
            assert np.array_equal(device.kat_any_hit(rays), want)
            for gbuffer in ("raster", "bvh"):
                device.set_kernels(gbuffer, "pairs")
                device.render_frames(cs)
                results[(builder, gbuffer)] = (device.read_visibility().copy(), device.read_accum().copy())
        first = results[("host", "raster")]
        for key, (vis, img) in results.items():
            assert np.array_equal(vis, first[0]), key
            assert np.array_equal(img.view(np.uint32), first[1].view(np.uint32)), key
        # degenerate input for a Morton-order build: thousands of triangles with the same centroid key
        mesh = scenes.degenerate_soup(20000, seed=3) if hasattr(scenes, "degenerate_soup") else None
        if mesh is not None:
            device.upload_mesh(mesh)
            report = device.check_scene_bvh()
            assert report["bad_order"] == 0 and report["bad_binary"] == 0 and report["bad_wide"] == 0, report
    finally:
        device.set_bvh_builder("auto")
        device.set_kernels("auto", "pairs")


def test_full_size_properties(device, ltc_tables):
    """BASELINE.json configs[1] at full size (1920x1080, 64 lights), checked through size-independent properties:
    determinism, accumulation = running mean of single frames, background / emitter pixels, fast vs exact agreement."""
    from oracle import orc
    from risltc_b200 import api, scenes
    _, rgba, rg = ltc_tables
    W, H = 1920, 1080
    scene = scenes.many_light_room(64, 200, seed=2, width=W, height=H)
    records = orc.light_records(scene["lights"])
    cs = [orc.make_constants(scene, W, H, orc.frame_words(f)[0], ltc_res=rgba.shape[1], ltc_layers=rgba.shape[0]) for f in range(3)]
    setup_device(device, scene, rgba, rg, api.variant(), W, H, records)
    images = {}
    for precision in ("exact", "fast"):
        device.set_precision(precision)
        device.render_frames(constants_bytes(cs))
        a = device.read_accum()
        device.render_frames(constants_bytes(cs))
        assert np.array_equal(a.view(np.uint32), device.read_accum().view(np.uint32)), "not deterministic"
        singles = []
        for c in cs:
            device.render_frames(constants_bytes([c]))
            singles.append(device.read_accum())
        mean = singles[0].copy()
        for k in (1, 2):   # accum_pass.frag.glsl:45-53
            mean = (mean * np.float32(k) + singles[k]) * (np.float32(1.0) / np.float32(k + 1))
        assert np.allclose(a, mean, rtol=2e-6, atol=1e-7)
        vis = device.read_visibility()
        emitter = (vis >> 31 != 0) & (vis != 0xFFFFFFFF)
        assert np.all(a[emitter][:, :3] == np.float32(1.5)) and np.all(a[..., 3] == 1.0)
        assert np.all(np.isfinite(a))
        images[precision] = a
    rmse, agree = image_metrics(images["fast"], images["exact"])
    parity_log(f"1920x1080 fast vs exact, 3 frames: rel_rmse={rmse:.3e} agree={agree:.5f}")
    assert agree >= 0.99


def test_host_layer_path_equals_direct_path(ltc_tables, tmp_path):
    """The reference-shaped path (scene files -> load_scene / quick_load / load_ltc_table -> write_lights /
    write_constants -> C ABI) renders the same image as uploading numpy arrays directly, and as the oracle."""
    from oracle import orc
    from risltc_b200 import api, host, scenes
    fits, rgba, rg = ltc_tables
    W, H = 160, 90
    scene = scenes.many_light_room(16, 20, seed=12, width=W, height=H)
    # load_ltc_table(dir, 51) wants 51 Fresnel layers: repeat the small table's layers
    idx = np.minimum(np.arange(51) * rgba.shape[0] // 51, rgba.shape[0] - 1)
    fits51 = np.asarray(fits)[idx]
    vks, tex, save = host.write_scene_files(scene, tmp_path, ltc_fits=fits51)
    app = host.Application(tmp_path)
    try:
        app.load(vks, tex, save, W, H)
        app.settings(accum=1)
        app.reset(0)
        blocks = []
        consts_before = app.write_constants()     # advances the noise seed like a rendered frame would
        app.reset(0)
        app.render_frames(2)
        got = app.device().read_accum()
        lights = np.frombuffer(app.write_lights(), dtype=np.float32).reshape(16, -1)
    finally:
        app.close()
    from risltc_b200 import ltc_fit
    rgba51, rg51 = ltc_fit.quantize_fits(fits51)
    osc = orc.OracleScene(scene, rgba51, rg51)
    assert np.array_equal(lights.view(np.uint32), osc.records.view(np.uint32)), "write_lights differs from the oracle's record stream"
    cs = [orc.make_constants(scene, W, H, orc.frame_words(f)[0], ltc_res=rgba51.shape[1], ltc_layers=51) for f in range(2)]
    want = bytearray(C_string(cs[0]))
    words = orc.frame_words(0)   # set_noise_constants fills all four words; the shader reads only [0] (noise_utility.glsl:82)
    for i in range(4):
        want[208 + 4 * i:212 + 4 * i] = int(words[i]).to_bytes(4, "little")
    diff = [i for i in range(256) if consts_before[i] != want[i]]
    assert not diff, f"write_constants differs from the oracle's block at bytes {diff}"
    ref, _, _ = osc.render(cs, orc.variant())
    check("host layer path 160x90", "fast", got, ref, 2)


def C_string(c):
    import ctypes
    return ctypes.string_at(ctypes.byref(c), 256)


def test_output_stage(ltc_tables, tmp_path):
    """The copy pass (copy_pass.frag.glsl:28-58, srgb_utility.glsl:20-34) and implement_screenshot (main.c:2339-2409): the
    8-bit frames of the device equal the oracle's on the same accumulation buffer, the *.png decodes to the displayed
    frame and the *.hdr to the fp16 image assembled from the two half-bit frames."""
    from oracle import orc
    from risltc_b200 import formats, host, scenes
    fits, rgba, rg = ltc_tables
    W, H = 200, 120
    scene = scenes.many_light_room(16, 20, seed=12, width=W, height=H)
    idx = np.minimum(np.arange(51) * rgba.shape[0] // 51, rgba.shape[0] - 1)
    vks, tex, save = host.write_scene_files(scene, tmp_path, ltc_fits=np.asarray(fits)[idx])
    app = host.Application(tmp_path)
    try:
        app.load(vks, tex, save, W, H)
        app.settings(accum=1)
        app.reset(0)
        app.render_frames(3)
        dev = app.device()
        accum = dev.read_accum()
        frames = [dev.copy_pass(k) for k in range(3)]
        app.screenshot(png=tmp_path / "shot.png", hdr=tmp_path / "shot.hdr")
    finally:
        app.close()
    for k in (1, 2):
        assert np.array_equal(frames[k], orc.copy_pass(accum, k)), f"half-bit frame {k}"
    display = orc.copy_pass(accum, 0)
    off = np.abs(frames[0].astype(int) - display.astype(int))
    parity_log(f"copy pass {W}x{H}: displayed frame differs from the oracle in {np.count_nonzero(off)} of {off.size} bytes (max {off.max()})")
    assert off.max() <= 1 and np.count_nonzero(off) <= 3      # pow() is correctly rounded on both sides
    assert np.array_equal(formats.read_png(tmp_path / "shot.png"), frames[0])
    want = accum[..., :3].astype(np.float16).astype(np.float32)
    got = formats.read_hdr(tmp_path / "shot.hdr")
    assert np.all(np.abs(got - want) <= want.max(axis=-1, keepdims=True) * 2.0 ** -7 + 1e-30)
    assert (frames[0].max() == 255) and (frames[0].min() == 0)


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_textured_materials(device, ltc_tables, precision):
    """Mip-mapped material textures (get_shading_data's derivative block and three textureGrad fetches,
    shading_pass.frag.glsl:604-635) against the oracle under the stated filtering definition: exact mode bit for bit."""
    from oracle import orc
    from risltc_b200 import api, scenes
    W, H = 320, 180
    scene = scenes.add_procedural_textures(scenes.many_light_room(32, 40, seed=3, width=W, height=H))
    ref, ref_vis, ref_rays, got, vis, counters = _render_both(device, scene, ltc_tables, orc.variant(), api.variant(), W, H, 4, precision)
    assert np.array_equal(vis, ref_vis)
    check("textured materials, 32 lights 320x180", precision, got, ref, 4, counters, ref_rays)
    flat = scenes.many_light_room(32, 40, seed=3, width=W, height=H)
    flat_image = _render_both(device, flat, ltc_tables, orc.variant(), api.variant(), W, H, 4, precision)[3]
    assert np.mean(np.any(got != flat_image, axis=-1)) > 0.5, "the textures have no effect"


def test_textured_scene_through_the_host_layer(ltc_tables, tmp_path):
    """Block-compressed *.vkt files (BC1 sRGB base colour, BC1 specular, BC5 normal, full mip chains; textures.c:95-241) through
    load_scene render the same image as uploading the decoded texels directly, and as the oracle on those texels."""
    from oracle import orc
    from risltc_b200 import api, formats, host, scenes
    fits, rgba, rg = ltc_tables
    W, H = 200, 120
    scene = scenes.add_procedural_textures(scenes.many_light_room(16, 20, seed=12, width=W, height=H), size=32)
    idx = np.minimum(np.arange(51) * rgba.shape[0] // 51, rgba.shape[0] - 1)
    fits51 = np.asarray(fits)[idx]
    vks, tex, save = host.write_scene_files(scene, tmp_path, ltc_fits=fits51)
    app = host.Application(tmp_path)
    try:
        app.load(vks, tex, save, W, H)
        app.settings(accum=1)
        app.reset(0)
        app.render_frames(2)
        got = app.device().read_accum()
    finally:
        app.close()
    # what the files hold after lossy block compression, decoded by the independent numpy decoder
    decoded = dict(scene)
    decoded["textures"] = []
    for i, m in enumerate(scene["materials"]):
        for j, suffix in enumerate(("BaseColor", "Specular", "Normal")):
            levels, fmt = formats.read_vkt(Path(tex) / f"{m['name']}_{suffix}.vkt", with_format=True)
            decoded["textures"].append(dict(format="rgba8_srgb" if fmt == formats.VK_FORMAT_BC1_RGB_SRGB_BLOCK else "rgba8_unorm", levels=levels))
    from risltc_b200 import ltc_fit
    rgba51, rg51 = ltc_fit.quantize_fits(fits51)
    osc = orc.OracleScene(decoded, rgba51, rg51)
    cs = [orc.make_constants(scene, W, H, orc.frame_words(f)[0], ltc_res=rgba51.shape[1], ltc_layers=51) for f in range(2)]
    want, _, _ = osc.render(cs, orc.variant())
    check("textured scene through load_scene (BC1 / BC5 files)", "fast", got, want, 2)
    dev = api.Device(0)
    try:
        setup_device(dev, decoded, rgba51, rg51, api.variant(), W, H, osc.records)
        dev.render_frames(constants_bytes(cs))
        direct = dev.read_accum()
    finally:
        dev.close()
    assert np.array_equal(direct.view(np.uint32), got.view(np.uint32)), "host loader and direct upload of the decoded texels differ"
