"""The C99 host layer (librisltc_host.so) without a GPU: reference structs, host arithmetic against the golden vectors of
the reference's own polygonal_light.c / camera.c / math_utilities.h, experiment list, file formats, .hdr writer."""
import ctypes as C
import os
import struct
from pathlib import Path

import numpy as np
import pytest

from oracle import ref as ref_mod
from risltc_b200 import formats, host, ltc_fit, scenes

GOLDEN = Path(__file__).resolve().parent / "golden"
PolygonalLight, Camera = ref_mod.RefHost.PolygonalLight, ref_mod.RefHost.Camera


@pytest.fixture(scope="module")
def lib():
    return host.lib()


@pytest.fixture(scope="module")
def fn():
    return np.load(GOLDEN / "functions.npz")


def test_struct_layouts_match_the_reference():
    # polygonal_light_t is 184 bytes on LP64 with these offsets; first_person_camera_t is 48 bytes
    assert C.sizeof(PolygonalLight) == 184 and C.sizeof(Camera) == 48
    for name, off in (("surface_radiance", 48), ("plane", 64), ("vertex_count", 80), ("texturing_technique", 84), ("texture_index", 88),
                      ("rotation", 96), ("area", 144), ("texture_file_path", 160), ("vertices_plane_space", 168), ("vertices_world_space", 176)):
        assert getattr(PolygonalLight, name).offset == off, name
    assert C.sizeof(host.RenderSettings) == 64


def test_update_polygonal_light_bit_exact(lib, fn):
    h = ref_mod.RefHost(path=host.PKG / "librisltc_host.so")
    li, pv = fn["host.light_inputs"], fn["host.light_plane_vertices"]
    for i in range(li.shape[0]):
        light = dict(rotation_angles=[float(x) for x in li[i, 0:3]], scaling_x=float(li[i, 3]), scaling_y=float(li[i, 4]),
                     translation=[float(x) for x in li[i, 5:8]], radiant_flux=[float(x) for x in li[i, 8:11]], vertices_plane_space=pv[i])
        world, plane, rad, area, _ = h.update_light(light)
        assert np.array_equal(world[:, :3].view(np.uint32), fn["host.light_world"][i][:, :3].view(np.uint32))
        assert np.array_equal(plane.view(np.uint32), fn["host.light_plane"][i].view(np.uint32))
        assert np.array_equal(rad.view(np.uint32), fn["host.light_radiance"][i].view(np.uint32))
        assert np.float32(area) == fn["host.light_area"][i]


def test_camera_and_math_bit_exact(lib, fn):
    h = ref_mod.RefHost(path=host.PKG / "librisltc_host.so")
    c = fn["host.camera"]
    cam = dict(position=[float(x) for x in c[0:3]], rotation_x=float(c[3]), rotation_z=float(c[4]), vertical_fov=float(c[5]), near=float(c[6]), far=float(c[7]))
    got = h.world_to_projection(cam, float(np.float32(16) / np.float32(9)))
    assert np.array_equal(got.view(np.uint32), fn["host.world_to_projection"].view(np.uint32))
    lib.wang_random_number.restype = C.c_uint32
    for s, w in zip(fn["host.wang_seeds"], fn["host.wang"]):
        assert int(lib.wang_random_number(C.c_uint32(int(s)))) == int(w)
    m = fn["host.matrix"]
    a = ((C.c_float * 4) * 4)(*[(C.c_float * 4)(*[float(x) for x in row]) for row in m])
    out = ((C.c_float * 4) * 4)()
    lib.matrix_inverse(out, a)
    assert np.array_equal(np.array([list(r) for r in out], dtype=np.float32).view(np.uint32), fn["host.matrix_inverse"].view(np.uint32))


def test_light_lifecycle_conventions(lib):
    light = PolygonalLight()
    assert lib.set_polygonal_light_vertex_count(C.byref(light), C.c_uint32(5)) == 0
    assert light.vertex_count == 5 and bool(light.vertices_plane_space) and bool(light.vertices_world_space)
    lib.destroy_polygonal_light(C.byref(light))
    assert bytes(light) == bytes(184)          # destroy memsets to zero (polygonal_light.c:113-118)
    lib.destroy_polygonal_light(C.byref(light))  # and tolerates a zeroed object


class Experiment(C.Structure):   # experiment_t, main.h:176-208
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("scene_index", C.c_int), ("quick_save_path", C.c_char_p), ("use_hdr", C.c_uint32),
                ("screenshot_path", C.c_char_p), ("num_samples", C.c_uint32), ("render_settings", host.RenderSettings), ("base_dir", C.c_char_p),
                ("timings_path", C.c_char_p), ("screenshots_dir", C.c_char_p), ("ext", C.c_char_p), ("exp_name", C.c_char_p), ("ss_per_frame", C.c_uint32)]


class ExperimentList(C.Structure):   # experiment_list_t, main.h:225-241
    _fields_ = [("experiments", C.POINTER(Experiment)), ("experiment", C.POINTER(Experiment)), ("count", C.c_uint32), ("next", C.c_uint32),
                ("next_setup_frame", C.c_uint32), ("state", C.c_int), ("timings_file", C.c_void_p)]


def test_experiment_list_from_environment(lib, monkeypatch):
    for k in list(os.environ):
        if k.startswith("EXP_") or k in ("NUM_SAMPLES", "SCENE", "COMPUTE_GT"):
            monkeypatch.delenv(k)
    lst = ExperimentList()
    lib.create_experiment_list(C.byref(lst))
    assert lst.count == 0 and lst.next == 1
    lib.destroy_experiment_list(C.byref(lst))
    # the timing experiment of experiment_list.c:354-396: 5 estimator variants, 1000 frames each, 1920x1080
    monkeypatch.setenv("EXP_MED_ROUGH", "1"); monkeypatch.setenv("EXP_TIMINGS", "1"); monkeypatch.setenv("NUM_SAMPLES", "77")
    lib.create_experiment_list(C.byref(lst))
    assert lst.count == 5
    names = [lst.experiments[i].exp_name.decode() for i in range(5)]
    assert names == ["uniform_uniform_time", "uniform_cp_time", "uniform_area_time", "cp_cp_time", "ltc_cp_time"]
    e = lst.experiments[4]
    assert (e.width, e.height, e.num_samples) == (1920, 1080, 1000)
    assert e.render_settings.polygon_sampling_technique == 4 and e.render_settings.light_sampling == 1
    assert abs(e.render_settings.roughness_factor - 0.1) < 1e-7 and abs(e.render_settings.exposure_factor - 1.5) < 1e-7
    lib.destroy_experiment_list(C.byref(lst))
    assert lst.count == 0 and not lst.experiments
    # EXP_COMPARE uses NUM_SAMPLES
    monkeypatch.delenv("EXP_TIMINGS"); monkeypatch.setenv("EXP_COMPARE", "1")
    lib.create_experiment_list(C.byref(lst))
    assert lst.count == 5 and lst.experiments[3].num_samples == 77 and lst.experiments[3].exp_name == b"ltc_cp"
    lib.destroy_experiment_list(C.byref(lst))


def test_scene_file_round_trips(tmp_path):
    scene = scenes.many_light_room(6, 4, seed=9, width=64, height=36, vertex_count=4)
    fits = ltc_fit.fit_ggx_ltc(8, 3, 8)
    vks, tex, save = host.write_scene_files(scene, tmp_path, ltc_fits=fits)
    mesh = formats.read_vks(vks)
    for k in ("positions", "normals_uv", "material_indices"):
        assert np.array_equal(mesh[k], scene["mesh"][k])
    assert mesh["material_names"] == scene["mesh"]["material_names"]
    assert np.array_equal(np.asarray(mesh["dequant_factor"], np.float32), np.asarray(scene["mesh"]["dequant_factor"], np.float32))
    cam, lights = formats.read_quicksave(save)
    assert len(lights) == 6 and abs(cam["rotation_x"] - np.float32(scene["camera"]["rotation_x"])) < 1e-6
    for a, b in zip(lights, scene["lights"]):
        assert np.allclose(a["vertices_plane_space"][:, :2], np.asarray(b["vertices_plane_space"])[:, :2])
        assert np.allclose(a["radiant_flux"], b["radiant_flux"])
    back = formats.read_ltc_fits(tmp_path / "ggx_ltc_fit", 3)
    assert np.array_equal(np.asarray(back, np.float32), np.asarray(fits, np.float32))
    tex_img = formats.read_vkt(Path(tex) / "flat1_BaseColor.vkt")
    assert np.allclose(np.asarray(tex_img).reshape(-1)[:3], scene["materials"][1]["base_color"])
    # quicksave layout (main.c:45-125): camera 48 B + u32 legacy + u32 count, then per light 88 B fixed + ...
    raw = Path(save).read_bytes()
    assert struct.unpack_from("<I", raw, 52)[0] == 6


def test_ltc_quantisation_known_answer():
    """load_ltc_table arithmetic (ltc_table.c:82-116): adjugate of [[a,0,b],[0,c,0],[d,0,1]] / max|.| -> UNORM16."""
    from oracle import orc
    fit = (C.c_float * 5)(0.5, 0.1, 0.25, -0.2, 0.9)
    rgba, rg = (C.c_uint16 * 4)(), (C.c_uint16 * 2)()
    orc.lib().orc_quantize_ltc_fit(fit, rgba, rg)
    a, b, c, d = 0.5, 0.1, 0.25, -0.2
    inv = np.array([[c, 0, -b * c], [0, a - b * d, 0], [-c * d, 0, a * c]])   # adjugate, C array inverse[k][l]
    inv = inv / np.abs(inv).max()
    want = [inv[0, 0], -inv[0, 2], inv[1, 1], inv[2, 0], inv[2, 2], 0.9]
    got = [x / 65535.0 for x in list(rgba) + list(rg)]
    assert np.allclose(got, np.clip(want, 0, 1), atol=1.0 / 65535.0)
    # the numpy quantiser used for generated tables is the same arithmetic
    q_rgba, q_rg = ltc_fit.quantize_fits(np.array([[[[0.5, 0.1, 0.25, -0.2, 0.9]]]], dtype=np.float32))
    assert list(q_rgba.reshape(-1)) == list(rgba) and list(q_rg.reshape(-1)) == list(rg)


def test_hdr_writer(lib, tmp_path):
    rgba = np.zeros((4, 6, 4), dtype=np.float32)
    rgba[..., 0] = 0.5; rgba[..., 1] = 2.0; rgba[..., 2] = 0.125; rgba[..., 3] = 1.0
    path = tmp_path / "shot.hdr"
    assert lib.write_hdr_screenshot(str(path).encode(), rgba.ctypes.data_as(C.c_void_p), C.c_uint32(6), C.c_uint32(4)) == 0
    raw = path.read_bytes()
    assert raw.startswith(b"#?RADIANCE") and b"-Y 4 +X 6" in raw
    px = raw[-4 * 24:][:4]
    e = px[3] - 136
    assert np.allclose([px[0] * 2.0 ** e, px[1] * 2.0 ** e, px[2] * 2.0 ** e], [0.5, 2.0, 0.125], rtol=0.02)


def test_block_compressed_textures(lib, tmp_path):
    """tables_scene.c decodes BC1 / BC5 blocks (the formats of the reference's texture conversion tool,
    tools/texture_conversion/main.c:32-38) exactly like the independent numpy decoder, partial blocks included; *.vkt files
    with mip chains round-trip; the oracle's sampler returns a flat texture's texel exactly."""
    from oracle import orc
    from risltc_b200 import formats, scenes
    rng = np.random.default_rng(3)
    for bc5, (w, h) in ((0, (20, 12)), (1, (20, 12)), (0, (5, 7)), (1, (2, 1)), (0, (64, 64))):
        n = ((w + 3) // 4) * ((h + 3) // 4) * (16 if bc5 else 8)
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        out = np.zeros((h, w, 4), dtype=np.uint8)
        lib.decode_block_compressed_texels(out.ctypes.data_as(C.c_void_p), data, C.c_uint32(w), C.c_uint32(h), C.c_int(bc5))
        assert np.array_equal(out, formats.decode_bc5(data, w, h) if bc5 else formats.decode_bc1(data, w, h)), (bc5, w, h)
    scene = scenes.add_procedural_textures(scenes.many_light_room(4, 2, width=32, height=18), size=32)
    base, spec, nrm = scene["textures"][:3]
    for tex, fmt, channels, tolerance in ((base, formats.VK_FORMAT_BC1_RGB_SRGB_BLOCK, 3, 24), (spec, formats.VK_FORMAT_BC1_RGB_UNORM_BLOCK, 3, 24), (nrm, formats.VK_FORMAT_BC5_UNORM_BLOCK, 2, 20)):
        path = tmp_path / "t.vkt"
        formats.write_vkt(path, tex["levels"], fmt)
        levels, got_fmt = formats.read_vkt(path, with_format=True)
        assert got_fmt == fmt and [l.shape for l in levels] == [l.shape for l in tex["levels"]]
        for a, b in zip(levels, tex["levels"]):
            assert np.abs(a[..., :channels].astype(int) - b[..., :channels].astype(int)).max() <= tolerance
    # sampler properties: a flat texture returns its texel exactly at any footprint; level of detail picks the 1x1 level
    flat = dict(format="rgba32f", levels=[np.full((4, 4, 4), 0.3217, np.float32), np.full((2, 2, 4), 0.3217, np.float32), np.full((1, 1, 4), 0.3217, np.float32)])
    for uv, dx in (((0.3, 0.9), 0.01), ((-3.7, 12.2), 0.2), ((0.5, 0.5), 7.0)):
        assert np.all(orc.sample_texture_grad(flat, uv, (dx, 0.0), (0.0, dx)) == np.float32(0.3217))
    ramp = dict(format="rgba8_unorm", levels=scenes.mip_chain(np.tile(np.arange(0, 256, 32, dtype=np.uint8)[None, :, None], (8, 1, 4))))
    coarse = orc.sample_texture_grad(ramp, (0.5, 0.5), (4.0, 0.0), (0.0, 4.0))
    assert np.allclose(coarse, ramp["levels"][-1][0, 0] / 255.0)
    assert abs(orc.sample_texture_grad(ramp, (0.5, 0.5), (1e-4, 0.0), (0.0, 1e-4))[0] - (96 + 128) / 2 / 255.0) < 1e-6
    assert orc.lib().orc_srgb8_to_linear.restype is not None


def test_screenshot_encoders_round_trip(lib, tmp_path):
    """screenshot.c: PNG (stored deflate, CRCs, Adler-32) and run-length Radiance files decode to what was stored; the two
    half-bit frames combine to the fp16 image (main.c:2339-2350); the oracle's copy pass supplies the frames."""
    from oracle import orc
    from risltc_b200 import formats
    rng = np.random.default_rng(5)
    H, W = 37, 301
    rgba = (rng.random((H, W, 4)) * np.exp(rng.uniform(-9, 4, (H, W, 1)))).astype(np.float32)
    rgba[5:9, 10:200] = rgba[5, 10]          # long runs
    rgba[20, :, :3] = 0.0
    display, low, high = orc.copy_pass(rgba, 0), orc.copy_pass(rgba, 1), orc.copy_pass(rgba, 2)
    png = tmp_path / "shot.png"
    assert lib.write_png(str(png).encode(), display.ctypes.data_as(C.c_void_p), C.c_uint32(W), C.c_uint32(H)) == 0
    assert np.array_equal(formats.read_png(png), display)
    hdr = np.zeros((H, W, 3), dtype=np.float32)
    lib.combine_ldr_screenshots_into_hdr(hdr.ctypes.data_as(C.c_void_p), low.ctypes.data_as(C.c_void_p), high.ctypes.data_as(C.c_void_p), C.c_size_t(hdr.size))
    want = rgba[..., :3].astype(np.float16).astype(np.float32)
    assert np.array_equal(hdr.view(np.uint32), want.view(np.uint32))
    path = tmp_path / "shot.hdr"
    assert lib.write_hdr(str(path).encode(), hdr.ctypes.data_as(C.c_void_p), C.c_uint32(W), C.c_uint32(H)) == 0
    got = formats.read_hdr(path)
    tolerance = want.max(axis=-1, keepdims=True) * 2.0 ** -7 + 1e-30
    assert np.all(np.abs(got - want) <= tolerance)
    assert path.stat().st_size < 4 * W * H      # the runs were found
    # narrow images are stored flat, like stb does
    assert lib.write_hdr(str(path).encode(), hdr[:, :5].copy().ctypes.data_as(C.c_void_p), C.c_uint32(5), C.c_uint32(H)) == 0
    assert np.all(np.abs(formats.read_hdr(path) - want[:, :5]) <= tolerance[:, :5])
    # srgb_utility.glsl:20-34 against a float64 evaluation of the same curve
    lin = np.clip(rgba[..., :3].astype(np.float64), 0.0, 1.0)
    srgb = np.where(lin <= 0.0031308, 12.92 * lin, 1.055 * lin ** (1.0 / 2.4) - 0.055)
    assert np.abs(display.astype(int) - np.floor(srgb * 255.0 + 0.5).astype(int)).max() <= 1


@pytest.mark.parametrize("max_leaf", [1, 2, 4])
def test_acceleration_structures_hold_their_invariants(max_leaf):
    """Host-only: the binary BVH and its 4-wide collapse with 8-bit boxes (risltc_b200/csrc/bvh_build.cpp) for a generated
    scene and for degenerate soups -- every triangle in exactly one leaf, every vertex inside its leaf box, every quantised
    box containing the binary tree's box, every node referenced once. Hit / no-hit of the traversal kernels only depends on
    the tree through these properties (the triangle test itself is the oracle's)."""
    from risltc_b200 import api, scenes
    scene = scenes.many_light_room(16, 60, seed=8, occluder_triangles=20000, width=64, height=36)
    verts = scenes.dequantize_positions(scene["mesh"]).astype(np.float32).reshape(-1, 3, 3)
    rng = np.random.default_rng(3)
    soups = [verts,
             verts[:1],                                                     # the whole scene is one leaf
             np.repeat(verts[:1], 40, axis=0),                              # coincident triangles
             (rng.normal(size=(3000, 1, 3)) * 50 + rng.normal(size=(3000, 3, 3)) * rng.uniform(1e-4, 5.0, (3000, 1, 1))).astype(np.float32),
             np.concatenate([verts[:200], verts[:200] * np.float32(1e-3) + np.float32(900.0)])]   # tiny geometry far from the origin
    for soup in soups:
        r = api.check_bvh(soup, max_leaf)
        assert r["bad_order"] == 0 and r["bad_binary"] == 0 and r["bad_wide"] == 0, r
        assert r["wide_nodes"] <= r["binary_nodes"] and 1.0 <= r["children_per_node"] <= 4.0
        assert 3 * r["depth"] + 1 <= 88      # the traversal stack of trace4_kernel (RL_T4_OVERFLOW)


def test_host_builder_is_the_same_tree_on_one_thread_and_on_many(monkeypatch):
    """The host builder hands the two subtrees of the top levels of a large scene to two tasks and splices their nodes in
    preorder (bvh_build.cpp): binary tree, triangle order and 4-wide tree are byte for byte those of the one-thread build."""
    from risltc_b200 import api
    rng = np.random.default_rng(5)
    T = 260_000      # above the 100 k triangles from which subtrees are built by tasks of their own
    soup = (rng.uniform(-20, 20, (T, 1, 3)) + rng.normal(size=(T, 3, 3)) * rng.uniform(0.01, 0.4, (T, 1, 1))).astype(np.float32)
    monkeypatch.setenv("RISLTC_BVH_THREADS", "1")
    one = api.bvh_checksum(soup)
    monkeypatch.delenv("RISLTC_BVH_THREADS")
    many = api.bvh_checksum(soup)
    assert one == many
    r = api.check_bvh(soup[:120_000], 2)
    assert r["bad_order"] == 0 and r["bad_binary"] == 0 and r["bad_wide"] == 0
