"""ctypes front end of the CPU oracle (oracle/liborc.so). TEST INFRASTRUCTURE:
imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
MAX_P = 8

MIS = dict(balance=0, power=1, weighted=2, optimal_clamped=3, optimal=4)
LIGHT = dict(uniform=0, reservoir=1)
POLY = dict(baseline=0, area_turk=1, projected_solid_angle=2, projected_solid_angle_biased=3, ltc_cp=4)


def build(force=False):
    lib = HERE / "liborc.so"
    srcs = [HERE / n for n in ("risltc_oracle.c", "risltc_oracle_frame.inc", "risltc_oracle.h", "clip_rotation_table.h")]
    if force or not lib.exists() or any(s.stat().st_mtime > lib.stat().st_mtime for s in srcs):
        subprocess.check_call(["make", "-s", "-C", str(HERE), "liborc.so"], env={**os.environ, "CC": "gcc"})
    return lib


class Variant(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("light_sampling", "polygon_technique", "mis_heuristic", "sample_count",
                                          "light_samples", "fast_atan", "min_light_vertices", "max_light_vertices")]


class Constants(C.Structure):
    _fields_ = [("dequant_factor", C.c_float * 3), ("pad0", C.c_float), ("dequant_summand", C.c_float * 3),
                ("error_factor", C.c_float), ("world_to_projection", (C.c_float * 4) * 4),
                ("pixel_to_ray", (C.c_float * 4) * 3), ("camera_position", C.c_float * 3),
                ("mis_visibility_estimate", C.c_float), ("viewport", C.c_uint32 * 2), ("cursor", C.c_int32 * 2),
                ("exposure_factor", C.c_float), ("roughness_factor", C.c_float),
                ("noise_resolution_mask", C.c_uint32 * 2), ("noise_texture_index_mask", C.c_uint32),
                ("pad3", C.c_uint32 * 3), ("noise_random_numbers", C.c_uint32 * 4), ("ltc_constants", C.c_float * 8)]


assert C.sizeof(Constants) == 256


class Scene(C.Structure):
    _fields_ = [("triangle_count", C.c_uint64), ("dequant_factor", C.c_float * 3), ("dequant_summand", C.c_float * 3),
                ("positions", C.c_void_p), ("normals_uv", C.c_void_p), ("material_indices", C.c_void_p),
                ("material_count", C.c_uint64), ("materials", C.c_void_p), ("light_count", C.c_uint32),
                ("light_records", C.c_void_p), ("ltc_res", C.c_uint32), ("ltc_layers", C.c_uint32),
                ("ltc_rgba16", C.c_void_p), ("ltc_rg16", C.c_void_p), ("bvh", C.c_void_p), ("textures", C.c_void_p)]


class Texture(C.Structure):
    _fields_ = [("format", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("mip_count", C.c_uint32), ("texels", C.c_void_p)]


TEXEL = dict(rgba32f=0, rgba8_unorm=1, rgba8_srgb=2)


def pack_textures(textures):
    """textures: list of dicts {format: 'rgba32f' | 'rgba8_unorm' | 'rgba8_srgb', levels: [(h, w, 4) arrays, largest first]}.
    Returns (ctypes array of Texture, list of the packed numpy buffers that must stay alive)."""
    arr = (Texture * len(textures))()
    keep = []
    for i, t in enumerate(textures):
        dtype = np.float32 if t["format"] == "rgba32f" else np.uint8
        buf = np.concatenate([np.ascontiguousarray(l, dtype=dtype).reshape(-1) for l in t["levels"]])
        keep.append(buf)
        arr[i].format = TEXEL[t["format"]]
        arr[i].height, arr[i].width = t["levels"][0].shape[:2]
        arr[i].mip_count = len(t["levels"])
        arr[i].texels = buf.ctypes.data
    return arr, keep


def sample_texture_grad(texture, uv, ddx, ddy):
    arr, keep = pack_textures([texture])
    out = (C.c_float * 4)()
    lib().orc_sample_texture_grad(C.byref(arr[0]), (C.c_float * 2)(*uv), (C.c_float * 2)(*ddx), (C.c_float * 2)(*ddy), out)
    return np.array(list(out), dtype=np.float32)


class PsaPolygon(C.Structure):
    _fields_ = [("vertex_count", C.c_uint32), ("vertices", (C.c_float * 2) * MAX_P), ("ellipses", (C.c_float * 2) * MAX_P),
                ("inner_ellipse_0", C.c_float * 2), ("sector_projected_solid_angles", C.c_float * MAX_P),
                ("projected_solid_angle", C.c_float)]


class Ltc(C.Structure):
    _fields_ = [("world_to_shading", (C.c_float * 3) * 4), ("shading_to_cosine", (C.c_float * 3) * 3),
                ("cosine_to_shading", (C.c_float * 3) * 3), ("albedo", C.c_float), ("determinant", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.orc_noise_next.restype = C.c_float
        _lib.orc_calculate_ltc.restype = C.c_float
        _lib.orc_render_frame.restype = C.c_uint64
        _lib.orc_wang_random_number.restype = C.c_uint32
        _lib.orc_noise_seed.restype = C.c_uint32
        _lib.orc_clip_polygon.restype = C.c_uint32
        _lib.orc_srgb8_to_linear.restype = C.c_float
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def variant(light_sampling="reservoir", technique="ltc_cp", mis="optimal_clamped", sample_count=1, light_samples=1,
            fast_atan=0, min_vertices=3, max_vertices=3):
    return Variant(LIGHT[light_sampling], POLY[technique], MIS[mis], sample_count, light_samples, int(fast_atan),
                   min_vertices, max_vertices)


def wang(seed):
    return int(lib().orc_wang_random_number(C.c_uint32(seed & 0xFFFFFFFF)))


def light_records(lights, max_vertices=None):
    """The write_lights stream (main.c:456-490) via orc_update_polygonal_light; returns (N, 12 + 4 V) float32."""
    if max_vertices is None:
        max_vertices = max(len(l["vertices_plane_space"]) for l in lights)
    out = np.zeros((len(lights), 12 + 4 * max_vertices), dtype=np.float32)
    for i, l in enumerate(lights):
        n = len(l["vertices_plane_space"])
        ps = np.zeros((n, 4), dtype=np.float32); ps[:, :2] = np.asarray(l["vertices_plane_space"], dtype=np.float32)[:, :2]
        world = np.zeros((n, 4), dtype=np.float32)
        plane = (C.c_float * 4)(); rad = (C.c_float * 3)(); area = C.c_float()
        lib().orc_update_polygonal_light((C.c_float * 3)(*l["rotation_angles"]), C.c_float(l["scaling_x"]), C.c_float(l["scaling_y"]),
                                         (C.c_float * 3)(*l["translation"]), (C.c_float * 3)(*l["radiant_flux"]), C.c_uint32(n),
                                         _p(ps), _p(world), plane, rad, C.byref(area), None)
        out[i, 0:3] = list(rad); out[i, 4:8] = list(plane)
        out[i, 8:9].view(np.uint32)[0] = n
        out[i, 12:12 + 4 * n] = world.reshape(-1)
        if n < max_vertices:
            out[i, 12 + 4 * n:16 + 4 * n] = world[0]
    return out


def make_constants(scene, width, height, frame_word, exposure=1.5, roughness_factor=1.0, mis_visibility_estimate=0.5,
                   ltc_res=64, ltc_layers=51):
    """write_constants (main.c:2902-2946) through the oracle's host arithmetic."""
    c = Constants()
    mesh, cam = scene["mesh"], scene["camera"]
    for i in range(3):
        c.dequant_factor[i] = float(mesh["dequant_factor"][i]); c.dequant_summand[i] = float(mesh["dequant_summand"][i])
        c.camera_position[i] = float(cam["position"][i])
    c.error_factor = float(np.float32(10.0) ** np.float32(7.0))
    lib().orc_world_to_projection(c.world_to_projection, (C.c_float * 3)(*cam["position"]), C.c_float(cam["rotation_x"]),
                                  C.c_float(cam["rotation_z"]), C.c_float(cam["vertical_fov"]), C.c_float(cam["near"]),
                                  C.c_float(cam["far"]), C.c_float(np.float32(width) / np.float32(height)))
    lib().orc_pixel_to_ray(c.pixel_to_ray, c.world_to_projection, C.c_uint32(width), C.c_uint32(height))
    c.mis_visibility_estimate = mis_visibility_estimate
    c.viewport[0], c.viewport[1] = width, height
    c.exposure_factor, c.roughness_factor = exposure, roughness_factor
    c.noise_random_numbers[0] = frame_word & 0xFFFFFFFF
    lib().orc_ltc_constants(c.ltc_constants, C.c_uint32(ltc_res), C.c_uint32(ltc_res), C.c_uint32(ltc_layers))
    return c


def frame_words(frame_seed):
    """set_noise_constants with animate_noise (noise_table.c:24-28): 4 words for frame number `frame_seed`."""
    return [wang(frame_seed * 4 + i) for i in range(4)]


class OracleScene:
    """Owns the numpy buffers behind an orc_scene_t and its BVH."""

    def __init__(self, scene, ltc_rgba16, ltc_rg16, max_vertices=None, arrays=None):
        """scene: a risltc_b200.scenes dict; or arrays = dict(positions, normals_uv, material_indices,
        dequant_factor, dequant_summand, materials (M,8), records (N,12+4V), min_vertices) as stored in tests/golden."""
        if arrays is None:
            from risltc_b200.scenes import material_constants
            mesh = scene["mesh"]
            arrays = dict(mesh, materials=material_constants(scene["materials"]), records=light_records(scene["lights"], max_vertices),
                          min_vertices=min(len(l["vertices_plane_space"]) for l in scene["lights"]))
        mesh = arrays
        self.positions = np.ascontiguousarray(mesh["positions"], dtype=np.uint32)
        self.normals_uv = np.ascontiguousarray(mesh["normals_uv"], dtype=np.uint16)
        self.material_indices = np.ascontiguousarray(mesh["material_indices"], dtype=np.uint8)
        self.materials = np.ascontiguousarray(arrays["materials"], dtype=np.float32)
        self.records = np.ascontiguousarray(arrays["records"], dtype=np.float32)
        self.max_vertices = (self.records.shape[1] - 12) // 4
        self.min_vertices = int(arrays["min_vertices"])
        self.rgba16 = np.ascontiguousarray(ltc_rgba16, dtype=np.uint16)
        self.rg16 = np.ascontiguousarray(ltc_rg16, dtype=np.uint16)
        s = Scene()
        s.triangle_count = self.material_indices.shape[0]
        for i in range(3):
            s.dequant_factor[i] = float(mesh["dequant_factor"][i]); s.dequant_summand[i] = float(mesh["dequant_summand"][i])
        s.positions, s.normals_uv, s.material_indices = _p(self.positions), _p(self.normals_uv), _p(self.material_indices)
        s.material_count, s.materials = self.materials.shape[0], _p(self.materials)
        s.light_count, s.light_records = self.records.shape[0], _p(self.records)
        s.ltc_layers, s.ltc_res = self.rgba16.shape[0], self.rgba16.shape[1]
        s.ltc_rgba16, s.ltc_rg16 = _p(self.rgba16), _p(self.rg16)
        # material textures (3 per material) when the scene has them; flat materials otherwise
        self.textures = None
        tex = arrays.get("textures") if isinstance(arrays, dict) else None
        if tex is None and scene is not None:
            tex = scene.get("textures")
        if tex is not None:
            self.textures, self._texture_buffers = pack_textures(tex)
            s.textures = C.cast(self.textures, C.c_void_p)
        self.c = s
        lib().orc_build_bvh(C.byref(s))

    def __del__(self):
        try:
            lib().orc_free_bvh(C.byref(self.c))
        except Exception:
            pass

    def render(self, constants_list, var, accum=None, accum_start=0):
        """Render one frame per entry of constants_list, accumulating like accum_pass. Returns (accum, visibility, rays)."""
        W, H = constants_list[0].viewport[0], constants_list[0].viewport[1]
        if accum is None:
            accum = np.zeros((H, W, 4), dtype=np.float32)
        vis = np.zeros((H, W), dtype=np.uint32)
        rays = 0
        for k, c in enumerate(constants_list):
            rays += int(lib().orc_render_frame(C.byref(self.c), C.byref(c), C.byref(var), C.c_uint32(accum_start + k), _p(accum), _p(vis)))
        return accum, vis, rays

    def any_hit(self, origin, direction, t_min, t_max):
        return int(lib().orc_any_hit(C.byref(self.c), (C.c_float * 3)(*origin), (C.c_float * 3)(*direction), C.c_float(t_min), C.c_float(t_max)))


def thread_count():
    return int(lib().orc_thread_count())


# ---- function-level wrappers (same shapes as oracle/ref.py and the risltc_cuda_kat_* entry points)
def clip_polygon(vertex_count, v, min_vertices, max_polygon_vertices):
    buf = np.zeros((MAX_P, 3), dtype=np.float32); buf[:] = np.asarray(v, dtype=np.float32).reshape(MAX_P, 3)
    vc = int(lib().orc_clip_polygon(C.c_uint32(vertex_count), _p(buf), C.c_uint32(min_vertices), C.c_uint32(max_polygon_vertices)))
    return vc, buf


def calculate_ltc(vertex_count, v):
    buf = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(MAX_P, 3))
    return float(lib().orc_calculate_ltc(C.c_uint32(vertex_count), _p(buf)))


def psa(vertex_count, v, u0, u1, max_polygon_vertices, fast_atan=0, biased=0):
    """Returns (44 floats {vc, v[8][2], e[8][2], inner0[2], sectors[8], total}, sampled direction)."""
    buf = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(MAX_P, 3))
    poly = PsaPolygon()
    lib().orc_prepare_psa(C.byref(poly), C.c_uint32(vertex_count), _p(buf), C.c_uint32(max_polygon_vertices), C.c_uint32(fast_atan))
    d = np.zeros(3, dtype=np.float32)
    lib().orc_sample_psa(_p(d), C.byref(poly), C.c_float(u0), C.c_float(u1), C.c_uint32(max_polygon_vertices), C.c_uint32(fast_atan), C.c_uint32(biased))
    out = np.zeros(44, dtype=np.float32)
    out[0] = poly.vertex_count
    for k in range(min(vertex_count, MAX_P)):
        out[1 + 2 * k], out[2 + 2 * k] = poly.vertices[k][0], poly.vertices[k][1]
        out[17 + 2 * k], out[18 + 2 * k] = poly.ellipses[k][0], poly.ellipses[k][1]
        out[35 + k] = poly.sector_projected_solid_angles[k]
    out[33], out[34] = poly.inner_ellipse_0[0], poly.inner_ellipse_0[1]
    out[43] = poly.projected_solid_angle
    return out, d


def ltc_coefficients(oscene, fresnel_0, roughness, pos, normal, outgoing, constants6):
    """32 floats: world_to_shading[12] (column-major), shading_to_cosine[9], cosine_to_shading[9], albedo, determinant."""
    l = Ltc()
    f3 = lambda a: (C.c_float * 3)(*[float(x) for x in a])
    lib().orc_get_ltc_coefficients(C.byref(l), C.byref(oscene.c), C.c_float(fresnel_0), C.c_float(roughness), f3(pos), f3(normal), f3(outgoing),
                                   (C.c_float * 6)(*[float(x) for x in constants6]))
    out = np.zeros(32, dtype=np.float32)
    out[0:12] = np.array([list(col) for col in l.world_to_shading], dtype=np.float32).reshape(-1)
    out[12:21] = np.array([list(col) for col in l.shading_to_cosine], dtype=np.float32).reshape(-1)
    out[21:30] = np.array([list(col) for col in l.cosine_to_shading], dtype=np.float32).reshape(-1)
    out[30], out[31] = l.albedo, l.determinant
    return out


def noise(px, py, width, frame_word, draws):
    seed = C.c_uint32(lib().orc_noise_seed(C.c_uint32(px), C.c_uint32(py), C.c_uint32(width), C.c_uint32(frame_word)))
    return np.array([lib().orc_noise_next(C.byref(seed)) for _ in range(draws)], dtype=np.float32)


def update_light(light):
    """(world (n,4), plane (4,), surface_radiance (3,), area) of polygonal_light.c:44-98."""
    n = len(light["vertices_plane_space"])
    ps = np.zeros((n, 4), dtype=np.float32); ps[:, :2] = np.asarray(light["vertices_plane_space"], dtype=np.float32)[:, :2]
    world = np.zeros((n, 4), dtype=np.float32)
    plane = (C.c_float * 4)(); rad = (C.c_float * 3)(); area = C.c_float()
    lib().orc_update_polygonal_light((C.c_float * 3)(*light["rotation_angles"]), C.c_float(light["scaling_x"]), C.c_float(light["scaling_y"]),
                                     (C.c_float * 3)(*light["translation"]), (C.c_float * 3)(*light["radiant_flux"]), C.c_uint32(n),
                                     _p(ps), _p(world), plane, rad, C.byref(area), None)
    return world, np.array(list(plane), dtype=np.float32), np.array(list(rad), dtype=np.float32), float(area.value)


def world_to_projection(cam, aspect):
    out = ((C.c_float * 4) * 4)()
    lib().orc_world_to_projection(out, (C.c_float * 3)(*cam["position"]), C.c_float(cam["rotation_x"]), C.c_float(cam["rotation_z"]),
                                  C.c_float(cam["vertical_fov"]), C.c_float(cam["near"]), C.c_float(cam["far"]), C.c_float(aspect))
    return np.array([list(r) for r in out], dtype=np.float32)


def copy_pass(rgba, frame_bits=0):
    """copy_pass.frag.glsl on an (H, W, 4) float32 frame -> (H, W, 3) uint8."""
    a = np.ascontiguousarray(rgba, dtype=np.float32)
    out = np.empty(a.shape[:-1] + (3,), dtype=np.uint8)
    lib().orc_copy_pass(_p(a), C.c_uint64(a.size // 4), C.c_uint32(frame_bits), _p(out))
    return out
