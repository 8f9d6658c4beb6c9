"""ctypes front end of oracle/_ref/ (the reference's own sources compiled for the CPU by
oracle/build_ref.py). TEST INFRASTRUCTURE. /root/reference only exists in the build container; on
the GPU box the prebuilt libraries in oracle/_ref/ are used as they are."""
import ctypes as C
from pathlib import Path

import numpy as np

from . import orc

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"


def available(name):
    return (REF_DIR / f"libref_shading_{name}.so").exists()


def host_available():
    return (REF_DIR / "libref_host.so").exists()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefShading:
    """One compiled variant of the reference's shading_pass.frag.glsl."""

    def __init__(self, name):
        path = REF_DIR / f"libref_shading_{name}.so"
        if not path.exists():
            raise FileNotFoundError(f"{path}: run `python oracle/build_ref.py` where /root/reference exists")
        self.lib = C.CDLL(str(path))
        self.lib.ref_shade_rows.restype = C.c_uint64
        self.lib.ref_calculate_ltc.restype = C.c_float
        self.lib.ref_clip_polygon.restype = C.c_uint32
        self.lib.ref_max_polygon_vertex_count.restype = C.c_uint32
        self.scene = None

    @property
    def max_polygon_vertices(self):
        return int(self.lib.ref_max_polygon_vertex_count())

    def bind(self, oscene):
        """oscene: orc.OracleScene (provides the buffers and the BVH behind the any-hit callback)."""
        self.scene = oscene
        any_hit = C.cast(orc.lib().orc_any_hit, C.c_void_p)
        self.lib.ref_bind_scene(_p(oscene.positions), _p(oscene.normals_uv), _p(oscene.material_indices),
                                _p(oscene.materials), C.c_uint32(oscene.materials.shape[0]),
                                _p(oscene.records), C.c_uint32(oscene.records.shape[0]), C.c_uint32(oscene.records.shape[1]),
                                _p(oscene.rgba16), _p(oscene.rg16), C.c_uint32(oscene.rgba16.shape[1]), C.c_uint32(oscene.rgba16.shape[0]),
                                any_hit, C.byref(oscene.c))
        if oscene.textures is not None:
            self.lib.ref_bind_textures(C.cast(oscene.textures, C.c_void_p), C.c_uint32(C.sizeof(orc.Texture)), C.c_uint32(len(oscene.textures)),
                                       C.cast(orc.lib().orc_sample_texture_grad, C.c_void_p))
        else:
            self.lib.ref_bind_textures(None, C.c_uint32(0), C.c_uint32(0), None)

    def render(self, constants_list, accum=None, accum_start=0):
        """Visibility from the oracle (driver stand-in), shading by the compiled reference, accumulation
        by the oracle's accum pass. Returns (accum, visibility, rays)."""
        W, H = constants_list[0].viewport[0], constants_list[0].viewport[1]
        if accum is None:
            accum = np.zeros((H, W, 4), dtype=np.float32)
        vis = np.zeros((H, W), dtype=np.uint32)
        shaded = np.zeros((H, W, 4), dtype=np.float32)
        rays = 0
        for k, c in enumerate(constants_list):
            orc.lib().orc_visibility_pass(C.byref(self.scene.c), C.byref(c), _p(vis), C.c_uint32(0), C.c_uint32(H))
            self.lib.ref_set_constants(C.byref(c))
            rays += int(self.lib.ref_shade_rows(_p(vis), _p(shaded), C.c_uint32(0), C.c_uint32(H)))
            orc.lib().orc_accum_pass(_p(accum), _p(shaded), C.c_uint32(accum_start + k), C.c_uint64(W * H))
        return accum, vis, rays

    def shade(self, constants, vis):
        H, W = vis.shape
        shaded = np.zeros((H, W, 4), dtype=np.float32)
        self.lib.ref_set_constants(C.byref(constants))
        rays = int(self.lib.ref_shade_rows(_p(np.ascontiguousarray(vis)), _p(shaded), C.c_uint32(0), C.c_uint32(H)))
        return shaded, rays

    # ---- function level
    def clip_polygon(self, vertex_count, v):
        buf = np.ascontiguousarray(v, dtype=np.float32).copy()
        vc = int(self.lib.ref_clip_polygon(C.c_uint32(vertex_count), _p(buf)))
        return vc, buf

    def calculate_ltc(self, vertex_count, v):
        buf = np.ascontiguousarray(v, dtype=np.float32)
        return float(self.lib.ref_calculate_ltc(C.c_uint32(vertex_count), _p(buf)))

    def psa(self, vertex_count, v, u0, u1):
        buf = np.ascontiguousarray(v, dtype=np.float32)
        poly = np.zeros(44, dtype=np.float32); d = np.zeros(3, dtype=np.float32)
        self.lib.ref_psa(C.c_uint32(vertex_count), _p(buf), C.c_float(u0), C.c_float(u1), _p(poly), _p(d))
        return poly, d

    def ltc_coefficients(self, fresnel_0, roughness, pos, normal, outgoing, constants6):
        out = np.zeros(32, dtype=np.float32)
        f3 = lambda a: (C.c_float * 3)(*[float(x) for x in a])
        self.lib.ref_ltc_coefficients(C.c_float(fresnel_0), C.c_float(roughness), f3(pos), f3(normal), f3(outgoing),
                                      (C.c_float * 6)(*[float(x) for x in constants6]), _p(out))
        return out

    def noise(self, px, py, width, frame_word, draws):
        out = np.zeros(draws, dtype=np.float32)
        self.lib.ref_noise(C.c_uint32(px), C.c_uint32(py), C.c_uint32(width), C.c_uint32(frame_word), C.c_uint32(draws), _p(out))
        return out


class RefHost:
    """polygonal_light.c, camera.c and math_utilities.h of the reference, compiled as-is."""

    class PolygonalLight(C.Structure):   # polygonal_light.h:73-99
        _fields_ = [("rotation_angles", C.c_float * 3), ("scaling_x", C.c_float), ("translation", C.c_float * 3), ("scaling_y", C.c_float),
                    ("radiant_flux", C.c_float * 3), ("inv_scaling_x", C.c_float), ("surface_radiance", C.c_float * 3), ("inv_scaling_y", C.c_float),
                    ("plane", C.c_float * 4), ("vertex_count", C.c_uint32), ("texturing_technique", C.c_int), ("texture_index", C.c_uint32),
                    ("padding_0", C.c_uint32), ("rotation", (C.c_float * 4) * 3), ("area", C.c_float), ("rcp_area", C.c_float),
                    ("padding_1", C.c_float * 2), ("texture_file_path", C.c_char_p), ("vertices_plane_space", C.POINTER(C.c_float)),
                    ("vertices_world_space", C.POINTER(C.c_float))]

    class Camera(C.Structure):   # camera.h:29-49
        _fields_ = [("position_world_space", C.c_float * 3), ("rotation_z", C.c_float), ("rotation_x", C.c_float), ("vertical_fov", C.c_float),
                    ("near", C.c_float), ("far", C.c_float), ("speed", C.c_float), ("rotate_camera", C.c_int),
                    ("rotation_x_0", C.c_float), ("rotation_z_0", C.c_float)]

    def __init__(self, path=None):
        self.lib = C.CDLL(str(path or (REF_DIR / "libref_host.so")))
        if path is None:   # the helper exports only exist in oracle/_ref; the struct-level functions are shared names
            self.lib.ref_wang_random_number.restype = C.c_uint32
            self.lib.ref_half_to_float.restype = C.c_float

    def make_light(self, light):
        n = len(light["vertices_plane_space"])
        l = self.PolygonalLight()
        for i in range(3):
            l.rotation_angles[i] = light["rotation_angles"][i]; l.translation[i] = light["translation"][i]; l.radiant_flux[i] = light["radiant_flux"][i]
        l.scaling_x, l.scaling_y = light["scaling_x"], light["scaling_y"]
        self.lib.set_polygonal_light_vertex_count(C.byref(l), C.c_uint32(n))
        for i in range(n):
            l.vertices_plane_space[4 * i] = float(light["vertices_plane_space"][i][0])
            l.vertices_plane_space[4 * i + 1] = float(light["vertices_plane_space"][i][1])
        return l

    def update_light(self, light):
        """Returns (world vertices (n,4), plane (4,), surface_radiance (3,), area, rotation (3,4)) after update_polygonal_light."""
        l = self.make_light(light)
        self.lib.update_polygonal_light(C.byref(l))
        n = l.vertex_count
        world = np.array([l.vertices_world_space[i] for i in range(4 * n)], dtype=np.float32).reshape(n, 4)
        out = (world, np.array(list(l.plane), dtype=np.float32), np.array(list(l.surface_radiance), dtype=np.float32), float(l.area),
               np.array([list(r) for r in l.rotation], dtype=np.float32))
        self.lib.destroy_polygonal_light(C.byref(l))
        return out

    def world_to_projection(self, cam, aspect):
        c = self.Camera()
        for i in range(3):
            c.position_world_space[i] = cam["position"][i]
        c.rotation_z, c.rotation_x, c.vertical_fov, c.near, c.far = cam["rotation_z"], cam["rotation_x"], cam["vertical_fov"], cam["near"], cam["far"]
        out = ((C.c_float * 4) * 4)()
        self.lib.get_world_to_projection_space(out, C.byref(c), C.c_float(aspect))
        return np.array([list(r) for r in out], dtype=np.float32)

    def matrix_inverse(self, m):
        a = ((C.c_float * 4) * 4)(*[(C.c_float * 4)(*[float(x) for x in row]) for row in m])
        out = ((C.c_float * 4) * 4)()
        self.lib.ref_matrix_inverse(out, a)
        return np.array([list(r) for r in out], dtype=np.float32)

    def wang(self, seed):
        return int(self.lib.ref_wang_random_number(C.c_uint32(seed)))
