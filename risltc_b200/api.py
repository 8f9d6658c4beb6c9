"""ctypes plumbing over the C ABI (include/risltc_cuda.h). The product is the shared
library; this module only moves numpy buffers across the boundary for tests,
bench.py and the multi-GPU driver. It fails loudly when librisltc_cuda.so is
missing or no CUDA device is usable: there is no CPU path."""
import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
_lib = None

MIS = dict(balance=0, power=1, weighted=2, optimal_clamped=3, optimal=4)
LIGHT = dict(uniform=0, reservoir=1)
POLY = dict(baseline=0, area_turk=1, projected_solid_angle=2, projected_solid_angle_biased=3, ltc_cp=4)


class Variant(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("light_sampling", "polygon_technique", "mis_heuristic", "sample_count",
                                          "light_samples", "fast_atan", "min_light_vertices", "max_light_vertices")]


def variant(light_sampling="reservoir", technique="ltc_cp", mis="optimal_clamped", sample_count=1, light_samples=1,
            fast_atan=0, min_vertices=3, max_vertices=3):
    return Variant(LIGHT[light_sampling], POLY[technique], MIS[mis], sample_count, light_samples, int(fast_atan),
                   min_vertices, max_vertices)


class Texture(C.Structure):
    _fields_ = [("format", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("mip_count", C.c_uint32), ("texels", C.c_void_p)]


TEXEL = dict(rgba32f=0, rgba8_unorm=1, rgba8_srgb=2)


class RisltcError(RuntimeError):
    pass


def lib():
    """Load librisltc_cuda.so (built in-tree by risltc_b200.build). Never falls back to anything else."""
    global _lib
    if _lib is None:
        path = PKG / "librisltc_cuda.so"
        if not path.exists():
            raise RisltcError(f"{path} is missing: run `python -m risltc_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
        _lib = C.CDLL(str(path))
        _lib.risltc_cuda_last_error.restype = C.c_char_p
        _lib.risltc_cuda_last_frame_ms.restype = C.c_float
        _lib.risltc_cuda_owned_rows.restype = C.c_uint32
        _lib.risltc_cuda_frame_overlap_active.restype = C.c_uint32
        _lib.risltc_cuda_stream.restype = C.c_void_p
    return _lib


def _check(rc):
    if rc != 0:
        raise RisltcError(lib().risltc_cuda_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def check_bvh(vertices, max_leaf=2):
    """Host-only check of the BVH builders (binary tree and its 4-wide, 8-bit collapse); vertices [T, 3, 3] float32."""
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 9)
    report = (C.c_uint64 * 6)()
    _check(lib().risltc_cuda_check_bvh(_p(v), C.c_uint64(v.shape[0]), C.c_uint32(max_leaf), report))
    r = [int(x) for x in report]
    return dict(bad_order=r[0], bad_binary=r[1], bad_wide=r[2], binary_nodes=r[3], wide_nodes=r[4], depth=r[5] >> 32, children_per_node=(r[5] & 0xFFFFFFFF) / 100.0)


def bvh_checksum(vertices, max_leaf=2):
    """Host-only hash of the host builder's output for these triangles ([T, 3, 3] float32)."""
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 9)
    out = C.c_uint64()
    _check(lib().risltc_cuda_bvh_checksum(_p(v), C.c_uint64(v.shape[0]), C.c_uint32(max_leaf), C.byref(out)))
    return int(out.value)


class Device:
    """One risltc_device_t: one CUDA device, one stream, one image stripe set."""

    def __init__(self, ordinal=0):
        self.h = C.c_void_p()
        _check(lib().risltc_cuda_create_device(C.byref(self.h), C.c_int(ordinal)))
        self.width = self.height = 0

    def close(self):
        if self.h:
            lib().risltc_cuda_destroy_device(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads
    def upload_mesh(self, mesh):
        pos = np.ascontiguousarray(mesh["positions"], dtype=np.uint32)
        nuv = np.ascontiguousarray(mesh["normals_uv"], dtype=np.uint16)
        mat = np.ascontiguousarray(mesh["material_indices"], dtype=np.uint8)
        f = (C.c_float * 3)(*[float(x) for x in mesh["dequant_factor"]])
        s = (C.c_float * 3)(*[float(x) for x in mesh["dequant_summand"]])
        _check(lib().risltc_cuda_upload_scene(self.h, _p(pos), _p(nuv), _p(mat), C.c_uint64(mat.shape[0]), f, s))

    def upload_materials(self, constants):
        m = np.ascontiguousarray(constants, dtype=np.float32)
        _check(lib().risltc_cuda_upload_materials(self.h, _p(m), C.c_uint64(m.shape[0])))

    def upload_textures(self, textures):
        """textures: 3 per material, dicts {format: 'rgba32f' | 'rgba8_unorm' | 'rgba8_srgb', levels: [(h, w, 4) arrays, largest first]}."""
        arr = (Texture * len(textures))()
        keep = []
        for i, tex in enumerate(textures):
            dtype = np.float32 if tex["format"] == "rgba32f" else np.uint8
            buf = np.concatenate([np.ascontiguousarray(l, dtype=dtype).reshape(-1) for l in tex["levels"]])
            keep.append(buf)
            arr[i].format = TEXEL[tex["format"]]
            arr[i].height, arr[i].width = tex["levels"][0].shape[:2]
            arr[i].mip_count = len(tex["levels"])
            arr[i].texels = buf.ctypes.data
        _check(lib().risltc_cuda_upload_textures(self.h, arr, C.c_uint64(len(textures))))

    def upload_lights(self, records):
        r = np.ascontiguousarray(records, dtype=np.float32)
        _check(lib().risltc_cuda_upload_lights(self.h, _p(r), C.c_uint32(r.shape[0]), C.c_uint32((r.shape[1] - 12) // 4)))

    def upload_ltc(self, rgba16, rg16):
        a = np.ascontiguousarray(rgba16, dtype=np.uint16); b = np.ascontiguousarray(rg16, dtype=np.uint16)
        _check(lib().risltc_cuda_upload_ltc(self.h, _p(a), _p(b), C.c_uint32(a.shape[2]), C.c_uint32(a.shape[1]), C.c_uint32(a.shape[0])))

    def set_variant(self, var):
        _check(lib().risltc_cuda_set_variant(self.h, C.byref(var)))

    def set_precision(self, mode):
        _check(lib().risltc_cuda_set_precision(self.h, C.c_uint32({"fast": 0, "exact": 1, "hybrid": 2}[mode])))

    def set_kernels(self, gbuffer="auto", shadow="pairs"):
        _check(lib().risltc_cuda_set_kernels(self.h, C.c_uint32({"bvh": 0, "raster": 1, "auto": 2}[gbuffer]), C.c_uint32({"binary": 2, "wide": 4, "pairs": 8}[shadow])))

    def frame_overlap_active(self):
        return bool(lib().risltc_cuda_frame_overlap_active(self.h))

    def set_bvh_builder(self, builder="auto"):
        """Builder of the acceleration structures for the next upload_mesh: host (binned SAH), device (Morton-order radix tree) or auto."""
        _check(lib().risltc_cuda_set_bvh_builder(self.h, C.c_uint32({"host": 0, "device": 1, "auto": 2, "radix": 3}[builder])))

    def bvh_stats(self):
        out = (C.c_double * 8)()
        _check(lib().risltc_cuda_bvh_stats(self.h, out))
        v = [float(x) for x in out]
        return dict(builder={0: "host", 1: "device", 3: "radix"}[int(v[0])], build_ms=v[1], device_ms=v[2:5], binary_node_slots=int(v[5]), wide_nodes=int(v[6]),
                    binary_depth=int(v[7]) >> 16, wide_depth=int(v[7]) & 0xFFFF)

    def check_scene_bvh(self):
        """Invariants of the acceleration structures on the device (see check_bvh); the first three must be 0."""
        report = (C.c_uint64 * 6)()
        _check(lib().risltc_cuda_check_scene_bvh(self.h, report))
        r = [int(x) for x in report]
        return dict(bad_order=r[0], bad_binary=r[1], bad_wide=r[2], binary_nodes=r[3], wide_nodes=r[4], depth=r[5] >> 32, children_per_node=(r[5] & 0xFFFFFFFF) / 100.0)

    def set_frame_overlap(self, mode="auto"):
        _check(lib().risltc_cuda_set_frame_overlap(self.h, C.c_uint32({"off": 0, "on": 1, "auto": 2}[mode])))

    def resize(self, width, height, stripe_height=8, stripe_index=0, stripe_count=1):
        _check(lib().risltc_cuda_resize(self.h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(stripe_height),
                                        C.c_uint32(stripe_index), C.c_uint32(stripe_count)))
        self.width, self.height = width, height

    def set_accum_buffer(self, device_pointer):
        _check(lib().risltc_cuda_set_accum_buffer(self.h, C.c_void_p(device_pointer)))

    @property
    def owned_rows(self):
        return int(lib().risltc_cuda_owned_rows(self.h))

    def owned_row_indices(self):
        rows = np.zeros(self.owned_rows, dtype=np.uint32)
        _check(lib().risltc_cuda_owned_row_indices(self.h, _p(rows)))
        return rows

    # ---- frames
    def render_frames(self, constants_blocks, first_accum_num=0):
        """constants_blocks: bytes-like / uint8 array of n * 256 bytes (per_frame_constants_t blocks)."""
        buf = np.frombuffer(bytes(constants_blocks), dtype=np.uint8) if not isinstance(constants_blocks, np.ndarray) else constants_blocks
        buf = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
        assert buf.size % 256 == 0
        _check(lib().risltc_cuda_render_frames(self.h, _p(buf), C.c_uint32(buf.size // 256), C.c_uint32(first_accum_num)))

    def synchronize(self):
        _check(lib().risltc_cuda_synchronize(self.h))

    def read_accum(self, out=None):
        if out is None:
            out = np.empty((self.owned_rows, self.width, 4), dtype=np.float32)
        _check(lib().risltc_cuda_read_accum(self.h, _p(out)))
        return out

    def read_visibility(self):
        out = np.empty((self.owned_rows, self.width), dtype=np.uint32)
        _check(lib().risltc_cuda_read_visibility(self.h, _p(out)))
        return out

    def copy_pass(self, frame_bits=0):
        """(owned_rows, width, 3) uint8: the displayed sRGB image (0) or the low (1) / high (2) bytes of the half-float bits."""
        out = np.empty((self.owned_rows, self.width, 3), dtype=np.uint8)
        _check(lib().risltc_cuda_copy_pass(self.h, C.c_uint32(frame_bits), _p(out)))
        return out

    def last_frame_ms(self):
        return float(lib().risltc_cuda_last_frame_ms(self.h))

    def last_kernel_ms(self):
        ms = (C.c_float * 4)()
        _check(lib().risltc_cuda_last_kernel_ms(self.h, ms))
        return list(ms)

    def last_pass_ms(self):
        """[visibility, RIS, winner, shadow rays, accumulation, whole call] in ms, summed over the frames of the last call."""
        ms = (C.c_float * 6)()
        _check(lib().risltc_cuda_last_pass_ms(self.h, ms))
        return list(ms)

    def traversal_counters(self, enable=True):
        """Counters of the last counted call (rays, node visits, triangle tests, occluded rays); switches counting on / off."""
        c = (C.c_uint64 * 4)()
        _check(lib().risltc_cuda_traversal_counters(self.h, C.c_uint32(int(enable)), c))
        return dict(rays=int(c[0]), node_visits=int(c[1]), triangle_tests=int(c[2]), occluded=int(c[3]))

    def counters(self):
        c = (C.c_uint64 * 4)()
        _check(lib().risltc_cuda_counters(self.h, c))
        return dict(shaded_pixels=int(c[0]), shadow_rays=int(c[1]), launches=int(c[2]), candidates=int(c[3]))

    # ---- known-answer entry points
    def kat_clip(self, polygons, counts, max_light_vertices):
        p = np.ascontiguousarray(polygons, dtype=np.float32).copy(); c = np.ascontiguousarray(counts, dtype=np.uint32).copy()
        _check(lib().risltc_cuda_kat_clip(self.h, _p(p), _p(c), C.c_uint32(c.shape[0]), C.c_uint32(max_light_vertices)))
        return p, c

    def kat_ltc_integral(self, polygons, counts):
        p = np.ascontiguousarray(polygons, dtype=np.float32); c = np.ascontiguousarray(counts, dtype=np.uint32)
        out = np.empty(c.shape[0], dtype=np.float32)
        _check(lib().risltc_cuda_kat_ltc_integral(self.h, _p(p), _p(c), _p(out), C.c_uint32(c.shape[0])))
        return out

    def kat_psa(self, polygons, counts, randoms, max_polygon_vertices, fast_atan=0, biased=0):
        p = np.ascontiguousarray(polygons, dtype=np.float32); c = np.ascontiguousarray(counts, dtype=np.uint32)
        r = np.ascontiguousarray(randoms, dtype=np.float32)
        op = np.empty((c.shape[0], 44), dtype=np.float32); od = np.empty((c.shape[0], 3), dtype=np.float32)
        _check(lib().risltc_cuda_kat_psa(self.h, _p(p), _p(c), _p(r), _p(op), _p(od), C.c_uint32(c.shape[0]),
                                         C.c_uint32(max_polygon_vertices), C.c_uint32(fast_atan), C.c_uint32(biased)))
        return op, od

    def kat_noise(self, width, height, frame_word, draws):
        out = np.empty((height, width, draws), dtype=np.float32)
        _check(lib().risltc_cuda_kat_noise(self.h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(frame_word), C.c_uint32(draws), _p(out)))
        return out

    def kat_ltc_coefficients(self, inputs, ltc_constants):
        a = np.ascontiguousarray(inputs, dtype=np.float32)
        out = np.empty((a.shape[0], 33), dtype=np.float32)
        _check(lib().risltc_cuda_kat_ltc_coefficients(self.h, _p(a), (C.c_float * 6)(*ltc_constants), _p(out), C.c_uint32(a.shape[0])))
        return out

    def kat_any_hit(self, rays):
        r = np.ascontiguousarray(rays, dtype=np.float32)
        hits = np.empty(r.shape[0], dtype=np.uint32)
        _check(lib().risltc_cuda_kat_any_hit(self.h, _p(r), _p(hits), C.c_uint32(r.shape[0])))
        return hits

    def kat_exact_math(self):
        out = (C.c_uint64 * 3)()
        _check(lib().risltc_cuda_kat_exact_math(self.h, out))
        return [int(v) for v in out]

    def kat_trace(self, rays, kind=4):
        """The frame path's shadow-ray kernels on an array of rays (8: 4-wide quantised tree, rays i and i + n / 2 as a pair
        with a common origin; 4: the same tree, one ray per lane; 2: binary tree)."""
        r = np.ascontiguousarray(rays, dtype=np.float32)
        hits = np.empty(r.shape[0], dtype=np.uint32)
        _check(lib().risltc_cuda_kat_trace(self.h, _p(r), _p(hits), C.c_uint32(r.shape[0]), C.c_uint32(kind)))
        return hits
