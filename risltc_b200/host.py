"""ctypes front end of the C99 host layer (librisltc_host.so): the reference-shaped path
scene files -> load_scene / quick_load / load_ltc_table -> write_constants / write_lights ->
C ABI -> kernels. Also writes generated scenes to disk in the reference's formats."""
import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import formats

PKG = Path(__file__).resolve().parent
_lib = None


class RenderSettings(C.Structure):
    """render_settings_t, main.h:123-156"""
    _fields_ = [("exposure_factor", C.c_float), ("roughness_factor", C.c_float), ("sample_count", C.c_uint32),
                ("sample_count_light", C.c_uint32), ("mis_heuristic", C.c_int), ("light_sampling", C.c_int),
                ("mis_visibility_estimate", C.c_float), ("polygon_sampling_technique", C.c_int), ("error_display", C.c_int),
                ("error_min_exponent", C.c_float), ("animate_noise", C.c_uint32), ("accum", C.c_uint32),
                ("show_polygonal_lights", C.c_uint32), ("show_gui", C.c_uint32), ("v_sync", C.c_uint32), ("fast_atan", C.c_uint32)]


def lib():
    global _lib
    if _lib is None:
        path = PKG / "librisltc_host.so"
        if not path.exists():
            raise RuntimeError(f"{path} is missing: run `python -m risltc_b200.build`")
        _lib = C.CDLL(str(path))
        _lib.risltc_app_create.restype = C.c_void_p
        _lib.risltc_app_device.restype = C.c_void_p
        _lib.risltc_app_wait.restype = C.c_float
        _lib.risltc_app_light_buffer_size.restype = C.c_size_t
        _lib.risltc_app_accum_num.restype = C.c_uint32
        _lib.wang_random_number.restype = C.c_uint32
    return _lib


def write_scene_files(scene, directory, ltc_fits=None, name=None):
    """Write a generated scene as <dir>/<name>.vks, <dir>/<name>_textures/*.vkt, <dir>/quicksaves/<name>.save
    and (optionally) <dir>/ggx_ltc_fit/fit<i>.dat. Returns the three paths load_scene / quick_load need."""
    directory = Path(directory)
    name = name or scene["name"]
    tex = directory / f"{name}_textures"
    (directory / "quicksaves").mkdir(parents=True, exist_ok=True)
    tex.mkdir(parents=True, exist_ok=True)
    vks = directory / f"{name}.vks"
    formats.write_vks(vks, scene["mesh"])
    if scene.get("textures") is not None:
        # the formats the reference's texture conversion tool produces: BC1 sRGB base colour, BC1 specular, BC5 normal
        for i, m in enumerate(scene["materials"]):
            base, spec, nrm = scene["textures"][3 * i:3 * i + 3]
            formats.write_vkt(tex / f"{m['name']}_BaseColor.vkt", base["levels"], formats.VK_FORMAT_BC1_RGB_SRGB_BLOCK if base["format"] == "rgba8_srgb" else formats.VK_FORMAT_BC1_RGB_UNORM_BLOCK)
            formats.write_vkt(tex / f"{m['name']}_Specular.vkt", spec["levels"], formats.VK_FORMAT_BC1_RGB_UNORM_BLOCK)
            formats.write_vkt(tex / f"{m['name']}_Normal.vkt", nrm["levels"], formats.VK_FORMAT_BC5_UNORM_BLOCK)
    for m in (scene["materials"] if scene.get("textures") is None else []):
        base = np.array([[list(m["base_color"]) + [1.0]]], dtype=np.float32)
        spec = np.array([[[1.0, m["roughness"], m["metalicity"], 1.0]]], dtype=np.float32)
        nrm = np.array([[[0.5, 0.5, 1.0, 1.0]]], dtype=np.float32)
        formats.write_vkt_rgba32f(tex / f"{m['name']}_BaseColor.vkt", base)
        formats.write_vkt_rgba32f(tex / f"{m['name']}_Specular.vkt", spec)
        formats.write_vkt_rgba32f(tex / f"{m['name']}_Normal.vkt", nrm)
    save = directory / "quicksaves" / f"{name}.save"
    formats.write_quicksave(save, scene["camera"], scene["lights"])
    if ltc_fits is not None:
        formats.write_ltc_fits(directory / "ggx_ltc_fit", ltc_fits)
    return str(vks), str(tex), str(save)


class Application:
    """application_t driven through the embedding entry points (host/embedding.c)."""

    def __init__(self, data_dir, ordinal=0, stripe_height=8, stripe_index=0, stripe_count=1):
        os.environ["RISLTC_DATA_DIR"] = str(data_dir)
        self.h = lib().risltc_app_create(C.c_int(ordinal), C.c_uint32(stripe_height), C.c_uint32(stripe_index), C.c_uint32(stripe_count))
        if not self.h:
            raise RuntimeError("startup_application failed (see the message printed above)")
        self.h = C.c_void_p(self.h)
        self.width = self.height = 0

    def close(self):
        if self.h:
            lib().risltc_app_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, vks, texture_dir, quick_save, width, height):
        if lib().risltc_app_load(self.h, vks.encode(), texture_dir.encode(), quick_save.encode(), C.c_uint32(width), C.c_uint32(height)):
            raise RuntimeError("risltc_app_load failed (see the message printed above)")
        self.width, self.height = width, height

    def settings(self, **changes):
        s = RenderSettings()
        lib().risltc_app_get_render_settings(self.h, C.byref(s))
        for k, v in changes.items():
            setattr(s, k, v)
        if lib().risltc_app_set_render_settings(self.h, C.byref(s)):
            raise RuntimeError("risltc_app_set_render_settings failed")
        return s

    def reset(self, random_seed=0):
        lib().risltc_app_reset(self.h, C.c_uint32(random_seed))

    def render_frames(self, count, upload_lights=False):
        if lib().risltc_app_render_frames(self.h, C.c_uint32(count), C.c_int(int(upload_lights))):
            raise RuntimeError("risltc_app_render_frames failed")

    def screenshot(self, png=None, hdr=None):
        """implement_screenshot (main.c:2358-2409) of the accumulated frame: 8-bit sRGB *.png and / or RGBE *.hdr."""
        if lib().risltc_app_screenshot(self.h, str(png).encode() if png else None, str(hdr).encode() if hdr else None):
            raise RuntimeError("taking the screenshot failed (see the message printed above)")

    def wait_ms(self):
        return float(lib().risltc_app_wait(self.h))

    def device(self):
        """The risltc_device_t as an api.Device that does not own the handle."""
        from . import api
        dev = api.Device.__new__(api.Device)
        dev.h = C.c_void_p(lib().risltc_app_device(self.h))
        dev.width, dev.height = self.width, self.height
        dev.close = lambda: None
        return dev

    def write_constants(self):
        buf = (C.c_ubyte * 256)()
        lib().write_constants(buf, self.h)
        return bytes(buf)

    def write_lights(self):
        size = int(lib().risltc_app_light_buffer_size(self.h))
        buf = (C.c_ubyte * size)()
        lib().write_lights(buf, self.h)
        return bytes(buf)
