"""Readers / writers for the reference's on-disk formats (host-side tooling).

All layouts follow the reference byte for byte so that files written here load in
the reference and vice versa:
  * ``.vks`` scenes   -- scene.c:409-483 (reader), tools/io_export_vulkan_blender28.py:481-541 (writer)
  * ``.save`` quicksaves -- main.c:45-125, polygonal_light.h:73-115
  * ``.vkt`` textures -- textures.c:95-172, tools/texture_conversion/main.c:41-63
  * ``fit<i>.dat`` LTC fits -- ltc_table.c:46-47,82-84
"""
import struct
from pathlib import Path

import numpy as np

VKS_MARKER, VKS_VERSION, EOF_MARKER = 0xABCABC, 1, 0xE0FE0F
VKT_MARKER = 0xBC1BC1
VK_FORMAT_R32G32B32A32_SFLOAT = 109
QUICKSAVE_LIGHT_BYTES = 4 * 20 + 4 * 2  # POLYGONAL_LIGHT_QUICKSAVE_SIZE


# --------------------------------------------------------------------- .vks
def write_vks(path, scene):
    """scene: dict with material_names, positions (T*3,2) u32, normals_uv (T*3,4) u16,
    material_indices (T,) u8, dequant_factor (3,), dequant_summand (3,)."""
    pos = np.ascontiguousarray(scene["positions"], dtype=np.uint32)
    nuv = np.ascontiguousarray(scene["normals_uv"], dtype=np.uint16)
    mat = np.ascontiguousarray(scene["material_indices"], dtype=np.uint8)
    T = mat.shape[0]
    assert pos.shape == (T * 3, 2) and nuv.shape == (T * 3, 4)
    with open(path, "wb") as f:
        f.write(struct.pack("<II", VKS_MARKER, VKS_VERSION))
        f.write(struct.pack("<QQ", len(scene["material_names"]), T))
        f.write(np.asarray(scene["dequant_factor"], dtype="<f4").tobytes())
        f.write(np.asarray(scene["dequant_summand"], dtype="<f4").tobytes())
        for name in scene["material_names"]:
            raw = name.encode("utf-8")
            f.write(struct.pack("<Q", len(raw)) + raw + b"\0")
        f.write(pos.tobytes())
        f.write(nuv.tobytes())
        f.write(mat.tobytes())
        f.write(struct.pack("<I", EOF_MARKER))


def read_vks(path):
    data = Path(path).read_bytes()
    marker, version, n_mat, T = struct.unpack_from("<IIQQ", data, 0)
    if marker != VKS_MARKER or version != VKS_VERSION:
        raise ValueError(f"{path}: bad marker 0x{marker:x} / version {version}")
    if T == 0:
        raise ValueError(f"{path}: holds 0 triangles")
    off = 24
    factor = np.frombuffer(data, "<f4", 3, off); off += 12
    summand = np.frombuffer(data, "<f4", 3, off); off += 12
    names = []
    for _ in range(n_mat):
        (length,) = struct.unpack_from("<Q", data, off); off += 8
        names.append(data[off:off + length].decode("utf-8")); off += length + 1
    pos = np.frombuffer(data, "<u4", T * 6, off).reshape(T * 3, 2); off += T * 24
    nuv = np.frombuffer(data, "<u2", T * 12, off).reshape(T * 3, 4); off += T * 24
    mat = np.frombuffer(data, "u1", T, off); off += T
    (eof,) = struct.unpack_from("<I", data, off)
    if eof != EOF_MARKER:
        raise ValueError(f"{path}: geometry is not followed by the end-of-file marker")
    return dict(material_names=names, positions=pos.copy(), normals_uv=nuv.copy(), material_indices=mat.copy(),
                dequant_factor=factor.copy(), dequant_summand=summand.copy())


# -------------------------------------------------------------------- .save
def write_quicksave(path, camera, lights):
    """camera: dict with the 12 first_person_camera_t fields (camera.h:29-49).
    lights: list of dicts (rotation_angles, scaling_x, scaling_y, translation, radiant_flux,
    vertices_plane_space (n,2), texture_file_path optional)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<3f6fi2f", *camera["position"], camera["rotation_z"], camera["rotation_x"],
                            camera["vertical_fov"], camera["near"], camera["far"], camera.get("speed", 2.0),
                            int(camera.get("rotate_camera", 0)), camera.get("rotation_x_0", 0.0), camera.get("rotation_z_0", 0.0)))
        f.write(struct.pack("<II", 0, len(lights)))
        for light in lights:
            verts = np.asarray(light["vertices_plane_space"], dtype=np.float32)
            n = verts.shape[0]
            head = struct.pack("<3f f 3f f 3f f 3f f 4f I I",
                               *light["rotation_angles"], light["scaling_x"], *light["translation"], light["scaling_y"],
                               *light["radiant_flux"], 1.0 / light["scaling_x"], 0.0, 0.0, 0.0, 1.0 / light["scaling_y"],
                               0.0, 0.0, 0.0, 0.0, n, 0)
            assert len(head) == QUICKSAVE_LIGHT_BYTES
            f.write(head)
            tex = light.get("texture_file_path")
            if tex:
                raw = tex.encode("utf-8") + b"\0"
                f.write(struct.pack("<Q", len(raw)) + raw)
            else:
                f.write(struct.pack("<Q", 0))
            f.write(struct.pack("<QQ", 0, 0))
            padded = np.zeros((n, 4), dtype="<f4")
            padded[:, :2] = verts[:, :2]
            f.write(padded.tobytes())


def read_quicksave(path):
    data = Path(path).read_bytes()
    cam = struct.unpack_from("<3f6fi2f", data, 0)
    camera = dict(position=list(cam[0:3]), rotation_z=cam[3], rotation_x=cam[4], vertical_fov=cam[5], near=cam[6],
                  far=cam[7], speed=cam[8], rotate_camera=cam[9], rotation_x_0=cam[10], rotation_z_0=cam[11])
    off = 48
    _, count = struct.unpack_from("<II", data, off); off += 8
    lights = []
    for _ in range(count):
        h = struct.unpack_from("<3f f 3f f 3f f 3f f 4f I I", data, off); off += QUICKSAVE_LIGHT_BYTES
        (path_size,) = struct.unpack_from("<Q", data, off); off += 8
        tex = None
        if path_size:
            tex = data[off:off + path_size - 1].decode("utf-8"); off += path_size
        off += 16
        n = h[20]
        verts = np.frombuffer(data, "<f4", 4 * n, off).reshape(n, 4)[:, :2].copy(); off += 16 * n
        scaling_y = h[7] if h[7] > 0.0 else h[3]  # legacy fix, main.c:104
        lights.append(dict(rotation_angles=list(h[0:3]), scaling_x=h[3], translation=list(h[4:7]), scaling_y=scaling_y,
                           radiant_flux=list(h[8:11]), vertices_plane_space=verts, texture_file_path=tex))
    return camera, lights


# --------------------------------------------------------------------- .vkt
def write_vkt_rgba32f(path, image):
    """Single-mip RGBA32F texture (VK_FORMAT 109); image is (h, w, 4) float32."""
    img = np.ascontiguousarray(image, dtype="<f4")
    h, w, _ = img.shape
    payload = img.tobytes()
    with open(path, "wb") as f:
        f.write(struct.pack("<IIIIIIQ", VKT_MARKER, 1, 1, w, h, VK_FORMAT_R32G32B32A32_SFLOAT, len(payload)))
        f.write(struct.pack("<IIQQ", w, h, len(payload), 0))
        f.write(payload)
        f.write(struct.pack("<I", EOF_MARKER))


def read_vkt(path):
    data = Path(path).read_bytes()
    marker, version, mips, w, h, fmt, size = struct.unpack_from("<IIIIIIQ", data, 0)
    if marker != VKT_MARKER or version != 1:
        raise ValueError(f"{path}: not a .vkt texture")
    off = 32
    headers = []
    for _ in range(mips):
        headers.append(struct.unpack_from("<IIQQ", data, off)); off += 24
    payload = data[off:off + size]
    (eof,) = struct.unpack_from("<I", data, off + size)
    if eof != EOF_MARKER:
        raise ValueError(f"{path}: texture data is not followed by the end-of-file marker")
    if fmt != VK_FORMAT_R32G32B32A32_SFLOAT:
        raise NotImplementedError(f"{path}: VkFormat {fmt} (block-compressed textures are not decoded yet)")
    out = []
    for (mw, mh, msize, moff) in headers:
        out.append(np.frombuffer(payload, "<f4", mw * mh * 4, moff).reshape(mh, mw, 4).copy())
    return out


# ----------------------------------------------------------------- fit*.dat
def write_ltc_fits(directory, fits):
    """fits: (layers, res, res, 5) float32 -> directory/fit<i>.dat"""
    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    fits = np.asarray(fits, dtype="<f4")
    for i in range(fits.shape[0]):
        with open(directory / f"fit{i}.dat", "wb") as f:
            f.write(struct.pack("<Q", fits.shape[1]))
            f.write(fits[i].tobytes())


def read_ltc_fits(directory, fresnel_count):
    out = []
    for i in range(fresnel_count):
        data = (Path(directory) / f"fit{i}.dat").read_bytes()
        (res,) = struct.unpack_from("<Q", data, 0)
        out.append(np.frombuffer(data, "<f4", res * res * 5, 8).reshape(res, res, 5))
    return np.stack(out)


# --------------------------------------------------------------------- .hdr / .png
def read_hdr(path):
    """Radiance RGBE, flat or adaptive run-length scanlines (what host/screenshot.c and stb_image_write store);
    returns (H, W, 3) float32."""
    data = Path(path).read_bytes()
    end = data.index(b"\n\n") + 2
    line_end = data.index(b"\n", end)
    parts = data[end:line_end].split()
    assert parts[0] == b"-Y" and parts[2] == b"+X", parts
    H, W = int(parts[1]), int(parts[3])
    raw = np.frombuffer(data, np.uint8, offset=line_end + 1)
    rgbe = np.zeros((H, W, 4), dtype=np.uint8)
    pos = 0
    for y in range(H):
        if W < 8 or W >= 32768 or not (raw[pos] == 2 and raw[pos + 1] == 2 and (int(raw[pos + 2]) << 8 | int(raw[pos + 3])) == W):
            rgbe[y] = raw[pos:pos + 4 * W].reshape(W, 4); pos += 4 * W
            continue
        pos += 4
        for c in range(4):
            x = 0
            while x < W:
                n = int(raw[pos]); pos += 1
                if n > 128:
                    rgbe[y, x:x + n - 128, c] = raw[pos]; pos += 1; x += n - 128
                else:
                    rgbe[y, x:x + n, c] = raw[pos:pos + n]; pos += n; x += n
            assert x == W
    scale = np.where(rgbe[..., 3:4] == 0, 0.0, np.ldexp(1.0, rgbe[..., 3:4].astype(np.int32) - 136))
    return (rgbe[..., :3].astype(np.float64) * scale).astype(np.float32)


def read_png(path):
    """8-bit truecolour PNG without interlacing (what host/screenshot.c stores); returns (H, W, 3) uint8. Checks every CRC."""
    import zlib
    data = Path(path).read_bytes()
    assert data[:8] == bytes([137, 80, 78, 71, 13, 10, 26, 10])
    pos, idat, W, H = 8, b"", 0, 0
    while pos < len(data):
        (length,) = struct.unpack_from(">I", data, pos)
        kind, body = data[pos + 4:pos + 8], data[pos + 8:pos + 8 + length]
        (crc,) = struct.unpack_from(">I", data, pos + 8 + length)
        assert zlib.crc32(kind + body) == crc, kind
        if kind == b"IHDR":
            W, H, depth, colour, compression, flt, interlace = struct.unpack(">IIBBBBB", body)
            assert (depth, colour, compression, flt, interlace) == (8, 2, 0, 0, 0)
        elif kind == b"IDAT":
            idat += body
        pos += 12 + length
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(H, 3 * W + 1)
    out = np.zeros((H, 3 * W), dtype=np.uint8)
    for y in range(H):   # undo the scanline filters (types 0-2 suffice for files written here and by simple encoders)
        f, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if f == 0:
            out[y] = line
        elif f == 1:
            for x in range(3 * W):
                out[y, x] = (line[x] + (int(out[y, x - 3]) if x >= 3 else 0)) & 255
        elif f == 2:
            out[y] = (line + (out[y - 1].astype(np.int32) if y else 0)) & 255
        else:
            raise ValueError(f"PNG filter {f} not supported")
    return out.reshape(H, W, 3)
