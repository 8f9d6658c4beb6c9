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


VK_FORMAT_BC1_RGB_UNORM_BLOCK, VK_FORMAT_BC1_RGB_SRGB_BLOCK, VK_FORMAT_BC5_UNORM_BLOCK = 131, 132, 141


def _blocks(level):
    """(H, W, C) uint8 -> (blocks_y, blocks_x, 16, C) with edge texels repeated into partial blocks."""
    h, w, c = level.shape
    H, W = (h + 3) // 4 * 4, (w + 3) // 4 * 4
    padded = np.zeros((H, W, c), dtype=np.uint8)
    padded[:h, :w] = level
    padded[h:, :w] = level[h - 1:h, :]
    padded[:, w:] = padded[:, w - 1:w]
    return padded.reshape(H // 4, 4, W // 4, 4, c).transpose(0, 2, 1, 3, 4).reshape(H // 4, W // 4, 16, c)


def bc1_palette(c0, c1):
    """(..., 4, 3) uint8 palette of BC1_RGB blocks with endpoints c0, c1 (uint16 RGB565): the decode rule of the host loader
    (host/tables_scene.c): endpoints by bit replication, interpolated entries rounded to the nearest 8-bit value."""
    def expand(c):
        r, g, b = (c >> 11) & 31, (c >> 5) & 63, c & 31
        return np.stack([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)], axis=-1).astype(np.uint32)
    e0, e1 = expand(c0.astype(np.uint32)), expand(c1.astype(np.uint32))
    four = (c0 > c1)[..., None]
    p2 = np.where(four, (2 * e0 + e1 + 1) // 3, (e0 + e1 + 1) // 2)
    p3 = np.where(four, (e0 + 2 * e1 + 1) // 3, 0)
    return np.stack([e0, e1, p2, p3], axis=-2).astype(np.uint8)


def encode_bc1(level):
    """(H, W, >=3) uint8 -> BC1_RGB blocks (bytes). Endpoints = the block's bounding box corners in RGB565, every texel takes
    the nearest palette entry: simple, deterministic, good enough for test assets."""
    b = _blocks(level[..., :3]).astype(np.int32)
    lo, hi = b.min(axis=2), b.max(axis=2)

    def pack565(c):
        return (((c[..., 0] >> 3) << 11) | ((c[..., 1] >> 2) << 5) | (c[..., 2] >> 3)).astype(np.uint16)
    c0, c1 = pack565(hi), pack565(lo)
    swap = c0 < c1
    c0, c1 = np.where(swap, c1, c0), np.where(swap, c0, c1)       # c0 >= c1; equal endpoints select the 3-colour mode, whose entries 0 and 1 coincide
    pal = bc1_palette(c0, c1).astype(np.int32)                    # (by, bx, 4, 3)
    usable = np.where((c0 > c1)[..., None], 4, 2)                 # in 3-colour mode avoid the black entry
    d = ((b[:, :, :, None, :] - pal[:, :, None, :, :]) ** 2).sum(axis=-1)       # (by, bx, 16, 4)
    d = np.where(np.arange(4)[None, None, None, :] < usable[..., None], d, 1 << 30)
    idx = d.argmin(axis=-1).astype(np.uint32)                     # (by, bx, 16)
    bits = (idx << (2 * np.arange(16, dtype=np.uint32))).sum(axis=-1).astype(np.uint32)
    out = np.zeros(c0.shape + (8,), dtype=np.uint8)
    out[..., 0] = c0 & 255; out[..., 1] = c0 >> 8; out[..., 2] = c1 & 255; out[..., 3] = c1 >> 8
    for k in range(4):
        out[..., 4 + k] = (bits >> (8 * k)) & 255
    return out.tobytes()


def bc4_palette(e0, e1):
    e0, e1 = e0.astype(np.uint32), e1.astype(np.uint32)
    eight = e0 > e1
    pal = [e0, e1]
    for k in range(1, 7):
        a = ((7 - k) * e0 + k * e1 + 3) // 7
        if k <= 4:
            b = ((5 - k) * e0 + k * e1 + 2) // 5
        else:
            b = np.zeros_like(e0) if k == 5 else np.full_like(e0, 255)
        pal.append(np.where(eight, a, b))
    return np.stack(pal, axis=-1).astype(np.uint8)


def _encode_bc4(channel_blocks):
    b = channel_blocks.astype(np.int32)                           # (by, bx, 16)
    e0, e1 = b.max(axis=2).astype(np.uint8), b.min(axis=2).astype(np.uint8)     # e0 >= e1; equal: 6-value mode, entries 0 / 1 coincide
    pal = bc4_palette(e0, e1).astype(np.int32)
    usable = np.where((e0 > e1)[..., None], 8, 2)
    d = np.abs(b[..., None] - pal[:, :, None, :])
    d = np.where(np.arange(8)[None, None, None, :] < usable[..., None], d, 1 << 30)
    idx = d.argmin(axis=-1).astype(np.uint64)
    bits = (idx << (3 * np.arange(16, dtype=np.uint64))).sum(axis=-1).astype(np.uint64)
    out = np.zeros(e0.shape + (8,), dtype=np.uint8)
    out[..., 0] = e0; out[..., 1] = e1
    for k in range(6):
        out[..., 2 + k] = (bits >> np.uint64(8 * k)) & np.uint64(255)
    return out


def encode_bc5(level):
    """(H, W, >=2) uint8 -> BC5 blocks (red BC4 block, green BC4 block)."""
    b = _blocks(level[..., :2])
    return np.concatenate([_encode_bc4(b[..., 0]), _encode_bc4(b[..., 1])], axis=-1).tobytes()


def decode_bc1(data, w, h):
    """numpy decoder of BC1_RGB blocks -> (h, w, 4) uint8; the independent check of the host loader's decoder."""
    by, bx = (h + 3) // 4, (w + 3) // 4
    raw = np.frombuffer(data, np.uint8, by * bx * 8).reshape(by, bx, 8).astype(np.uint32)
    c0, c1 = raw[..., 0] | (raw[..., 1] << 8), raw[..., 2] | (raw[..., 3] << 8)
    pal = bc1_palette(c0, c1)
    bits = raw[..., 4] | (raw[..., 5] << 8) | (raw[..., 6] << 16) | (raw[..., 7] << 24)
    idx = (bits[..., None] >> (2 * np.arange(16, dtype=np.uint32))) & 3
    tex = np.take_along_axis(pal, idx[..., None].astype(np.int64).repeat(3, axis=-1), axis=2)      # (by, bx, 16, 3)
    img = tex.reshape(by, bx, 4, 4, 3).transpose(0, 2, 1, 3, 4).reshape(by * 4, bx * 4, 3)[:h, :w]
    return np.concatenate([img, np.full((h, w, 1), 255, np.uint8)], axis=-1)


def _decode_bc4(raw):
    pal = bc4_palette(raw[..., 0].astype(np.uint8), raw[..., 1].astype(np.uint8))
    bits = np.zeros(raw.shape[:2], dtype=np.uint64)
    for k in range(6):
        bits |= raw[..., 2 + k].astype(np.uint64) << np.uint64(8 * k)
    idx = (bits[..., None] >> (3 * np.arange(16, dtype=np.uint64))) & np.uint64(7)
    return np.take_along_axis(pal, idx.astype(np.int64), axis=2)


def decode_bc5(data, w, h):
    by, bx = (h + 3) // 4, (w + 3) // 4
    raw = np.frombuffer(data, np.uint8, by * bx * 16).reshape(by, bx, 16)
    r, g = _decode_bc4(raw[..., :8]), _decode_bc4(raw[..., 8:])
    img = np.stack([r, g, np.zeros_like(r), np.full_like(r, 255)], axis=-1)
    return img.reshape(by, bx, 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(by * 4, bx * 4, 4)[:h, :w]


def write_vkt(path, levels, vk_format):
    """Mip-mapped texture in the layout of tools/texture_conversion/main.c:41-63. levels: (h, w, 4) arrays, largest first;
    vk_format 109 (RGBA32F, float32 levels), 131 / 132 (BC1 RGB UNORM / SRGB, uint8 levels) or 141 (BC5, uint8 levels)."""
    chunks = []
    for level in levels:
        if vk_format == VK_FORMAT_R32G32B32A32_SFLOAT:
            chunks.append(np.ascontiguousarray(level, dtype="<f4").tobytes())
        elif vk_format in (VK_FORMAT_BC1_RGB_UNORM_BLOCK, VK_FORMAT_BC1_RGB_SRGB_BLOCK):
            chunks.append(encode_bc1(np.asarray(level, dtype=np.uint8)))
        elif vk_format == VK_FORMAT_BC5_UNORM_BLOCK:
            chunks.append(encode_bc5(np.asarray(level, dtype=np.uint8)))
        else:
            raise ValueError(f"VkFormat {vk_format} is not written")
    payload = b"".join(chunks)
    h, w = levels[0].shape[:2]
    with open(path, "wb") as f:
        f.write(struct.pack("<IIIIIIQ", VKT_MARKER, 1, len(levels), w, h, vk_format, len(payload)))
        offset = 0
        for level, chunk in zip(levels, chunks):
            f.write(struct.pack("<IIQQ", level.shape[1], level.shape[0], len(chunk), offset))
            offset += len(chunk)
        f.write(payload)
        f.write(struct.pack("<I", EOF_MARKER))


def read_vkt(path, with_format=False):
    """All mip levels of a .vkt texture as (h, w, 4) arrays: float32 for RGBA32F, uint8 (decoded blocks) for BC1 / BC5."""
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
    out = []
    for (mw, mh, msize, moff) in headers:
        if fmt == VK_FORMAT_R32G32B32A32_SFLOAT:
            out.append(np.frombuffer(payload, "<f4", mw * mh * 4, moff).reshape(mh, mw, 4).copy())
        elif fmt in (VK_FORMAT_BC1_RGB_UNORM_BLOCK, VK_FORMAT_BC1_RGB_SRGB_BLOCK):
            out.append(decode_bc1(payload[moff:moff + msize], mw, mh))
        elif fmt == VK_FORMAT_BC5_UNORM_BLOCK:
            out.append(decode_bc5(payload[moff:moff + msize], mw, mh))
        else:
            raise NotImplementedError(f"{path}: VkFormat {fmt}")
    return (out, fmt) if with_format else out


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
