"""Procedural many-light scenes in the reference's data model.

The reference's scenes (Bistro, ZeroDay), quicksaves and LTC fits are website
downloads (README.md:8-11) that do not exist offline, so every benchmark and
parity input is generated here, deterministically from a seed, and can be
written in the reference's own file formats with risltc_b200.formats
(``.vks`` mesh, ``.save`` camera + lights, ``.vkt`` flat material textures,
``fit<i>.dat`` LTC tables).

A scene is a dict:
  mesh       -- material_names, positions (T*3,2) u32, normals_uv (T*3,4) u16,
                material_indices (T,) u8, dequant_factor, dequant_summand
                (exactly the .vks payload, tools/io_export_vulkan_blender28.py:481-541)
  materials  -- list of dicts base_color (linear rgb), roughness (linear), metalicity
  camera     -- first_person_camera_t fields (camera.h:29-49)
  lights     -- polygonal_light_t inputs (polygonal_light.h:73-99): rotation_angles,
                scaling_x/y, translation, radiant_flux, vertices_plane_space
"""
import numpy as np


# ------------------------------------------------------------------ mesh
class MeshBuilder:
    def __init__(self):
        self.tri, self.nrm, self.uv, self.mat, self.light = [], [], [], [], []

    def add_triangles(self, tris, material, light=False, normals=None):
        tris = np.asarray(tris, dtype=np.float64).reshape(-1, 3, 3)
        if normals is None:
            n = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
            n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
            normals = np.repeat(n[:, None, :], 3, axis=1)
        uv = np.tile(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), (tris.shape[0], 1, 1))
        self.tri.append(tris); self.nrm.append(np.asarray(normals, dtype=np.float64)); self.uv.append(uv)
        self.mat.append(np.full(tris.shape[0], material, dtype=np.uint8))
        self.light.append(np.full(tris.shape[0], bool(light)))

    def add_quad(self, p0, p1, p2, p3, material, light=False):
        """Counter-clockwise quad p0..p3 (front face = side the normal (p1-p0)x(p2-p0) points to)."""
        self.add_triangles([[p0, p1, p2], [p0, p2, p3]], material, light)

    def add_box(self, lo, hi, material, inward=False, rotation=None, center=None):
        lo, hi = np.asarray(lo, float), np.asarray(hi, float)
        c = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                      [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]])
        if rotation is not None:
            ctr = c.mean(axis=0) if center is None else np.asarray(center, float)
            c = (c - ctr) @ np.asarray(rotation).T + ctr
        faces = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (3, 0, 4, 7)]  # outward CCW
        for f in faces:
            q = [c[i] for i in (f[::-1] if inward else f)]
            self.add_quad(q[0], q[1], q[2], q[3], material)

    def finish(self, material_names):
        tris = np.concatenate(self.tri); nrm = np.concatenate(self.nrm); uv = np.concatenate(self.uv)
        mat = np.concatenate(self.mat); light = np.concatenate(self.light)
        return quantize_mesh(tris, nrm, uv, mat, light, material_names)


def encode_normal_oct16(n):
    """Inverse of decode_normal_32_bit (mesh_quantization.glsl:19-33): -1 -> 1, 0 -> 32768, 1 -> 65535."""
    n = np.asarray(n, dtype=np.float64)
    o = n[..., :2] / np.sum(np.abs(n), axis=-1, keepdims=True)
    sign = np.where(o >= 0.0, 1.0, -1.0)
    o = np.where(n[..., 2:3] <= 0.0, (1.0 - np.abs(o[..., ::-1])) * sign, o)
    return np.asarray(o * 32767.0 + 32768.5, dtype=np.uint16)


def quantize_mesh(tris, normals, uvs, material_indices, light_flags, material_names):
    """21-bit position quantisation and attribute packing as the exporter does
    (tools/io_export_vulkan_blender28.py:493-534)."""
    T = tris.shape[0]
    pos = tris.reshape(T * 3, 3)
    box_min, box_max = pos.min(axis=0), pos.max(axis=0)
    extent = np.maximum(box_max - box_min, 1e-6)
    qf = 2.0 ** 21 / extent
    q = np.asarray(pos * qf - box_min * qf, dtype=np.uint32)
    q = np.minimum(2 ** 21 - 1, q)
    dequant_factor = (1.0 / qf).astype(np.float32)
    dequant_summand = (box_min + 0.5 / qf).astype(np.float32)
    packed = np.zeros((T * 3, 2), dtype=np.uint32)
    packed[:, 0] = q[:, 0] | ((q[:, 1] & 0x7FF) << 21)
    packed[:, 1] = ((q[:, 1] & 0x1FF800) >> 11) | (q[:, 2] << 10)
    packed[:, 1] |= np.repeat(light_flags.astype(np.uint32), 3) << 31
    nuv = np.zeros((T * 3, 4), dtype=np.uint16)
    nuv[:, :2] = encode_normal_oct16(normals.reshape(T * 3, 3))
    uv = uvs.reshape(T, 3, 2).copy()
    uv -= np.floor(uv.min(axis=1))[:, None, :]
    # the shader reads v flipped: tex = (u * 8, 1 - v * 8) (shading_pass.frag.glsl:583)
    nuv[:, 2:] = np.asarray(np.clip(uv.reshape(-1, 2) * (65535.0 / 8.0) + 0.5, 0.0, 65535.0), dtype=np.uint16)
    return dict(material_names=list(material_names), positions=packed, normals_uv=nuv,
                material_indices=np.asarray(material_indices, dtype=np.uint8),
                dequant_factor=dequant_factor, dequant_summand=dequant_summand)


def dequantize_positions(mesh):
    """World-space vertices (T*3,3) float32 as the acceleration-structure build sees them (scene.c:176-187)."""
    p = mesh["positions"]
    x = (p[:, 0] & 0x1FFFFF).astype(np.float32)
    y = (((p[:, 0] & 0xFFE00000) >> 21) | ((p[:, 1] & 0x3FF) << 11)).astype(np.float32)
    z = ((p[:, 1] & 0x7FFFFC00) >> 10).astype(np.float32)
    f, s = mesh["dequant_factor"].astype(np.float32), mesh["dequant_summand"].astype(np.float32)
    return np.stack([x * f[0] + s[0], y * f[1] + s[1], z * f[2] + s[2]], axis=1)


# ---------------------------------------------------------------- camera
def look_at_camera(position, target, vertical_fov=0.33 * np.pi, near=0.05, far=1.0e3):
    """Angles such that the view direction of camera.c:24-52 points from position to target."""
    f = np.asarray(target, float) - np.asarray(position, float)
    f /= np.linalg.norm(f)
    rotation_x = float(np.arccos(np.clip(-f[2], -1.0, 1.0)))
    rotation_z = float(np.arctan2(-f[0], -f[1]))
    return dict(position=[float(x) for x in position], rotation_z=rotation_z, rotation_x=rotation_x,
                vertical_fov=float(vertical_fov), near=float(near), far=float(far), speed=2.0,
                rotate_camera=0, rotation_x_0=0.0, rotation_z_0=0.0)


# ---------------------------------------------------------------- lights
def light_rotation(angles):
    """polygonal_light.c:49-62 in float64 (used only to place emitter geometry in the mesh)."""
    cx, sx = np.cos(angles[0]), np.sin(angles[0])
    cy, sy = np.cos(angles[1]), np.sin(angles[1])
    cz, sz = np.cos(angles[2]), np.sin(angles[2])
    return np.array([[cy * cz, -cy * sz, -sy],
                     [-sx * sy * cz + cx * sz, sx * sy * sz + cx * cz, -sx * cy],
                     [cx * sy * cz + sx * sz, -cx * sy * sz + sx * cz, cx * cy]])


def light_world_vertices(light):
    r = light_rotation(light["rotation_angles"])
    v = np.asarray(light["vertices_plane_space"], float)
    scaled = v * np.array([light["scaling_x"], light["scaling_y"]])
    return np.asarray(light["translation"], float) + scaled @ r[:, :2].T


def angles_for_normal(normal, spin):
    """Euler angles whose plane normal (third rotation column, polygonal_light.c:73-76) is `normal`."""
    n = np.asarray(normal, float) / np.linalg.norm(normal)
    ay = np.arcsin(np.clip(-n[0], -1.0, 1.0))
    ax = np.arctan2(-n[1], n[2])
    return [float(ax), float(ay), float(spin)]


def add_light_geometry(builder, lights, material):
    """Emitter triangles (both windings, light bit set) so that lights are visible and sit in the BVH
    like in the reference, whose BLAS contains the emitters (shading_pass.frag.glsl:115)."""
    for light in lights:
        w = light_world_vertices(light)
        for i in range(1, w.shape[0] - 1):
            builder.add_triangles([[w[0], w[i], w[i + 1]]], material, light=True)
            builder.add_triangles([[w[0], w[i + 1], w[i]]], material, light=True)


def default_materials(count=8):
    """Flat materials: roughness from 0.1 to 1.0 (linear), metalicity alternating 0 / 1."""
    rng = np.random.default_rng(7)
    out = []
    for i in range(count):
        out.append(dict(name=f"flat{i}", base_color=[float(x) for x in rng.uniform(0.25, 0.9, 3)],
                        roughness=float(np.linspace(0.32, 1.0, count)[i]), metalicity=float(i % 2 if i >= 2 else 0)))
    return out


# ---------------------------------------------------------------- scenes
def quad_over_plane(width=640, height=360):
    """Config C1 (BASELINE.json configs[0]): one 1x1 m quad light at z = 2 facing down over a 10x10 m diffuse plane."""
    b = MeshBuilder()
    materials = [dict(name="plane", base_color=[0.8, 0.8, 0.8], roughness=1.0, metalicity=0.0),
                 dict(name="emitter", base_color=[1.0, 1.0, 1.0], roughness=1.0, metalicity=0.0)]
    b.add_quad([-5, -5, 0], [5, -5, 0], [5, 5, 0], [-5, 5, 0], 0)
    light = dict(rotation_angles=angles_for_normal([0, 0, -1], 0.0), scaling_x=1.0, scaling_y=1.0,
                 translation=[0.0, 0.0, 2.0], radiant_flux=[10.0, 10.0, 10.0],
                 vertices_plane_space=np.array([[-0.5, -0.5], [0.5, -0.5], [0.5, 0.5], [-0.5, 0.5]], dtype=np.float32))
    add_light_geometry(b, [light], 1)
    mesh = b.finish([m["name"] for m in materials])
    camera = look_at_camera([0.0, -6.0, 3.0], [0.0, 0.0, 0.6])
    return dict(name="quad_over_plane", mesh=mesh, materials=materials, camera=camera, lights=[light],
                width=width, height=height)


def _random_rotation(rng):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def many_light_room(light_count=64, box_count=200, seed=2, occluder_triangles=0, width=1920, height=1080, vertex_count=3):
    """Configs C2/C3/C5 (and C4 with occluder_triangles > 0): a 20x20x6 m room with random boxes,
    `light_count` polygonal lights on the ceiling and walls (side 0.2-1 m, radiance log-uniform in
    [1, 50] per channel) and, optionally, a field of small random triangles between floor and lights.
    vertex_count: vertices per light, or (min, max) for lights that cycle through every count of the range."""
    rng = np.random.default_rng(seed)
    materials = default_materials(8) + [dict(name="emitter", base_color=[1.0, 1.0, 1.0], roughness=1.0, metalicity=0.0)]
    emitter = len(materials) - 1
    b = MeshBuilder()
    room_lo, room_hi = np.array([-10.0, -10.0, 0.0]), np.array([10.0, 10.0, 6.0])
    b.add_box(room_lo, room_hi, 0, inward=True)
    for _ in range(box_count):
        size = rng.uniform(0.3, 1.6, 3)
        center = np.array([rng.uniform(-9, 9), rng.uniform(-9, 9), 0.0])
        center[2] = size[2] * 0.5 if rng.random() < 0.8 else rng.uniform(1.0, 4.0)
        rot = _random_rotation(rng) if rng.random() < 0.5 else None
        b.add_box(center - 0.5 * size, center + 0.5 * size, int(rng.integers(1, 8)), rotation=rot)
    lights = []
    for _ in range(light_count):
        where = rng.integers(0, 5)
        if where == 0 or where > 3:   # ceiling (more likely)
            pos = [rng.uniform(-9.5, 9.5), rng.uniform(-9.5, 9.5), 5.9]; normal = np.array([0.0, 0.0, -1.0])
        elif where == 1:
            pos = [-9.9, rng.uniform(-9.5, 9.5), rng.uniform(2.0, 5.5)]; normal = np.array([1.0, 0.0, 0.0])
        elif where == 2:
            pos = [rng.uniform(-9.5, 9.5), 9.9, rng.uniform(2.0, 5.5)]; normal = np.array([0.0, -1.0, 0.0])
        else:
            pos = [9.9, rng.uniform(-9.5, 9.5), rng.uniform(2.0, 5.5)]; normal = np.array([-1.0, 0.0, 0.0])
        normal = normal + rng.normal(scale=0.15, size=3)
        side = rng.uniform(0.2, 1.0)
        phase = rng.uniform(0, 2 * np.pi)
        vc = vertex_count if np.isscalar(vertex_count) else int(vertex_count[0] + len(lights) % (vertex_count[1] - vertex_count[0] + 1))
        ang = phase + np.arange(vc) * (2 * np.pi / vc)
        verts = np.stack([np.cos(ang), np.sin(ang)], axis=1) * (0.5 + 0.3 * rng.random((vc, 1)))
        lights.append(dict(rotation_angles=angles_for_normal(normal, rng.uniform(0, 2 * np.pi)),
                           scaling_x=float(side), scaling_y=float(side * rng.uniform(0.7, 1.3)),
                           translation=[float(x) for x in pos],
                           radiant_flux=[float(x) for x in np.exp(rng.uniform(np.log(1.0), np.log(50.0), 3))],
                           vertices_plane_space=verts.astype(np.float32)))
    add_light_geometry(b, lights, emitter)
    if occluder_triangles:
        n = int(occluder_triangles)
        centers = np.stack([rng.uniform(-9.5, 9.5, n), rng.uniform(-9.5, 9.5, n), rng.uniform(2.5, 5.0, n)], axis=1)
        offs = rng.normal(scale=0.03, size=(n, 3, 3))
        b.add_triangles(centers[:, None, :] + offs, 3)
    mesh = b.finish([m["name"] for m in materials])
    camera = look_at_camera([-8.5, -8.0, 2.6], [2.0, 3.0, 1.2])
    return dict(name=f"room_{light_count}l", mesh=mesh, materials=materials, camera=camera, lights=lights,
                width=width, height=height)


def degenerate_soup(triangle_count=20000, seed=3):
    """A mesh that is hard on a Morton-order builder: a third of the triangles are copies of one triangle, a third share
    one centroid (rotated about it), the rest are small and scattered. Quantised like every other mesh."""
    rng = np.random.default_rng(seed)
    n = triangle_count // 3
    base = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.2], [0.0, 1.0, 0.4]])
    copies = np.repeat(base[None], n, axis=0) + np.array([2.0, 2.0, 1.0])
    spun = np.stack([(base - base.mean(axis=0)) @ _random_rotation(rng).T * rng.uniform(0.1, 3.0) for _ in range(n)]) + np.array([-3.0, 1.0, 2.0])
    rest = triangle_count - 2 * n
    small = rng.uniform(-8.0, 8.0, (rest, 1, 3)) + rng.normal(size=(rest, 3, 3)) * rng.uniform(1e-3, 0.3, (rest, 1, 1))
    tris = np.concatenate([copies, spun, small]).astype(np.float64)
    normals = np.repeat(np.array([[[0.0, 0.0, 1.0]]]), triangle_count, axis=0).repeat(3, axis=1)
    uvs = np.zeros((triangle_count, 3, 2))
    return quantize_mesh(tris, normals, uvs, np.zeros(triangle_count, dtype=np.uint8), np.zeros(triangle_count, dtype=bool), ["m0"])


def material_constants(materials):
    """The values the three material textures hold (scene.h:104-118): base colour rgb (linear),
    specular data (occlusion, linear roughness, metalicity), tangent-space normal (0.5, 0.5)."""
    out = np.zeros((len(materials), 8), dtype=np.float32)
    for i, m in enumerate(materials):
        out[i, 0:3] = m["base_color"]
        out[i, 3:6] = [1.0, m["roughness"], m["metalicity"]]
        out[i, 6:8] = [0.5, 0.5]
    return out


# -------------------------------------------------------------- material textures
def _srgb_to_linear(b):
    c = np.asarray(b, dtype=np.float64) / 255.0
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def _linear_to_srgb8(x):
    x = np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0)
    s = np.where(x <= 0.0031308, 12.92 * x, 1.055 * x ** (1.0 / 2.4) - 0.055)
    return np.asarray(np.floor(s * 255.0 + 0.5), dtype=np.uint8)


def mip_chain(level0, srgb=False):
    """All mip levels of an (H, W, 4) uint8 image down to 1x1 by 2x2 box filtering (in linear light for sRGB colour)."""
    levels = [np.ascontiguousarray(level0, dtype=np.uint8)]
    while levels[-1].shape[0] > 1 or levels[-1].shape[1] > 1:
        a = levels[-1]
        h, w = max(1, a.shape[0] // 2), max(1, a.shape[1] // 2)
        lin = a.astype(np.float64) / 255.0
        if srgb:
            lin[..., :3] = _srgb_to_linear(a[..., :3])
        lin = lin[:2 * h if a.shape[0] > 1 else 1, :2 * w if a.shape[1] > 1 else 1]
        lin = lin.reshape(h, lin.shape[0] // h, w, lin.shape[1] // w, 4).mean(axis=(1, 3))
        out = np.asarray(np.floor(lin * 255.0 + 0.5), dtype=np.uint8)
        if srgb:
            out[..., :3] = _linear_to_srgb8(lin[..., :3])
        levels.append(out)
    return levels


def add_procedural_textures(scene, seed=1, size=64):
    """Gives every material of `scene` three mip-mapped 8-bit textures like the reference's assets have (scene.h:104-118):
    an sRGB checkerboard around its base colour, a specular map (occlusion, striped linear roughness, metalicity) and a
    tangent-space normal map with round bumps. Stored in scene['textures'] (3 per material) as
    {format: 'rgba8_srgb' | 'rgba8_unorm', levels: [...]}; risltc_b200.formats writes them as BC1 / BC5 .vkt files."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size]
    textures = []
    for m in scene["materials"]:
        cells = int(rng.choice([4, 8]))
        checker = ((x * cells // size) + (y * cells // size)) % 2
        base = np.asarray(m["base_color"], dtype=np.float64)
        other = np.clip(base * rng.uniform(0.35, 0.7) + rng.uniform(0.0, 0.15, 3), 0.0, 1.0)
        colour = np.where(checker[..., None] == 0, base, other)
        img = np.zeros((size, size, 4), dtype=np.uint8)
        img[..., :3] = _linear_to_srgb8(colour); img[..., 3] = 255
        textures.append(dict(format="rgba8_srgb", levels=mip_chain(img, srgb=True)))
        stripes = 0.5 + 0.5 * np.sin(2.0 * np.pi * (x + 0.5) * int(rng.integers(2, 5)) / size)
        rough = np.clip(m["roughness"] * (0.75 + 0.25 * stripes), 0.05, 1.0)
        spec = np.zeros((size, size, 4), dtype=np.uint8)
        spec[..., 0] = 255; spec[..., 1] = np.floor(rough * 255.0 + 0.5); spec[..., 2] = int(round(m["metalicity"] * 255.0)); spec[..., 3] = 255
        textures.append(dict(format="rgba8_unorm", levels=mip_chain(spec)))
        period = size // int(rng.choice([2, 4]))
        px, py = ((x % period) + 0.5) / period - 0.5, ((y % period) + 0.5) / period - 0.5
        bump = np.clip(0.2 - (px * px + py * py), 0.0, None)
        gy, gx = np.gradient(bump * 6.0)
        n = np.stack([-gx * period, -gy * period, np.ones_like(gx)], axis=-1)
        n /= np.linalg.norm(n, axis=-1, keepdims=True)
        nrm = np.zeros((size, size, 4), dtype=np.uint8)
        nrm[..., :2] = np.floor((n[..., :2] * 0.5 + 0.5) * 255.0 + 0.5); nrm[..., 2] = 255; nrm[..., 3] = 255
        textures.append(dict(format="rgba8_unorm", levels=mip_chain(nrm)))
    scene["textures"] = textures
    return scene
