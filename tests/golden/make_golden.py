#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's own sources compiled for the CPU (oracle/_ref, built by
oracle/build_ref.py from /root/reference). Run in the build container: `python tests/golden/make_golden.py`.

The reference ships no golden vectors, so these are outputs of the reference code itself:
  functions.npz  clip_polygon / calculate_ltc / prepare+sample PSA / get_ltc_coefficients / noise stream of
                 the compiled GLSL, update_polygonal_light / get_world_to_projection_space / matrix_inverse /
                 wang_random_number of the compiled C, on seeded random inputs (inputs stored with outputs)
  frames.npz     64x36 two-frame accumulations of every compiled shader variant on a small generated scene
                 (all inputs stored: mesh, materials, light records, LTC tables, constant blocks)
The oracle must reproduce every array bit for bit (tests/test_oracle_golden.py)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
HERE = Path(__file__).resolve().parent

from oracle import build_ref, orc, ref  # noqa: E402
from risltc_b200 import ltc_fit, scenes  # noqa: E402

VARIANT_ARGS = {  # compiled variant -> orc.variant keyword arguments
    "ris_ltc_v3": dict(),
    "ris_ltc_v4": dict(min_vertices=4, max_vertices=4),
    "uni_ltc_v3": dict(light_sampling="uniform"),
    "uni_psa_v4": dict(light_sampling="uniform", technique="projected_solid_angle", min_vertices=4, max_vertices=4),
    "ris_psa_v3": dict(technique="projected_solid_angle"),
    "ris_turk_v3": dict(technique="area_turk"),
    "uni_turk_v3": dict(light_sampling="uniform", technique="area_turk"),
    "ris_psa_s2l2_v3": dict(technique="projected_solid_angle", mis="balance", sample_count=2, light_samples=2),
    "ris_ltc_weighted_v3": dict(mis="weighted"),
    "ris_ltc_optimal_v3": dict(mis="optimal"),
    "uni_psa_biased_fast_v5": dict(light_sampling="uniform", technique="projected_solid_angle_biased", mis="power", fast_atan=1, min_vertices=3, max_vertices=5),
    "ris_psa_v6": dict(technique="projected_solid_angle", min_vertices=6, max_vertices=6),
    "uni_psa_v7": dict(light_sampling="uniform", technique="projected_solid_angle", min_vertices=3, max_vertices=7),
}


def random_polygons(rng, count, n_min, n_max):
    """Convex polygons in a random plane, seen from the origin, some crossing the horizon z = 0."""
    polys = np.zeros((count, 8, 3), dtype=np.float32)
    counts = rng.integers(n_min, n_max + 1, count).astype(np.uint32)
    for i in range(count):
        n = int(counts[i])
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        ang = ang[0] + np.arange(n) * (2 * np.pi / n) + rng.uniform(-0.3, 0.3, n) / n
        local = np.stack([np.cos(ang), np.sin(ang), np.zeros(n)], axis=1) * rng.uniform(0.2, 2.0)
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w, x, y, z = q
        rot = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        center = rng.normal(size=3) * 1.5 + np.array([0, 0, rng.uniform(-0.5, 2.5)])
        polys[i, :n] = (local @ rot.T + center).astype(np.float32)
    return polys, counts


def functions():
    rng = np.random.default_rng(11)
    out = {}
    # clip + LTC integral + PSA per compiled MAX_POLYGON_VERTEX_COUNT (P = V_max + 1)
    for name, n_lo, n_hi in (("ris_ltc_v3", 3, 3), ("ris_ltc_v4", 4, 4), ("uni_psa_biased_fast_v5", 3, 5), ("uni_psa_v7", 3, 7)):
        r = ref.RefShading(name)
        polys, counts = random_polygons(rng, 400, n_lo, n_hi)
        clipped = np.zeros_like(polys); vcs = np.zeros(len(counts), dtype=np.uint32); ltc = np.zeros(len(counts), dtype=np.float32)
        psa_poly = np.zeros((len(counts), 44), dtype=np.float32); psa_dir = np.zeros((len(counts), 3), dtype=np.float32)
        rnd = rng.random((len(counts), 2)).astype(np.float32)
        for i in range(len(counts)):
            vc, buf = r.clip_polygon(int(counts[i]), polys[i])
            vcs[i] = vc; clipped[i] = buf
            # calculate_ltc loops over MAX_POLYGON_VERTEX_COUNT edges (polygon_sampling.glsl:523-530); clip_polygon
            # repeats vertex 0 at slot [vc], so the result is defined iff vc >= P - 1 (no stale slot is read)
            ltc[i] = r.calculate_ltc(vc, buf) if vc > 0 and vc + 1 >= r.max_polygon_vertices else np.float32(np.nan)
            if vc > 0:
                psa_poly[i], psa_dir[i] = r.psa(vc, buf, float(rnd[i, 0]), float(rnd[i, 1]))
        key = name
        out[f"{key}.polygons"] = polys; out[f"{key}.counts"] = counts; out[f"{key}.clipped"] = clipped; out[f"{key}.clipped_counts"] = vcs
        out[f"{key}.ltc_integral"] = ltc; out[f"{key}.randoms"] = rnd; out[f"{key}.psa_polygon"] = psa_poly; out[f"{key}.psa_dir"] = psa_dir
    # noise
    r = ref.RefShading("ris_ltc_v3")
    words = [orc.wang(4 * f) for f in range(3)]
    out["noise.frame_words"] = np.array(words, dtype=np.uint32)
    out["noise.pixels"] = np.array([[0, 0], [1, 0], [0, 1], [639, 359], [17, 200]], dtype=np.uint32)
    out["noise.draws"] = np.stack([np.stack([r.noise(int(px), int(py), 640, w, 8) for px, py in out["noise.pixels"]]) for w in words])
    # LTC coefficients (needs bound tables)
    fits = ltc_fit.fit_ggx_ltc(16, 6, 16)
    rgba, rg = ltc_fit.quantize_fits(fits)
    scene = scenes.many_light_room(4, 2, width=32, height=18)
    osc = orc.OracleScene(scene, rgba, rg)
    r.bind(osc)
    consts = (C.c_float * 8)()
    orc.lib().orc_ltc_constants(consts, C.c_uint32(16), C.c_uint32(16), C.c_uint32(6))
    c6 = np.array(list(consts)[:6], dtype=np.float32)
    n = 200
    inputs = np.zeros((n, 11), dtype=np.float32)
    inputs[:, 0] = rng.random(n); inputs[:, 1] = rng.uniform(0.0064, 1.0, n); inputs[:, 2:5] = rng.normal(size=(n, 3)) * 3
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    outg = nrm + rng.normal(size=(n, 3)) * 0.8; outg /= np.linalg.norm(outg, axis=1, keepdims=True)
    flip = np.sum(nrm * outg, axis=1) < 0.05
    outg[flip] = nrm[flip]
    inputs[:, 5:8] = nrm; inputs[:, 8:11] = outg
    out["ltc.rgba16"] = rgba; out["ltc.rg16"] = rg; out["ltc.constants"] = c6; out["ltc.inputs"] = inputs
    out["ltc.coefficients"] = np.stack([r.ltc_coefficients(inputs[i, 0], inputs[i, 1], inputs[i, 2:5], inputs[i, 5:8], inputs[i, 8:11], c6) for i in range(n)])
    # host C: polygonal_light.c, camera.c, math_utilities.h
    h = ref.RefHost()
    lights = scenes.many_light_room(12, 0, seed=5, vertex_count=5)["lights"]
    out["host.light_inputs"] = np.array([list(l["rotation_angles"]) + [l["scaling_x"], l["scaling_y"]] + list(l["translation"]) + list(l["radiant_flux"]) for l in lights], dtype=np.float32)
    out["host.light_plane_vertices"] = np.stack([np.asarray(l["vertices_plane_space"], dtype=np.float32) for l in lights])
    upd = [h.update_light(l) for l in lights]
    out["host.light_world"] = np.stack([u[0] for u in upd]); out["host.light_plane"] = np.stack([u[1] for u in upd])
    out["host.light_radiance"] = np.stack([u[2] for u in upd]); out["host.light_area"] = np.array([u[3] for u in upd], dtype=np.float32)
    cam = scenes.look_at_camera([-8.5, -8.0, 2.6], [2.0, 3.0, 1.2])
    out["host.camera"] = np.array(list(cam["position"]) + [cam["rotation_x"], cam["rotation_z"], cam["vertical_fov"], cam["near"], cam["far"]], dtype=np.float32)
    out["host.world_to_projection"] = h.world_to_projection(cam, float(np.float32(16) / np.float32(9)))
    m = rng.normal(size=(4, 4)).astype(np.float32)
    out["host.matrix"] = m; out["host.matrix_inverse"] = h.matrix_inverse(m)
    out["host.wang_seeds"] = np.array([0, 1, 4, 8, 12345, 0xFFFFFFFF], dtype=np.uint32)
    out["host.wang"] = np.array([h.wang(int(s)) for s in out["host.wang_seeds"]], dtype=np.uint32)
    np.savez_compressed(HERE / "functions.npz", **out)
    print("functions.npz", sum(v.nbytes for v in out.values()), "bytes raw")


def frames():
    W, H, F = 64, 36, 2
    fits = ltc_fit.fit_ggx_ltc(16, 6, 16)
    rgba, rg = ltc_fit.quantize_fits(fits)
    out = {"ltc.rgba16": rgba, "ltc.rg16": rg}
    for verts in (3, 4, 5, 6, 7):
        scene = scenes.many_light_room(12, 10, seed=3, width=W, height=H, vertex_count=verts)
        if verts == 5:   # mixed vertex counts (MIN < MAX)
            for l in scene["lights"][::2]:
                l["vertices_plane_space"] = l["vertices_plane_space"][:3]
        if verts == 7:   # every count from 3 to 7
            for i, l in enumerate(scene["lights"]):
                l["vertices_plane_space"] = l["vertices_plane_space"][:3 + i % 5]
        osc = orc.OracleScene(scene, rgba, rg)
        cs = [orc.make_constants(scene, W, H, orc.frame_words(f)[0], ltc_res=16, ltc_layers=6) for f in range(F)]
        k = f"scene_v{verts}"
        out[f"{k}.positions"] = osc.positions; out[f"{k}.normals_uv"] = osc.normals_uv; out[f"{k}.material_indices"] = osc.material_indices
        out[f"{k}.dequant_factor"] = np.asarray(scene["mesh"]["dequant_factor"], dtype=np.float32)
        out[f"{k}.dequant_summand"] = np.asarray(scene["mesh"]["dequant_summand"], dtype=np.float32)
        out[f"{k}.materials"] = osc.materials; out[f"{k}.records"] = osc.records; out[f"{k}.min_vertices"] = np.uint32(osc.min_vertices)
        out[f"{k}.constants"] = np.frombuffer(b"".join(bytes(C.string_at(C.byref(c), 256)) for c in cs), dtype=np.uint8).reshape(F, 256)
        for name in VARIANT_ARGS:
            if not name.endswith(f"_v{verts}"):
                continue
            r = ref.RefShading(name); r.bind(osc)
            accum, vis, rays = r.render(cs)
            out[f"{name}.accum"] = accum; out[f"{name}.rays"] = np.uint64(rays)
            out[f"{k}.visibility"] = vis
    # textured materials: mip-mapped 8-bit textures (sRGB base colour, specular, normal) sampled by the oracle's textureGrad
    # (texture filtering is driver code); the compiled reference GLSL computes the texture-coordinate derivatives
    scene = scenes.add_procedural_textures(scenes.many_light_room(12, 10, seed=3, width=W, height=H), seed=1, size=32)
    osc = orc.OracleScene(scene, rgba, rg)
    cs = [orc.make_constants(scene, W, H, orc.frame_words(f)[0], ltc_res=16, ltc_layers=6) for f in range(F)]
    out["scene_tex.constants"] = np.frombuffer(b"".join(bytes(C.string_at(C.byref(c), 256)) for c in cs), dtype=np.uint8).reshape(F, 256)
    out["scene_tex.formats"] = np.array([orc.TEXEL[t["format"]] for t in scene["textures"]], dtype=np.uint32)
    out["scene_tex.level_counts"] = np.array([len(t["levels"]) for t in scene["textures"]], dtype=np.uint32)
    out["scene_tex.texels"] = np.concatenate([l.reshape(-1) for t in scene["textures"] for l in t["levels"]])
    r = ref.RefShading("ris_ltc_v3"); r.bind(osc)
    accum, vis, rays = r.render(cs)
    out["scene_tex.accum"] = accum; out["scene_tex.rays"] = np.uint64(rays)
    np.savez_compressed(HERE / "frames.npz", **out)
    print("frames.npz", sum(np.asarray(v).nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    if not ref.host_available():
        build_ref.main()
    functions()
    frames()
