/* risltc_oracle.h -- CPU oracle for the risltc shading hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and there only as the checker / CPU baseline.
 *
 * This is a plain C99, IEEE fp32 restatement (no FMA contraction except where
 * the reference writes fma()) of the GLSL algorithm in
 *   /root/reference/src/shaders/shading_pass.frag.glsl (whole file) and the
 *   helper files it includes (polygon_sampling.glsl, polygon_clipping.glsl,
 *   ltc_utility.glsl, brdfs.glsl, noise_utility.glsl, reservoir.glsl,
 *   mesh_quantization.glsl, polygon_sampling_related_work.glsl:34-64,
 *   accum_pass.frag.glsl:34-55, visibility_pass.*),
 * plus the host arithmetic that feeds it (main.c:456-490, 2902-2946,
 * ltc_table.c:82-116,184-191, polygonal_light.c:44-98, camera.c:24-83,
 * noise_table.c:24-28, scene.c:176-187).
 *
 * PARITY STATUS: the reference ships no tests, golden vectors or images
 * (SURVEY.md section 4). The shader arithmetic is pinned instead against the
 * reference's own GLSL sources compiled as C++ on the CPU (oracle/_ref, built by
 * oracle/build_ref.py from /root/reference, outputs only in oracle/_ref/) and
 * against the reference's own polygonal_light.c / camera.c compiled as-is.
 * Two boundaries stay "parity unpinned" because they are driver code that is
 * not in the tree: the rasteriser / ray-query BVH (defined here as pixel-centre
 * closest-hit and any-hit Moeller-Trumbore on the dequantised triangles) and
 * texture-unit filtering (defined here as exact fp32 bilinear).
 */
#ifndef RISLTC_ORACLE_H
#define RISLTC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Largest supported MAX_POLYGON_VERTEX_COUNT (main.c:191-204: V_max + 1, V_max <= 7). */
#define ORC_MAX_P 8

/* render_settings_t enums, values as in main.h:42-83 and polygonal_light.h:29-46 */
enum { ORC_MIS_BALANCE = 0, ORC_MIS_POWER, ORC_MIS_WEIGHTED, ORC_MIS_OPTIMAL_CLAMPED, ORC_MIS_OPTIMAL };
enum { ORC_LIGHT_UNIFORM = 0, ORC_LIGHT_RESERVOIR = 1 };
enum { ORC_POLY_BASELINE = 0, ORC_POLY_AREA_TURK, ORC_POLY_PSA, ORC_POLY_PSA_BIASED, ORC_POLY_LTC_CP };

/* The compile-time -D table of main.c:962-991 as a run-time struct. */
typedef struct orc_variant_s {
	uint32_t light_sampling;       /* ORC_LIGHT_* */
	uint32_t polygon_technique;    /* ORC_POLY_* */
	uint32_t mis_heuristic;        /* ORC_MIS_* */
	uint32_t sample_count;         /* SAMPLE_COUNT */
	uint32_t light_samples;        /* LIGHT_SAMPLES */
	uint32_t fast_atan;            /* USE_FAST_ATAN */
	uint32_t min_light_vertices;   /* MIN_POLYGON_VERTEX_COUNT_BEFORE_CLIPPING */
	uint32_t max_light_vertices;   /* MAX_POLYGONAL_LIGHT_VERTEX_COUNT */
} orc_variant_t;

/* per_frame_constants_t, byte-identical to main.h:537-553 (256 bytes). */
typedef struct orc_constants_s {
	float dequant_factor[3], pad0, dequant_summand[3];
	float error_factor;
	float world_to_projection[4][4];
	float pixel_to_ray[3][4];
	float camera_position[3];
	float mis_visibility_estimate;
	uint32_t viewport[2];
	int32_t cursor[2];
	float exposure_factor;
	float roughness_factor;
	uint32_t noise_resolution_mask[2];
	uint32_t noise_texture_index_mask;
	uint32_t pad3[3];
	uint32_t noise_random_numbers[4];
	float ltc_constants[8];
} orc_constants_t;

/* Flat-colour material: what the three textureGrad fetches of
 * shading_pass.frag.glsl:630-633 return (linear RGB base colour, specular
 * data (occlusion, linear roughness, metalicity), tangent-space normal .rg). */
typedef struct orc_material_s {
	float base_color[3];
	float specular[3];
	float normal_rg[2];
} orc_material_t;

/* A material texture as the product keeps it after loading (tables_scene.c decodes BC1 / BC5 blocks to 8-bit texels): every
 * mip level, largest first, tightly packed, RGBA per texel. Sampled by orc_sample_texture_grad. */
enum { ORC_TEXEL_RGBA32F = 0, ORC_TEXEL_RGBA8_UNORM = 1, ORC_TEXEL_RGBA8_SRGB = 2 };
typedef struct orc_texture_s {
	uint32_t format, width, height, mip_count;
	const void* texels;
} orc_texture_t;

typedef struct orc_bvh_s orc_bvh_t;

typedef struct orc_scene_s {
	uint64_t triangle_count;
	float dequant_factor[3], dequant_summand[3]; /* mesh_t.dequantization_* (scene.h:50-53) */
	const uint32_t* positions;        /* 2*3*T words, mesh_t.positions layout (scene.h:58-65) */
	const uint16_t* normals_uv;       /* 4*3*T UNORM16 (scene.h:66-71) */
	const uint8_t* material_indices;  /* T bytes (scene.c:58) */
	uint64_t material_count;
	const orc_material_t* materials;
	uint32_t light_count;
	const float* light_records;       /* write_lights stream, stride 12+4*V_max floats (main.c:456-490) */
	uint32_t ltc_res;                 /* roughness_count == inclination_count */
	uint32_t ltc_layers;              /* fresnel_count */
	const uint16_t* ltc_rgba16;       /* layers*res*res*4 */
	const uint16_t* ltc_rg16;         /* layers*res*res*2 */
	orc_bvh_t* bvh;                   /* built by orc_build_bvh */
	const orc_texture_t* textures;    /* NULL: flat materials (above); else 3 per material: base colour, specular, normal (scene.h:104-118) */
} orc_scene_t;

/* textureGrad with the sampler of scene.c:546-552 (linear mag / min / mip filters, repeat addressing) under the oracle's stated
 * definition -- texture filtering is driver code, anisotropy (maxAnisotropy 16) is NOT modelled:
 *   level of detail lambda = log2(max(|d(u W, v H)/dx|, |d(u W, v H)/dy|)) (Vulkan 1.2, 16.5.5-16.5.7, isotropic), clamped to
 *   [0, mip_count - 1]; bilinear taps at levels floor(lambda) and floor(lambda) + 1 mixed by its fraction; a tap is
 *   c00 + fx (c10 - c00) etc. at x = u w - 0.5 with repeat addressing, in fp32; 8-bit texels are c / 255 (correctly rounded),
 *   sRGB texels go through the exact sRGB curve evaluated in double precision (alpha stays linear). log2 is correctly
 *   rounded like the other transcendental functions. A texture whose texels are all equal returns exactly that texel. */
void orc_sample_texture_grad(const orc_texture_t* texture, const float uv[2], const float ddx[2], const float ddy[2], float rgba[4]);
/* table entry i: sRGB byte -> linear float (the table the product uploads is computed by the same formula) */
float orc_srgb8_to_linear(uint32_t byte);

/* ---- noise (noise_utility.glsl:26-95, math_utilities.h:50-57) ---- */
uint32_t orc_wang_random_number(uint32_t seed);
uint32_t orc_noise_seed(uint32_t px, uint32_t py, uint32_t width, uint32_t frame_word);
float orc_noise_next(uint32_t* seed);

/* ---- polygon kernels exposed for known-answer tests ---- */
/* v holds ORC_MAX_P xyz triples; returns clipped count (polygon_clipping.glsl:35-225) */
uint32_t orc_clip_polygon(uint32_t vertex_count, float v[ORC_MAX_P][3], uint32_t min_vertices, uint32_t max_polygon_vertices);
float orc_calculate_ltc(uint32_t vertex_count, const float v[ORC_MAX_P][3]);

typedef struct orc_psa_polygon_s {
	uint32_t vertex_count;
	float vertices[ORC_MAX_P][2];
	float ellipses[ORC_MAX_P][2];
	float inner_ellipse_0[2];
	float sector_projected_solid_angles[ORC_MAX_P];
	float projected_solid_angle;
} orc_psa_polygon_t;

void orc_prepare_psa(orc_psa_polygon_t* out, uint32_t vertex_count, const float v[ORC_MAX_P][3], uint32_t max_polygon_vertices, uint32_t fast_atan);
void orc_sample_psa(float out_dir[3], const orc_psa_polygon_t* polygon, float u0, float u1, uint32_t max_polygon_vertices, uint32_t fast_atan, uint32_t biased);

/* ---- LTC ---- */
typedef struct orc_ltc_s {
	float world_to_shading[4][3];   /* mat4x3, [column][row] */
	float shading_to_cosine[3][3];  /* mat3,   [column][row] */
	float cosine_to_shading[3][3];
	float albedo;
	float determinant;
} orc_ltc_t;
void orc_get_ltc_coefficients(orc_ltc_t* out, const orc_scene_t* scene, float fresnel_0, float roughness,
	const float position[3], const float normal[3], const float outgoing[3], const float ltc_constants[6]);
/* ltc_table.c:82-116: one fit record (a,b,c,d,albedo) -> 4+2 UNORM16 */
void orc_quantize_ltc_fit(const float fit[5], uint16_t rgba[4], uint16_t rg[2]);
void orc_ltc_constants(float out[8], uint32_t roughness_count, uint32_t inclination_count, uint32_t fresnel_count);

/* ---- host arithmetic ---- */
/* polygonal_light.c:44-98 on flat arrays. plane_space / world are float[4]-strided. */
void orc_update_polygonal_light(const float rotation_angles[3], float scaling_x, float scaling_y, const float translation[3],
	const float radiant_flux[3], uint32_t vertex_count, const float* vertices_plane_space,
	float* vertices_world_space, float plane[4], float surface_radiance[3], float* area, float rotation[3][4]);
void orc_world_to_projection(float out[4][4], const float position[3], float rotation_x, float rotation_z, float vertical_fov, float near_plane, float far_plane, float aspect);
void orc_pixel_to_ray(float out[3][4], const float world_to_projection[4][4], uint32_t width, uint32_t height);

/* ---- frame ---- */
int orc_build_bvh(orc_scene_t* scene);
void orc_free_bvh(orc_scene_t* scene);
/* Primary visibility (visibility_pass.*, main.c:715-721,751-756,2024): u32 per pixel. */
void orc_visibility_pass(const orc_scene_t* scene, const orc_constants_t* constants, uint32_t* visibility,
	uint32_t row_begin, uint32_t row_end);
/* Shading (shading_pass.frag.glsl:674-770) for rows [row_begin,row_end): RGBA32F per pixel. */
void orc_shading_pass(const orc_scene_t* scene, const orc_constants_t* constants, const orc_variant_t* variant,
	const uint32_t* visibility, float* shaded_rgba, uint32_t row_begin, uint32_t row_end);
/* accum_pass.frag.glsl:45-53 */
void orc_accum_pass(float* accum_rgba, const float* shaded_rgba, uint32_t accum_num, uint64_t pixel_count);
/* All three for one frame; accum is updated in place. Returns the number of shadow rays traced. */
uint64_t orc_render_frame(const orc_scene_t* scene, const orc_constants_t* constants, const orc_variant_t* variant,
	uint32_t accum_num, float* accum_rgba, uint32_t* visibility_out);
/* Any-hit query used by the shading pass, exposed for BVH parity tests. */
int orc_any_hit(const orc_scene_t* scene, const float origin[3], const float dir[3], float t_min, float t_max);
int orc_thread_count(void);

/* copy_pass.frag.glsl:28-58 + the 8-bit swapchain write: frame_bits 0 = displayed sRGB image, 1 / 2 = low / high byte of the
 * half-float bits (the two LDR frames of an HDR screenshot, main.c:2339-2350). rgb8: pixel_count x 3 bytes. */
void orc_copy_pass(const float* rgba, uint64_t pixel_count, uint32_t frame_bits, uint8_t* rgb8);

#ifdef __cplusplus
}
#endif
#endif
