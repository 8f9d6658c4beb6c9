/* risltc_cuda.h -- C ABI of librisltc_cuda.so, the B200 (sm_100a) implementation of
 * risltc's per-pixel shading path. Plain C: pointers and sizes only.
 *
 * The reference has no FFI for this path; its host code (main.c) talks to Vulkan
 * directly. Each entry point below replaces one block of that host code, cited as
 * file:line under /root/reference/src. A maintainer wires them in as shown in
 * INTEGRATION.md. Conventions follow the reference: every function that can fail
 * returns int, 0 = success, non-zero = failure after printf of a one-line message
 * (ltc_table.c:38-43, scene.c:414-418); destroy tolerates NULL / partially built
 * objects; the library is single-threaded like the reference (one stream per
 * device object) and never spawns host threads that outlive a call.
 *
 * There is no CPU fallback: when no CUDA device is usable every call fails. */
#ifndef RISLTC_CUDA_H
#define RISLTC_CUDA_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct risltc_device_s risltc_device_t;

/* Shader-variant selection. Replaces the -D table that specialises
 * shading_pass.frag.glsl (main.c:962-991); enum values are those of
 * main.h:42-83 (mis_heuristic_t, light_sampling_strategies_t) and
 * polygonal_light.h:29-46 (sample_polygon_technique_t). */
typedef struct risltc_variant_s {
	uint32_t light_sampling;      /* 0 light_uniform, 1 light_reservoir */
	uint32_t polygon_technique;   /* 0 baseline, 1 area_turk, 2 projected_solid_angle, 3 ..._biased, 4 ltc_cp */
	uint32_t mis_heuristic;       /* 0 balance, 1 power, 2 weighted, 3 optimal_clamped, 4 optimal */
	uint32_t sample_count;        /* SAMPLE_COUNT (render_settings_t.sample_count) */
	uint32_t light_samples;       /* LIGHT_SAMPLES (render_settings_t.sample_count_light) */
	uint32_t fast_atan;           /* USE_FAST_ATAN */
	uint32_t min_light_vertices;  /* MIN_POLYGON_VERTEX_COUNT_BEFORE_CLIPPING (main.c:944) */
	uint32_t max_light_vertices;  /* MAX_POLYGONAL_LIGHT_VERTEX_COUNT (main.c:945) */
} risltc_variant_t;

/* create_vulkan_device / destroy_vulkan_device (vulkan_basics.c:24, main.c:2569-2580). */
int risltc_cuda_create_device(risltc_device_t** device, int cuda_ordinal);
void risltc_cuda_destroy_device(risltc_device_t* device);
const char* risltc_cuda_last_error(void);

/* Mesh upload + acceleration structure: load_scene's device part (scene.c:484-559) and
 * create_acceleration_structure (scene.c:142-406, dequantisation :176-187). The three
 * arrays are the .vks payload / mesh_t buffers (scene.h:58-75): 2*3*T u32 positions,
 * 4*3*T u16 normals+uv, T u8 material indices. */
int risltc_cuda_upload_scene(risltc_device_t* device, const uint32_t* quantized_positions,
	const uint16_t* normals_and_tex_coords, const uint8_t* material_indices, uint64_t triangle_count,
	const float dequantization_factor[3], const float dequantization_summand[3]);

/* Material textures (scene.c:520-552, sampled at shading_pass.frag.glsl:629-633). One record of
 * 8 floats per material: base colour rgb (linear), specular texel rgb (occlusion, linear
 * roughness, metalicity), normal-map texel rg: materials without texture detail (1 x 1 textures); the kernels then skip
 * the texture-coordinate derivatives and the three filtered fetches. Replaces the textures of upload_textures. */
int risltc_cuda_upload_materials(risltc_device_t* device, const float* material_constants, uint64_t material_count);

/* Material textures with their mip chains (load_2d_textures, textures.c:95-241; sampler scene.c:546-552): 3 per material in the
 * order base colour, specular, normal (scene.h:104-118). Every level is decoded to RGBA texels (the host layer decodes the
 * BC1 / BC5 blocks of *.vkt files), largest level first, tightly packed. Sampled like textureGrad under the stated
 * definition of DESIGN.md (isotropic level of detail, trilinear, repeat). Replaces the constants of upload_materials. */
#define RISLTC_TEXEL_RGBA32F 0u
#define RISLTC_TEXEL_RGBA8_UNORM 1u
#define RISLTC_TEXEL_RGBA8_SRGB 2u      /* r, g, b through the sRGB curve, alpha linear */
typedef struct risltc_texture_s {
	uint32_t format, width, height, mip_count;
	const void* texels;
} risltc_texture_t;
int risltc_cuda_upload_textures(risltc_device_t* device, const risltc_texture_t* textures, uint64_t texture_count);

/* Light buffer: the byte stream write_lights produces (main.c:456-490): per light 12 floats
 * {radiance.xyz, pad, plane.xyzw, vertex_count(u32), pad x3} + max_vertex_count x {x, y, z, pad}. */
int risltc_cuda_upload_lights(risltc_device_t* device, const void* light_records, uint32_t light_count, uint32_t max_vertex_count);

/* LTC tables: the two staging arrays load_ltc_table fills (ltc_table.c:82-116) before the
 * image copy (ltc_table.c:143-166): fresnel_count layers of res x res RGBA16 / RG16 UNORM. */
int risltc_cuda_upload_ltc(risltc_device_t* device, const uint16_t* rgba16, const uint16_t* rg16,
	uint32_t roughness_count, uint32_t inclination_count, uint32_t fresnel_count);

/* change_shading / create_shading_pass (main.c:2498, 937-1010). */
int risltc_cuda_set_variant(risltc_device_t* device, const risltc_variant_t* variant);

/* Arithmetic mode of the shading kernel. The reference compiles its GLSL without a precision contract
 * (drivers contract a*b+c and use approximate rsqrt / rcp freely), so two modes are offered:
 * RISLTC_PRECISION_FAST (default): fused multiply-adds, MUFU rsqrt / rcp, lights staged in shared memory;
 * image within BASELINE.json's tolerance of the oracle (rel. RMSE <= 1e-3 converged, >= 99 % pixels).
 * RISLTC_PRECISION_EXACT: every operation rounded like the C oracle (no contraction, IEEE div / sqrt);
 * image bit-identical to the oracle up to libm ulps. Replaces nothing in the reference (its shader
 * compiler decides); documented here because it selects between two compiled kernel sets. */
#define RISLTC_PRECISION_FAST 0u
#define RISLTC_PRECISION_EXACT 1u
int risltc_cuda_set_precision(risltc_device_t* device, uint32_t mode);

/* Builder of the acceleration structures (create_acceleration_structure, scene.c:142-406, where the driver builds on the
 * device): HOST = binned SAH, the top levels split over tasks (5 M triangles: 2.7 s on 16 cores, 7-8 s on one), DEVICE = built by kernels
 * (bvh_gpu.cu): Morton order + PLOC clustering, 5 M triangles in 24 ms; shadow rays cost -3 .. +9 % against the SAH tree,
 * the per-pixel BVH walk of huge scenes +40 %; DEVICE_RADIX = the plain radix tree over the Morton order (13 ms, shadow
 * rays 1.2-1.5x costlier). AUTO (default) = DEVICE from 8 M triangles on. Takes effect at the next upload_scene; images do
 * not depend on it. Environment: RISLTC_BVH_BUILD=host|gpu|radix, RISLTC_BVH_PLOC_RADIUS=<1..32> (16).
 * bvh_stats: {builder used, wall ms of the build, device ms of sort + hierarchy / boxes + records / 4-wide collapse,
 * binary node slots, 4-wide nodes, binary depth << 16 | 4-wide depth}. */
#define RISLTC_BVH_BUILDER_HOST 0u
#define RISLTC_BVH_BUILDER_DEVICE 1u
#define RISLTC_BVH_BUILDER_AUTO 2u
#define RISLTC_BVH_BUILDER_DEVICE_RADIX 3u
int risltc_cuda_set_bvh_builder(risltc_device_t* device, uint32_t builder);
int risltc_cuda_bvh_stats(risltc_device_t* device, double stats[8]);

/* Two of the passes have two implementations each that produce bit-identical buffers (tests/test_gpu_frames.py):
 *   visibility pass (visibility_pass.*.glsl): the triangle-parallel rasteriser or the per-pixel BVH walk; AUTO (default)
 *     times both on the first two frames after upload_scene / resize and keeps the faster;
 *   ray queries (shading_pass.frag.glsl:112-129): the 4-wide tree with 8-bit boxes, traversed once for the two rays of
 *     a pixel (PAIRS, default) or once per ray (WIDE), or the binary tree.
 * This call pins them (the environment variables RISLTC_GBUFFER=raster|bvh and RISLTC_TRACE=8|4|2 do the same at create).
 * The shadow-ray kernel additionally times two settings of its triangle-track threshold on the first two frames after
 * upload_scene and keeps the faster (RISLTC_TRI_VOTE=<n> pins it); results do not depend on it either. */
#define RISLTC_GBUFFER_BVH 0u
#define RISLTC_GBUFFER_RASTER 1u
#define RISLTC_GBUFFER_AUTO 2u
#define RISLTC_SHADOW_BINARY 2u
#define RISLTC_SHADOW_WIDE 4u
#define RISLTC_SHADOW_PAIRS 8u
int risltc_cuda_set_kernels(risltc_device_t* device, uint32_t gbuffer, uint32_t shadow);

/* Frame overlap inside render_frames: consecutive frames alternate between two streams and two sets of per-frame buffers
 * (only the accumulation stays ordered), which fills the tails of the persistent kernels. AUTO (default): on when the
 * device renders at most 4.5 M pixels (where it measures 2-19 % faster), off for larger shares (a whole 4K frame: no
 * gain) -- only then do last_kernel_ms() / last_pass_ms() time each pass in isolation. The image does not depend on the
 * mode. Environment: RISLTC_OVERLAP=0|1. frame_overlap_active: 1 if the frames of the next render_frames call will overlap. */
#define RISLTC_OVERLAP_OFF 0u
#define RISLTC_OVERLAP_ON 1u
#define RISLTC_OVERLAP_AUTO 2u
int risltc_cuda_set_frame_overlap(risltc_device_t* device, uint32_t mode);
uint32_t risltc_cuda_frame_overlap_active(const risltc_device_t* device);

/* Render targets (create_render_targets, main.c:246-330) for a width x height frame of which this
 * device renders the rows y with (y / stripe_height) % stripe_count == stripe_index
 * (stripe_count = 1: the whole frame). Resets the accumulation buffer (and detaches a buffer given
 * to risltc_cuda_set_accum_buffer). */
int risltc_cuda_resize(risltc_device_t* device, uint32_t width, uint32_t height,
	uint32_t stripe_height, uint32_t stripe_index, uint32_t stripe_count);

/* Let the accumulation target live in caller-owned device memory (e.g. a torch tensor that an
 * NCCL gather reads); layout owned_rows x width x RGBA32F. NULL returns to the device's own buffer.
 * risltc_cuda_resize recreates all render targets and therefore detaches a caller-owned buffer:
 * attach again after every resize. */
int risltc_cuda_set_accum_buffer(risltc_device_t* device, void* device_pointer);
uint32_t risltc_cuda_owned_rows(const risltc_device_t* device);

/* One frame = the four subpasses record_render_frame_commands records (main.c:2010-2092):
 * visibility -> shading -> (shadow rays) -> accumulation. `constants` is the 256-byte
 * per_frame_constants_t block write_constants fills (main.c:2902-2946, main.h:537-553);
 * accum_num is the push constant of the accumulation pass (main.c:2062-2064). Asynchronous on the
 * device's stream; timing is taken with events around the whole frame like the reference's
 * timestamps (main.c:2038-2040, 2084). */
int risltc_cuda_render_frame(risltc_device_t* device, const void* per_frame_constants, uint32_t accum_num);
/* `frame_count` frames back to back from an array of constant blocks, accum_num = first + i. */
int risltc_cuda_render_frames(risltc_device_t* device, const void* per_frame_constants_array, uint32_t frame_count, uint32_t first_accum_num);

/* Blocking read-backs (implement_screenshot's staging copy, main.c:2358-2409). rgba has
 * owned_rows x width x 4 floats, in the order of the owned rows. */
int risltc_cuda_synchronize(risltc_device_t* device);
int risltc_cuda_read_accum(risltc_device_t* device, float* rgba);
int risltc_cuda_read_visibility(risltc_device_t* device, uint32_t* primitive_ids);
/* The copy pass (copy_pass.frag.glsl:28-58) and the 8-bit swapchain write behind it: the accumulated frame as
 * owned_rows x width x 3 bytes. frame_bits 0: the displayed image (clamp, linear -> sRGB, srgb_utility.glsl:20-34);
 * frame_bits 1 / 2: the low / high byte of every channel's half-float bits -- the two LDR frames that
 * implement_screenshot combines into one HDR screenshot (main.c:2339-2350, 2358-2409). Blocking. */
int risltc_cuda_copy_pass(risltc_device_t* device, uint32_t frame_bits, uint8_t* rgb8);
/* Global row index of every owned row (owned_rows entries). */
int risltc_cuda_owned_row_indices(const risltc_device_t* device, uint32_t* rows);

/* record_frame_time (frame_timer.c:37-55): milliseconds of the last render_frame(s) call
 * (blocks until it has finished); ms[4] = visibility, shading, shadow+accumulate kernels summed
 * over the frames of the call (CUDA events on the device's stream around every launch), and the whole call. */
float risltc_cuda_last_frame_ms(risltc_device_t* device);
int risltc_cuda_last_kernel_ms(risltc_device_t* device, float ms[4]);
/* The five passes of the last call separately: ms[6] = visibility (1), RIS candidates (2a), winner's estimator (2b), shadow
 * rays (3), MIS sum + accumulation (4), each summed over the frames of the call, and the whole call. Variants that run
 * the generic kernel report all of (2) in ms[1]. With frame overlap on the intervals of consecutive frames overlap. */
int risltc_cuda_last_pass_ms(risltc_device_t* device, float ms[6]);
/* Traversal statistics of the shadow-ray kernel (bytes per ray from traversal counters). `enable` != 0
 * switches the kernel to its counting instantiation for the following calls; `counters` (may be NULL) receives those of
 * the last render call that ran with counting on: rays traced, 4-wide nodes visited (64 B each), triangles tested (48 B
 * each), rays found occluded. Counting costs a few percent; it is off by default. */
int risltc_cuda_traversal_counters(risltc_device_t* device, uint32_t enable, uint64_t counters[4]);
/* Counters of the last frame: [0] covered (non-background) pixel-samples, [1] shadow rays traced,
 * [2] kernels launched since create, [3] RIS candidates evaluated. */
int risltc_cuda_counters(risltc_device_t* device, uint64_t counters[4]);
void* risltc_cuda_stream(risltc_device_t* device);

/* Known-answer entry points: run the device functions of the shading kernel on arrays so that tests
 * can compare them with the oracle one function at a time (SURVEY 8c "setup agrees to <= 1e-5").
 * polygons: count x 8 x 3 floats (vertex slots), vertex_counts: count. */
int risltc_cuda_kat_clip(risltc_device_t* device, float* polygons, uint32_t* vertex_counts, uint32_t count, uint32_t max_light_vertices);
int risltc_cuda_kat_ltc_integral(risltc_device_t* device, const float* polygons, const uint32_t* vertex_counts, float* out, uint32_t count);
/* out_polygon: count x 34 floats {vertex_count, vertices[8][2], ellipses[8][2], inner_ellipse_0[2], sectors[8], total... see kat layout in api.cu}
 * samples: for each polygon one direction from (u0, u1) = randoms[2 i], randoms[2 i + 1]. */
int risltc_cuda_kat_psa(risltc_device_t* device, const float* polygons, const uint32_t* vertex_counts, const float* randoms,
	float* out_polygons, float* out_dirs, uint32_t count, uint32_t max_polygon_vertices, uint32_t fast_atan, uint32_t biased);
int risltc_cuda_kat_noise(risltc_device_t* device, uint32_t width, uint32_t height, uint32_t frame_word, uint32_t draws, float* out);
/* out: count x 33 floats {world_to_shading[12], shading_to_cosine[9], cosine_to_shading[9], albedo, determinant, pad} */
int risltc_cuda_kat_ltc_coefficients(risltc_device_t* device, const float* inputs /* count x 11: fresnel, roughness, pos, normal, outgoing */,
	const float ltc_constants[6], float* out, uint32_t count);
int risltc_cuda_kat_any_hit(risltc_device_t* device, const float* rays /* count x 8: o, tmin, d, tmax */, uint32_t* hits, uint32_t count);
/* The shadow-ray kernels of the frame path on the same ray array (t_min is their fixed 1e-3): kind 8 = 4-wide quantised
 * tree traversed once per pair of rays (default of render_frames; ray i and ray i + count / 2 form a pair and must share
 * their origin, count even), kind 4 = the same tree, one ray per lane, kind 2 = binary tree. */
int risltc_cuda_kat_trace(risltc_device_t* device, const float* rays, uint32_t* hits, uint32_t count, uint32_t kind);
/* Exhaustive device-side check of the kernels' hand-written exactly rounded sequences against the IEEE operations they
 * stand for: [0] inversesqrt vs 1 / sqrt over every float of its fast range, [1] unorm16 vs x / 65535 for 0..65535,
 * [2] inversesqrt over every other bit pattern. All three must be 0. */
int risltc_cuda_kat_exact_math(risltc_device_t* device, uint64_t mismatches[3]);
/* Host-only (no device): builds the acceleration structures of upload_scene for `triangle_count` triangles (9 floats each)
 * and checks their invariants. report = {triangles not in exactly one leaf slot, vertices outside their (padded) leaf box,
 * violations of the 4-wide tree (a leaf / node not referenced exactly once, a quantised box that does not contain the
 * binary tree's box), binary nodes, 4-wide nodes, depth << 32 | children per 4-wide node x 100}. The first three must be 0. */
int risltc_cuda_check_bvh(const float* vertices, uint64_t triangle_count, uint32_t max_leaf, uint64_t report[6]);
/* The same report for the acceleration structures upload_scene left on the device, whichever builder made them; report[0]
 * additionally counts triangle records that are not bit-identical to {v0, v1 - v0, v2 - v0, id} of their triangle. */
int risltc_cuda_check_scene_bvh(risltc_device_t* device, uint64_t report[6]);
/* Host-only: a hash of everything the host builder produces for these triangles (binary tree, triangle order, 4-wide tree).
 * The builder splits the top levels of large scenes over tasks; RISLTC_BVH_THREADS=1 keeps it on the calling thread -- the
 * trees, and this hash, are the same. */
int risltc_cuda_bvh_checksum(const float* vertices, uint64_t triangle_count, uint32_t max_leaf, uint64_t* checksum);

#ifdef __cplusplus
}
#endif
#endif
