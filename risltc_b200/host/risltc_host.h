/* risltc_host.h -- C99 host layer of the B200 build: the reference's data model and entry
 * points for the offline frame path, without Vulkan. Types that are plain data keep the
 * reference's layout byte for byte (polygonal_light_t 184 B, first_person_camera_t 48 B,
 * ltc_constants_t 32 B, per_frame_constants_t 256 B, render_settings_t, experiment_t); types
 * that embedded Vulkan handles keep their field names with the shim types of vk_shim.h.
 * Entry points keep the reference's names, arguments and error behaviour (0 = success,
 * non-zero after a printed message and after cleaning up; destroy_* zeroes the object).
 *
 * One header instead of the reference's six (polygonal_light.h, camera.h, noise_table.h,
 * ltc_table.h, scene.h, main.h); each section cites the file:line it mirrors. */
#ifndef RISLTC_HOST_H
#define RISLTC_HOST_H
#include "vk_shim.h"
#include <stdio.h>

#ifdef __cplusplus
#define EXTERN_C extern "C"
#else
#define EXTERN_C
#endif

#define M_PI_F 3.1415926535897932384626433832795f

/* ---- polygonal_light.h:29-136 ------------------------------------------------ */
typedef enum sample_polygon_technique_e {
	sample_polygon_baseline,
	sample_polygon_area_turk,
	sample_polygon_projected_solid_angle,
	sample_polygon_projected_solid_angle_biased,
	sample_polygon_ltc_cp,
	sample_polygon_count
} sample_polygon_technique_t;

typedef enum polygon_texturing_technique_e {
	polygon_texturing_none = 0,
	polygon_texturing_area = 1,
	polygon_texturing_portal = 2,
	polygon_texturing_ies_profile = 3,
	polygon_texturing_count,
	polygon_texturing_force_int = 0x7fffffff
} polygon_texturing_technique_t;

/* polygonal_light.h:73-99. The first 88 bytes are what a quicksave stores. */
typedef struct polygonal_light_s {
	float rotation_angles[3];
	float scaling_x;
	float translation[3];
	float scaling_y;
	float radiant_flux[3];
	float inv_scaling_x;
	float surface_radiance[3];   /* written by update_polygonal_light() */
	float inv_scaling_y;
	float plane[4];
	uint32_t vertex_count;
	polygon_texturing_technique_t texturing_technique;
	uint32_t texture_index;
	uint32_t padding_0;
	float rotation[3][4];
	float area, rcp_area;
	float padding_1[2];
	char* texture_file_path;
	float* vertices_plane_space;   /* vertex i at [4 i], [4 i + 1] */
	float* vertices_world_space;   /* vertex i at [4 i .. 4 i + 2], written by update_polygonal_light() */
} polygonal_light_t;

/* polygonal_light.h:104-111: header of one record in the light buffer */
typedef struct polygonal_light_upload_s {
	float surface_radiance[3];
	float padding_1;
	float plane[4];
	uint32_t vertex_count;
	float padding_2[3];
} polygonal_light_upload_t;

#define POLYGONAL_LIGHT_QUICKSAVE_SIZE (sizeof(float) * 20 + sizeof(uint32_t) * 2)
#define POLYGONAL_LIGHT_FIXED_CONSTANT_BUFFER_SIZE (sizeof(float) * 12)

EXTERN_C int set_polygonal_light_vertex_count(polygonal_light_t* light, uint32_t vertex_count);
EXTERN_C void update_polygonal_light(polygonal_light_t* light);
EXTERN_C polygonal_light_t duplicate_polygonal_light(const polygonal_light_t* light);
EXTERN_C void destroy_polygonal_light(polygonal_light_t* light);

/* ---- camera.h:29-61 ---------------------------------------------------------- */
typedef struct first_person_camera_s {
	float position_world_space[3];
	float rotation_z;
	float rotation_x;
	float vertical_fov;
	float near, far;
	float speed;
	int rotate_camera;
	float rotation_x_0, rotation_z_0;
} first_person_camera_t;

EXTERN_C void get_world_to_view_space(float world_to_view_space[4][4], const first_person_camera_t* camera);
EXTERN_C void get_view_to_projection_space(float view_to_projection_space[4][4], const first_person_camera_t* camera, float aspect_ratio);
EXTERN_C void get_world_to_projection_space(float world_to_projection_space[4][4], const first_person_camera_t* camera, float aspect_ratio);

/* ---- math_utilities.h:24-57 -------------------------------------------------- */
EXTERN_C void matrix_inverse(float inverse[4][4], const float matrix[4][4]);
EXTERN_C uint32_t wang_random_number(uint32_t seed);

/* ---- noise_table.h:25-35 ----------------------------------------------------- */
typedef struct noise_table_s {
	uint32_t random_seed;
} noise_table_t;

EXTERN_C void set_noise_constants(uint32_t resolution_mask[2], uint32_t* texture_index_mask, uint32_t random_numbers[4], noise_table_t* noise, VkBool32 animate_noise);

/* ---- ltc_table.h:23-72 ------------------------------------------------------- */
typedef struct ltc_constants_s {
	float fresnel_index_factor, fresnel_index_summand;
	float roughness_factor, roughness_summand;
	float inclination_factor, inclination_summand;
	float padding[2];
} ltc_constants_t;

typedef struct ltc_table_s {
	uint32_t roughness_count, inclination_count, fresnel_count;
	/* images[0]: RGBA16 UNORM array, images[1]: RG16 UNORM array (host copies + device residency) */
	images_t texture_arrays;
	VkSampler sampler;
	ltc_constants_t constants;
} ltc_table_t;

EXTERN_C int load_ltc_table(ltc_table_t* table, const device_t* device, const char* directory, uint32_t fresnel_count);
EXTERN_C void destroy_ltc_table(ltc_table_t* table, const device_t* device);

/* ---- scene.h:30-184 ---------------------------------------------------------- */
typedef enum mesh_buffer_type_e {
	mesh_buffer_type_positions,
	mesh_buffer_type_normals_and_tex_coords,
	mesh_buffer_type_material_indices,
	mesh_buffer_count,
	mesh_buffer_type_triangle = mesh_buffer_count,
	mesh_buffer_count_full
} mesh_buffer_type_t;

typedef struct mesh_s {
	uint64_t triangle_count;
	float dequantization_factor[3], dequantization_summand[3];
	union {
		struct {
			buffer_t positions;
			buffer_t normals_and_tex_coords;
			buffer_t material_indices;
			buffer_t triangle;
		};
		buffer_t buffers[mesh_buffer_count_full];
	};
	union {
		struct {
			VkBufferView positions_view;
			VkBufferView normals_and_tex_coords_view;
			VkBufferView material_indices_view;
			VkBufferView triangle_view;
		};
		VkBufferView buffer_views[mesh_buffer_count_full];
	};
	/* Host staging copy of the three mesh buffers (one malloc, offsets in buffers[]) */
	VkDeviceMemory memory;
	VkDeviceSize size;
} mesh_t;

typedef enum material_texture_type_e {
	material_texture_type_base_color,
	material_texture_type_specular,
	material_texture_type_normal,
	material_texture_count
} material_texture_type_t;

typedef struct materials_s {
	uint64_t material_count;
	char** material_names;
	images_t textures;
	VkSampler sampler;
} materials_t;

typedef struct acceleration_structure_s {
	union {
		struct {
			VkAccelerationStructureKHR bottom_level;
			VkAccelerationStructureKHR top_level;
		};
		VkAccelerationStructureKHR levels[2];
	};
	buffers_t buffers;
} acceleration_structure_t;

typedef struct scene_s {
	mesh_t mesh;
	materials_t materials;
	acceleration_structure_t acceleration_structure;
} scene_t;

EXTERN_C const char* get_material_texture_suffix(material_texture_type_t type);
EXTERN_C int load_scene(scene_t* scene, const device_t* device, const char* file_path, const char* texture_path, VkBool32 request_acceleration_structure);
EXTERN_C void destroy_scene(scene_t* scene, const device_t* device);

/* ---- main.h:30-241 ------------------------------------------------------------ */
typedef struct scene_specification_s {
	char* file_path;
	char* texture_path;
	char* quick_save_path;
	first_person_camera_t camera;
	uint32_t polygonal_light_count;
	polygonal_light_t* polygonal_lights;
} scene_specification_t;

typedef enum sampling_strategies_e {
	sampling_strategies_diffuse_only,
	sampling_strategies_diffuse_specular_mis,
	sampling_strategies_count
} sampling_strategies_t;

typedef enum mis_heuristic_e {
	mis_heuristic_balance,
	mis_heuristic_power,
	mis_heuristic_weighted,
	mis_heuristic_optimal_clamped,
	mis_heuristic_optimal,
	mis_heuristic_count
} mis_heuristic_t;

typedef enum light_sampling_strategies_e {
	light_uniform,
	light_reservoir
} light_sampling_strategies_t;

typedef enum error_display_e {
	error_display_none,
	error_display_diffuse_backward,
	error_display_diffuse_backward_scaled,
	error_display_diffuse_forward,
	error_display_specular_backward,
	error_display_specular_backward_scaled,
	error_display_specular_forward,
	error_display_count
} error_display_t;

typedef enum bool_override_e {
	bool_override_false = 0,
	bool_override_true = 1,
	bool_override_none = 2,
} bool_override_t;

typedef struct render_settings_s {
	float exposure_factor, roughness_factor;
	uint32_t sample_count;
	uint32_t sample_count_light;
	mis_heuristic_t mis_heuristic;
	light_sampling_strategies_t light_sampling;
	float mis_visibility_estimate;
	sample_polygon_technique_t polygon_sampling_technique;
	error_display_t error_display;
	float error_min_exponent;
	VkBool32 animate_noise;
	VkBool32 accum;
	VkBool32 show_polygonal_lights;
	VkBool32 show_gui;
	VkBool32 v_sync;
	VkBool32 fast_atan;
} render_settings_t;

typedef enum scene_index_e {
	scene_bistro_inside,
	scene_bistro_outside,
	scene_zeroday,
	scene_count
} scene_index_t;

/* Display name, .vks path, texture directory, quicksave path (main.c:36-40). Paths may be
 * redirected to generated data with the RISLTC_DATA_DIR environment variable. */
extern const char* const g_scene_paths[scene_count][4];

typedef struct experiment_s {
	uint32_t width, height;
	scene_index_t scene_index;
	char* quick_save_path;
	VkBool32 use_hdr;
	char* screenshot_path;
	uint32_t num_samples;
	render_settings_t render_settings;
	char* base_dir;
	char* timings_path;
	char* screenshots_dir;
	char* ext;
	char* exp_name;
	VkBool32 ss_per_frame;
} experiment_t;

typedef enum experiment_state_e {
	experiment_state_rendering,
	experiment_state_screenshot_frame_0,
	experiment_state_screenshot_frame_1,
	experiment_state_new_experiment,
} experiment_state_t;

typedef struct experiment_list_s {
	experiment_t* experiments;
	const experiment_t* experiment;
	uint32_t count;
	uint32_t next;
	uint32_t next_setup_frame;
	experiment_state_t state;
	FILE* timings_file;
} experiment_list_t;

/* main.h:537-553, mirror of shared_constants.glsl:21-60 (std140, row_major), 256 bytes */
typedef struct per_frame_constants_s {
	float mesh_dequantization_factor[3], padding_0, mesh_dequantization_summand[3];
	float error_factor;
	float world_to_projection_space[4][4];
	float pixel_to_ray_direction_world_space[3][4];
	float camera_position_world_space[3];
	float mis_visibility_estimate;
	VkExtent2D viewport_size;
	int32_t cursor_position[2];
	float exposure_factor;
	float roughness_factor;
	uint32_t noise_resolution_mask[2];
	uint32_t noise_texture_index_mask;
	uint32_t padding_3[3];
	uint32_t noise_random_numbers[4];
	ltc_constants_t ltc_constants;
} per_frame_constants_t;

/* main.h:556-562, experiment_list.c:34,473 */
EXTERN_C void create_experiment_list(experiment_list_t* list);
EXTERN_C void destroy_experiment_list(experiment_list_t* list);

/* ---- application (main.h:499-529 reduced to the offline frame path) ------------- */
typedef struct application_updates_s {
	VkBool32 startup, recreate_swapchain, reload_shaders, update_light_count, update_light_textures;
	VkBool32 reload_scene, change_shading, quick_save, quick_load;
} application_updates_t;

typedef struct swapchain_s {
	VkExtent2D extent;   /* the render resolution; there is no window */
} swapchain_t;

typedef struct application_s {
	device_t device;
	swapchain_t swapchain;
	scene_specification_t scene_specification;
	render_settings_t render_settings;
	scene_t scene;
	noise_table_t noise_table;
	ltc_table_t ltc_table;
	experiment_list_t experiment_list;
	bool_override_t run_all_exp;
	uint32_t accum_num;
	FILE* timings;
	/* image partition of this process (multi-GPU: interleaved stripes) */
	uint32_t stripe_height, stripe_index, stripe_count;
	/* milliseconds of the most recent frame, as record_frame_time reports them (frame_timer.c:37-55) */
	float last_frame_ms;
	/* a screenshot of the next rendered frame was requested under this sample index (main.c:2757-2760) */
	VkBool32 screenshot_pending;
	uint32_t screenshot_index;
} application_t;

EXTERN_C void quick_save(scene_specification_t* scene);
EXTERN_C void quick_load(scene_specification_t* scene, application_updates_t* updates);
EXTERN_C void specify_default_scene(scene_specification_t* scene);
EXTERN_C void specify_default_render_settings(render_settings_t* settings);
EXTERN_C void destroy_scene_specification(scene_specification_t* scene);
EXTERN_C uint32_t get_min_polygonal_light_vertex_count(const scene_specification_t* scene_specification);
EXTERN_C uint32_t get_max_polygonal_light_vertex_count(const scene_specification_t* scene_specification);
EXTERN_C uint32_t get_max_polygon_vertex_count(const scene_specification_t* scene_specification, const render_settings_t* render_settings);
/* main.c:2902-2946 and :456-490. write_lights needs get_max_polygonal_light_vertex_count(...) * 16 + 48 bytes per light. */
EXTERN_C void write_constants(void* data, application_t* app);
EXTERN_C void write_lights(void* data, application_t* app);
EXTERN_C size_t get_light_buffer_size(const scene_specification_t* scene_specification);
/* main.c:2569 (argv handling main.c:3050-3058), :2467, :2955, :2719, :2647, main loop :3067-3074 */
EXTERN_C int startup_application(application_t* app, int experiment_index, bool_override_t run_all_exp, int cuda_ordinal);
EXTERN_C int update_application(application_t* app, const application_updates_t* updates);
EXTERN_C int render_frame(application_t* app);
EXTERN_C int setup_experiment(application_t* app, const experiment_t* experiment);
EXTERN_C int advance_experiments(application_t* app);
EXTERN_C void destroy_application(application_t* app);
/* The accumulated frame as RGBA32F (rows owned by this process, see stripe_*). */
EXTERN_C int read_accumulation_buffer(application_t* app, float* rgba);
/* implement_screenshot (main.c:2358-2409, screenshot.c): the accumulated frame through the copy pass to *.png (8-bit sRGB) and /
 * or *.hdr (two frames of half-float bits combined, Radiance RGBE); either path may be NULL */
EXTERN_C int take_screenshot(application_t* app, const char* path_png, const char* path_hdr);
EXTERN_C void combine_ldr_screenshots_into_hdr(float* hdr, const uint8_t* low_bits, const uint8_t* high_bits, size_t entry_count);
EXTERN_C float half_to_float(uint16_t half_bits);
EXTERN_C uint16_t float_to_half(float value);
EXTERN_C int write_png(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height);
EXTERN_C int write_hdr(const char* path, const float* rgb, uint32_t width, uint32_t height);
/* an RGBA32F frame to *.hdr; the values go through fp16 like the copy pass's two half-bit frames do */
EXTERN_C int write_hdr_screenshot(const char* path, const float* rgba, uint32_t width, uint32_t height);
/* -run_exp / -e<N> command line of the reference's main() (main.c:3044-3078) */
EXTERN_C int risltc_main(int argc, char** argv);

#endif
