/* application.c -- the offline frame path of the reference's main.c on top of the C ABI:
 * quicksaves (main.c:45-125), defaults (:129-236), per-frame constants (:2902-2946), the light
 * buffer (:456-490), start-up / update / one frame (:2569, :2467, :2955), the experiment state
 * machine (:2647-2790); screenshots (:2339-2409) are in screenshot.c. No window, no swapchain, no GUI. */
#include "risltc_host.h"
#include "risltc_cuda.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>

const char* const g_scene_paths[scene_count][4] = {
	{ "Bistro Interior", "data/Bistro_interior.vks", "data/Bistro_textures", "data/quicksaves/Bistro_interior.save" },
	{ "Bistro Exterior", "data/Bistro_exterior.vks", "data/Bistro_textures", "data/quicksaves/Bistro_exterior.save" },
	{ "Zero Day", "data/zeroday.vks", "data/ZeroDay_textures", "data/quicksaves/ZeroDay.save" },
};

static char* dup_string(const char* s) {
	if (!s) return NULL;
	size_t n = strlen(s) + 1;
	char* r = (char*) malloc(n);
	memcpy(r, s, n);
	return r;
}

/* Paths in the tables above are relative to the working directory like in the reference; when
 * RISLTC_DATA_DIR is set, a leading "data/" is replaced by that directory so that generated
 * inputs can live anywhere. The result is malloc'ed. */
static char* resolve_path(const char* path) {
	const char* root = getenv("RISLTC_DATA_DIR");
	if (root && strncmp(path, "data/", 5) == 0) {
		size_t n = strlen(root) + strlen(path + 4) + 1;
		char* r = (char*) malloc(n);
		snprintf(r, n, "%s%s", root, path + 4);
		return r;
	}
	return dup_string(path);
}

/* ------------------------------------------------------------------ quicksaves */

void quick_save(scene_specification_t* scene) {
	char* path = resolve_path(scene->quick_save_path);
	FILE* file = fopen(path, "wb");
	if (!file) {
		printf("Quick save failed. Please check path and permissions: %s\n", scene->quick_save_path);
		free(path);
		return;
	}
	free(path);
	fwrite(&scene->camera, sizeof(scene->camera), 1, file);
	const uint32_t legacy_count = 0;
	fwrite(&legacy_count, sizeof(uint32_t), 1, file);
	fwrite(&scene->polygonal_light_count, sizeof(uint32_t), 1, file);
	for (uint32_t i = 0; i != scene->polygonal_light_count; ++i) {
		const polygonal_light_t* light = &scene->polygonal_lights[i];
		fwrite(light, POLYGONAL_LIGHT_QUICKSAVE_SIZE, 1, file);
		size_t path_size = light->texture_file_path ? strlen(light->texture_file_path) + 1 : 0;
		fwrite(&path_size, sizeof(path_size), 1, file);
		if (path_size) fwrite(light->texture_file_path, sizeof(char), path_size, file);
		const float* null_pointers[2] = { NULL, NULL };
		fwrite(null_pointers, sizeof(float*), 2, file);
		fwrite(light->vertices_plane_space, sizeof(float), 4 * light->vertex_count, file);
	}
	fclose(file);
}

void quick_load(scene_specification_t* scene, application_updates_t* updates) {
	char* path = resolve_path(scene->quick_save_path);
	FILE* file = fopen(path, "rb");
	free(path);
	if (!file) {
		printf("Failed to load a quick save. Please check path and permissions: %s\n", scene->quick_save_path);
		return;
	}
	size_t got = fread(&scene->camera, sizeof(scene->camera), 1, file);
	uint32_t legacy_count = 0;
	got += fread(&legacy_count, sizeof(uint32_t), 1, file);
	const uint32_t old_count = scene->polygonal_light_count;
	polygonal_light_t* old_lights = scene->polygonal_lights;
	scene->polygonal_light_count = 0;
	got += fread(&scene->polygonal_light_count, sizeof(uint32_t), 1, file);
	VkBool32 vertex_count_changed = VK_FALSE;
	scene->polygonal_lights = (polygonal_light_t*) calloc(scene->polygonal_light_count ? scene->polygonal_light_count : 1, sizeof(polygonal_light_t));
	for (uint32_t i = 0; i != scene->polygonal_light_count; ++i) {
		polygonal_light_t* light = &scene->polygonal_lights[i];
		got += fread(light, POLYGONAL_LIGHT_QUICKSAVE_SIZE, 1, file);
		if (i < old_count && light->vertex_count != old_lights[i].vertex_count) vertex_count_changed = VK_TRUE;
		if (light->scaling_y <= 0.0f) light->scaling_y = light->scaling_x;   /* legacy files */
		size_t path_size = 0;
		got += fread(&path_size, sizeof(path_size), 1, file);
		light->texture_file_path = NULL;
		if (path_size) {
			light->texture_file_path = (char*) malloc(path_size);
			got += fread(light->texture_file_path, sizeof(char), path_size, file);
			if (updates && i < old_count && old_lights[i].texture_file_path && strcmp(light->texture_file_path, old_lights[i].texture_file_path) != 0)
				updates->update_light_textures = VK_TRUE;
		}
		float* stored_pointers[2];
		got += fread(stored_pointers, sizeof(float*), 2, file);
		light->vertices_plane_space = NULL;
		light->vertices_world_space = NULL;
		uint32_t vertex_count = light->vertex_count;
		light->vertex_count = 0;
		set_polygonal_light_vertex_count(light, vertex_count);
		got += fread(light->vertices_plane_space, sizeof(float), 4 * light->vertex_count, file);
	}
	(void) got;
	for (uint32_t i = 0; i != old_count; ++i) destroy_polygonal_light(&old_lights[i]);
	free(old_lights);
	fclose(file);
	if (updates)
		updates->update_light_count |= (old_count != scene->polygonal_light_count) || vertex_count_changed;
}

/* -------------------------------------------------------------------- defaults */

void specify_default_scene(scene_specification_t* scene) {
	const uint32_t scene_index = scene_zeroday;
	memset(scene, 0, sizeof(*scene));
	scene->file_path = dup_string(g_scene_paths[scene_index][1]);
	scene->texture_path = dup_string(g_scene_paths[scene_index][2]);
	scene->quick_save_path = dup_string(g_scene_paths[scene_index][3]);
	first_person_camera_t camera;
	memset(&camera, 0, sizeof(camera));
	camera.near = 0.05f; camera.far = 1.0e3f;
	camera.vertical_fov = 0.33f * M_PI_F;
	camera.rotation_x = 0.43f * M_PI_F;
	camera.rotation_z = 1.3f * M_PI_F;
	camera.position_world_space[0] = -3.0f; camera.position_world_space[1] = -2.0f; camera.position_world_space[2] = 1.65f;
	camera.speed = 2.0f;
	scene->camera = camera;
	/* one unit square light, replaced by whatever the quicksave holds */
	polygonal_light_t light;
	memset(&light, 0, sizeof(light));
	light.rotation_angles[0] = 0.5f * M_PI_F;
	light.scaling_x = light.scaling_y = 1.0f;
	light.radiant_flux[0] = light.radiant_flux[1] = light.radiant_flux[2] = 1.0f;
	set_polygonal_light_vertex_count(&light, 4);
	const float corners[4][2] = { { 0.0f, 0.0f }, { 1.0f, 0.0f }, { 1.0f, 1.0f }, { 0.0f, 1.0f } };
	for (uint32_t i = 0; i != 4; ++i) { light.vertices_plane_space[4 * i] = corners[i][0]; light.vertices_plane_space[4 * i + 1] = corners[i][1]; }
	scene->polygonal_light_count = 1;
	scene->polygonal_lights = (polygonal_light_t*) malloc(sizeof(light));
	scene->polygonal_lights[0] = light;
	quick_load(scene, NULL);
}

uint32_t get_min_polygonal_light_vertex_count(const scene_specification_t* s) {
	if (!s->polygonal_light_count) return 3;
	uint32_t minimum = 0x7FFFFFFF;
	for (uint32_t i = 0; i != s->polygonal_light_count; ++i)
		if (minimum > s->polygonal_lights[i].vertex_count) minimum = s->polygonal_lights[i].vertex_count;
	return minimum;
}

uint32_t get_max_polygonal_light_vertex_count(const scene_specification_t* s) {
	uint32_t maximum = 3;
	for (uint32_t i = 0; i != s->polygonal_light_count; ++i)
		if (maximum < s->polygonal_lights[i].vertex_count) maximum = s->polygonal_lights[i].vertex_count;
	return maximum;
}

uint32_t get_max_polygon_vertex_count(const scene_specification_t* s, const render_settings_t* settings) {
	uint32_t v = get_max_polygonal_light_vertex_count(s);
	switch (settings->polygon_sampling_technique) {
	case sample_polygon_projected_solid_angle:
	case sample_polygon_projected_solid_angle_biased:
	case sample_polygon_ltc_cp:
		return v + 1;   /* clipping may add a vertex */
	default:
		return v;
	}
}

void destroy_scene_specification(scene_specification_t* scene) {
	free(scene->file_path); free(scene->texture_path); free(scene->quick_save_path);
	for (uint32_t i = 0; i != scene->polygonal_light_count; ++i) destroy_polygonal_light(&scene->polygonal_lights[i]);
	free(scene->polygonal_lights);
	memset(scene, 0, sizeof(*scene));
}

void specify_default_render_settings(render_settings_t* settings) {
	settings->exposure_factor = 1.5f;
	settings->roughness_factor = 1.0f;
	settings->sample_count = 1;
	settings->sample_count_light = 1;
	settings->mis_heuristic = mis_heuristic_optimal_clamped;
	settings->mis_visibility_estimate = 0.5f;
	settings->polygon_sampling_technique = sample_polygon_ltc_cp;
	settings->light_sampling = light_reservoir;
	settings->error_display = error_display_none;
	settings->error_min_exponent = -7.0f;
	settings->accum = VK_FALSE;
	settings->show_polygonal_lights = VK_FALSE;
	settings->animate_noise = VK_TRUE;
	settings->v_sync = VK_FALSE;
	settings->show_gui = VK_TRUE;
}

/* ------------------------------------------------- constants and light buffer */

void write_constants(void* data, application_t* app) {
	const scene_t* scene = &app->scene;
	const first_person_camera_t* camera = &app->scene_specification.camera;
	per_frame_constants_t constants;
	memset(&constants, 0, sizeof(constants));
	for (uint32_t i = 0; i != 3; ++i) {
		constants.mesh_dequantization_factor[i] = scene->mesh.dequantization_factor[i];
		constants.mesh_dequantization_summand[i] = scene->mesh.dequantization_summand[i];
		constants.camera_position_world_space[i] = camera->position_world_space[i];
	}
	constants.mis_visibility_estimate = app->render_settings.mis_visibility_estimate;
	constants.viewport_size = app->swapchain.extent;
	constants.ltc_constants = app->ltc_table.constants;
	constants.error_factor = powf(10.0f, -app->render_settings.error_min_exponent);
	constants.exposure_factor = app->render_settings.exposure_factor;
	constants.roughness_factor = app->render_settings.roughness_factor;
	set_noise_constants(constants.noise_resolution_mask, &constants.noise_texture_index_mask, constants.noise_random_numbers, &app->noise_table, app->render_settings.animate_noise);
	const float aspect_ratio = ((float) app->swapchain.extent.width) / ((float) app->swapchain.extent.height);
	get_world_to_projection_space(constants.world_to_projection_space, camera, aspect_ratio);
	/* pixel index -> world-space ray direction through the pixel centre */
	float viewport[4];
	viewport[0] = 2.0f / app->swapchain.extent.width;
	viewport[1] = 2.0f / app->swapchain.extent.height;
	viewport[2] = 0.5f * viewport[0] - 1.0f;
	viewport[3] = 0.5f * viewport[1] - 1.0f;
	float rotation_only[4][4], inverse[4][4];
	memcpy(rotation_only, constants.world_to_projection_space, sizeof(rotation_only));
	rotation_only[0][3] = rotation_only[1][3] = rotation_only[2][3] = 0.0f;
	matrix_inverse(inverse, rotation_only);
	const float pixel_to_projection[4][3] = {
		{ viewport[0], 0.0f, viewport[2] },
		{ 0.0f, viewport[1], viewport[3] },
		{ 0.0f, 0.0f, 1.0f },
		{ 0.0f, 0.0f, 1.0f } };
	for (uint32_t i = 0; i != 3; ++i)
		for (uint32_t j = 0; j != 3; ++j)
			for (uint32_t k = 0; k != 4; ++k)
				constants.pixel_to_ray_direction_world_space[i][j] += inverse[i][k] * pixel_to_projection[k][j];
	memcpy(data, &constants, sizeof(constants));
}

size_t get_light_buffer_size(const scene_specification_t* s) {
	return (size_t) s->polygonal_light_count * (POLYGONAL_LIGHT_FIXED_CONSTANT_BUFFER_SIZE + sizeof(float) * 4 * get_max_polygonal_light_vertex_count(s));
}

void write_lights(void* data, application_t* app) {
	scene_specification_t* s = &app->scene_specification;
	const uint32_t max_vertex_count = get_max_polygonal_light_vertex_count(s);
	char* cursor = (char*) data;
	for (uint32_t i = 0; i != s->polygonal_light_count; ++i) {
		polygonal_light_t* light = &s->polygonal_lights[i];
		update_polygonal_light(light);
		polygonal_light_upload_t head;
		memset(&head, 0, sizeof(head));
		memcpy(head.surface_radiance, light->surface_radiance, sizeof(head.surface_radiance));
		memcpy(head.plane, light->plane, sizeof(head.plane));
		head.vertex_count = light->vertex_count;
		memcpy(cursor, &head, POLYGONAL_LIGHT_FIXED_CONSTANT_BUFFER_SIZE);
		cursor += POLYGONAL_LIGHT_FIXED_CONSTANT_BUFFER_SIZE;
		memset(cursor, 0, sizeof(float) * 4 * max_vertex_count);
		memcpy(cursor, light->vertices_world_space, sizeof(float) * 4 * light->vertex_count);
		if (light->vertex_count < max_vertex_count)
			memcpy(cursor + sizeof(float) * 4 * light->vertex_count, light->vertices_world_space, sizeof(float) * 4);
		cursor += sizeof(float) * 4 * max_vertex_count;
	}
}

/* ------------------------------------------------------------ application */

static int upload_lights_and_variant(application_t* app) {
	scene_specification_t* s = &app->scene_specification;
	if (s->polygonal_light_count == 0) {
		printf("The scene specification holds no polygonal lights.\n");
		return 1;
	}
	size_t size = get_light_buffer_size(s);
	void* records = malloc(size);
	write_lights(records, app);
	int result = risltc_cuda_upload_lights(app->device.cuda, records, s->polygonal_light_count, get_max_polygonal_light_vertex_count(s));
	free(records);
	if (result) return 1;
	const render_settings_t* r = &app->render_settings;
	risltc_variant_t variant = {
		(uint32_t) r->light_sampling, (uint32_t) r->polygon_sampling_technique, (uint32_t) r->mis_heuristic,
		r->sample_count, r->sample_count_light, r->fast_atan,
		get_min_polygonal_light_vertex_count(s), get_max_polygonal_light_vertex_count(s) };
	return risltc_cuda_set_variant(app->device.cuda, &variant);
}

int update_application(application_t* app, const application_updates_t* updates) {
	if (updates->startup || updates->reload_scene) {
		destroy_scene(&app->scene, &app->device);
		char* file_path = resolve_path(app->scene_specification.file_path);
		char* texture_path = resolve_path(app->scene_specification.texture_path);
		int result = load_scene(&app->scene, &app->device, file_path, texture_path, VK_TRUE);
		free(file_path); free(texture_path);
		if (result) return 1;
	}
	if (updates->quick_load) quick_load(&app->scene_specification, NULL);
	if (updates->startup || updates->recreate_swapchain) {
		if (risltc_cuda_resize(app->device.cuda, app->swapchain.extent.width, app->swapchain.extent.height, app->stripe_height, app->stripe_index, app->stripe_count)) return 1;
		app->accum_num = 0;
	}
	if (updates->startup || updates->quick_load || updates->update_light_count || updates->change_shading || updates->reload_scene) {
		if (upload_lights_and_variant(app)) return 1;
	}
	if (updates->quick_save) quick_save(&app->scene_specification);
	return 0;
}

void destroy_application(application_t* app) {
	if (app->timings) fclose(app->timings);
	destroy_experiment_list(&app->experiment_list);
	destroy_scene(&app->scene, &app->device);
	destroy_ltc_table(&app->ltc_table, &app->device);
	destroy_scene_specification(&app->scene_specification);
	if (app->device.cuda) risltc_cuda_destroy_device(app->device.cuda);
	memset(app, 0, sizeof(*app));
}

int startup_application(application_t* app, int experiment_index, bool_override_t run_all_exp, int cuda_ordinal) {
	uint32_t stripes[3] = { app->stripe_height, app->stripe_index, app->stripe_count };
	memset(app, 0, sizeof(*app));
	app->stripe_height = stripes[0] ? stripes[0] : 8;
	app->stripe_index = stripes[1];
	app->stripe_count = stripes[2] ? stripes[2] : 1;
	app->run_all_exp = run_all_exp;
	if (risltc_cuda_create_device(&app->device.cuda, cuda_ordinal)) {
		printf("Failed to create a CUDA device object: %s\n", risltc_cuda_last_error());
		return 1;
	}
	app->device.cuda_ordinal = cuda_ordinal;
	app->device.ray_tracing_supported = VK_TRUE;
	create_experiment_list(&app->experiment_list);
	if (run_all_exp == bool_override_true) {
		app->experiment_list.next = 0;
		app->experiment_list.state = experiment_state_new_experiment;
	}
	else if (experiment_index >= 0 && (uint32_t) experiment_index < app->experiment_list.count) {
		app->experiment_list.next = (uint32_t) experiment_index;
		app->experiment_list.state = experiment_state_new_experiment;
	}
	specify_default_scene(&app->scene_specification);
	specify_default_render_settings(&app->render_settings);
	app->swapchain.extent.width = 1280; app->swapchain.extent.height = 1024;
	char* ltc_directory = resolve_path("data/ggx_ltc_fit");
	int result = load_ltc_table(&app->ltc_table, &app->device, ltc_directory, 51);
	free(ltc_directory);
	if (result) { destroy_application(app); return 1; }
	return 0;
}

static int take_experiment_screenshot(application_t* app, uint32_t index);

int render_frame(application_t* app) {
	per_frame_constants_t constants;
	write_constants(&constants, app);
	if (risltc_cuda_render_frame(app->device.cuda, &constants, app->accum_num)) return 1;
	app->last_frame_ms = risltc_cuda_last_frame_ms(app->device.cuda);
	if (app->timings) fprintf(app->timings, "%i,%f\n", (int) app->accum_num, app->last_frame_ms);
	/* implement_screenshot (main.c:2358): a requested screenshot captures the frame just rendered */
	if (app->screenshot_pending) {
		app->screenshot_pending = VK_FALSE;
		if (take_experiment_screenshot(app, app->screenshot_index)) return 1;
	}
	if (app->render_settings.accum) {
		++app->accum_num;
		if (app->accum_num % 1000 == 0) printf("%d Samples Completed\n", (int) app->accum_num);
	}
	return 0;
}

int read_accumulation_buffer(application_t* app, float* rgba) {
	return risltc_cuda_read_accum(app->device.cuda, rgba);
}

/* --------------------------------------------------------------- screenshots */

static int take_experiment_screenshot(application_t* app, uint32_t index) {
	const experiment_t* e = app->experiment_list.experiment;
	if (!e) return 0;
	if (app->stripe_count != 1) return 0;   /* partial frames are gathered by the multi-GPU driver */
	char name[32];
	snprintf(name, sizeof(name), "/%05d", (int) index);
	size_t n = strlen(e->base_dir) + strlen(e->exp_name) + strlen(name) + strlen(e->ext) + 2;
	char* relative = (char*) malloc(n);
	snprintf(relative, n, "%s%s%s.%s", e->base_dir, e->exp_name, name, e->ext);
	char* path = resolve_path(relative);
	/* experiment_t.use_hdr: *.hdr through the two half-bit frames, else *.png (main.c:2757-2760) */
	int result = e->use_hdr ? take_screenshot(app, NULL, path) : take_screenshot(app, path, NULL);
	free(path); free(relative);
	return result;
}

/* ---------------------------------------------------------------- experiments */

static void make_directories(const char* path) {
	char* copy = dup_string(path);
	for (char* p = copy + 1; *p; ++p)
		if (*p == '/') { *p = 0; mkdir(copy, 0777); *p = '/'; }
	mkdir(copy, 0777);
	free(copy);
}

int setup_experiment(application_t* app, const experiment_t* experiment) {
	experiment_list_t* list = &app->experiment_list;
	scene_specification_t* scene = &app->scene_specification;
	application_updates_t updates;
	memset(&updates, 0, sizeof(updates));
	list->experiment = experiment;
	char* directory = resolve_path(experiment->screenshots_dir);
	make_directories(directory);
	free(directory);
	list->next_setup_frame = experiment->num_samples;
	list->state = experiment_state_rendering;
	if (app->timings) { fclose(app->timings); app->timings = NULL; }
	if (!experiment->ss_per_frame) {
		char* timings_path = resolve_path(experiment->timings_path);
		app->timings = fopen(timings_path, "w");
		free(timings_path);
	}
	if (experiment->width && experiment->height && (experiment->width != app->swapchain.extent.width || experiment->height != app->swapchain.extent.height || !app->scene.mesh.triangle_count)) {
		app->swapchain.extent.width = experiment->width; app->swapchain.extent.height = experiment->height;
		updates.recreate_swapchain = VK_TRUE;
	}
	if (!app->scene.mesh.triangle_count || strcmp(scene->file_path, g_scene_paths[experiment->scene_index][1]) != 0) {
		free(scene->file_path); free(scene->texture_path);
		scene->file_path = dup_string(g_scene_paths[experiment->scene_index][1]);
		scene->texture_path = dup_string(g_scene_paths[experiment->scene_index][2]);
		updates.reload_scene = VK_TRUE;
	}
	free(scene->quick_save_path);
	scene->quick_save_path = dup_string(experiment->quick_save_path ? experiment->quick_save_path : g_scene_paths[experiment->scene_index][3]);
	updates.quick_load = VK_TRUE;
	updates.change_shading = VK_TRUE;
	app->render_settings = experiment->render_settings;
	app->accum_num = 0;
	if (!app->scene.mesh.triangle_count) updates.recreate_swapchain = VK_TRUE;
	return update_application(app, &updates);
}

/* One step of the experiment state machine (main.c:2719-2790), called before every frame.
 * Returns 1 once all experiments have finished, -1 on failure, 0 otherwise. */
int advance_experiments(application_t* app) {
	experiment_list_t* list = &app->experiment_list;
	if (list->next > list->count) return list->state == experiment_state_new_experiment;
	if (list->next == list->count) { list->next = list->count + 1; list->state = experiment_state_new_experiment; return 1; }
	if (list->state == experiment_state_new_experiment) {
		if (setup_experiment(app, &list->experiments[list->next])) return -1;
		if (risltc_cuda_resize(app->device.cuda, app->swapchain.extent.width, app->swapchain.extent.height, app->stripe_height, app->stripe_index, app->stripe_count)) return -1;
		if (upload_lights_and_variant(app)) return -1;
	}
	const experiment_t* e = list->experiment;
	int finished = 0;
	if (e->ss_per_frame) {
		/* a screenshot every 10 accumulated frames; the experiment ends with the first one at or past num_samples - 1 */
		if (list->state == experiment_state_screenshot_frame_0) {
			if (list->next_setup_frame > app->accum_num + 1) list->state = experiment_state_rendering;
			else finished = 1;
		}
		else if (app->accum_num % 10 == 0) {
			app->screenshot_pending = VK_TRUE; app->screenshot_index = app->accum_num;
			list->state = experiment_state_screenshot_frame_0;
		}
	}
	else {
		if (list->state == experiment_state_screenshot_frame_0) finished = 1;
		else if (list->next_setup_frame <= app->accum_num) {
			app->screenshot_pending = VK_TRUE; app->screenshot_index = app->accum_num;
			list->state = experiment_state_screenshot_frame_0;
		}
	}
	if (finished) {
		if (app->timings) { fclose(app->timings); app->timings = NULL; }
		list->state = experiment_state_new_experiment;
		if (list->next + 1 == list->count) { list->experiment = NULL; list->next = list->count + 1; return 1; }
		++list->next;
	}
	return 0;
}

int risltc_main(int argc, char** argv) {
	int experiment = -1, ordinal = 0;
	bool_override_t run_all_exp = bool_override_false;
	for (int i = 1; i < argc; ++i) {
		const char* arg = argv[i];
		if (arg[0] == '-' && arg[1] == 'e') sscanf(arg + 2, "%d", &experiment);
		if (strcmp(arg, "-run_exp") == 0) run_all_exp = bool_override_true;
		if (strncmp(arg, "-gpu", 4) == 0) sscanf(arg + 4, "%d", &ordinal);
		/* -no_v_sync, -v_sync, -no_gui, -gui are accepted and ignored: there is no window */
	}
	application_t app;
	memset(&app, 0, sizeof(app));
	if (startup_application(&app, experiment, run_all_exp, ordinal)) {
		printf("Application startup has failed.\n");
		return 1;
	}
	if (app.experiment_list.next > app.experiment_list.count) {
		printf("No experiment selected (use -run_exp or -e<N> with the EXP_* environment variables); nothing to render offline.\n");
		destroy_application(&app);
		return 0;
	}
	int status = 0;
	for (;;) {
		int done = advance_experiments(&app);
		if (done < 0) { printf("Failed to apply changed settings. Shutting down.\n"); status = 1; break; }
		if (done > 0) { printf("All experiments finished. Shutting down.\n"); break; }
		if (app.experiment_list.state == experiment_state_new_experiment) continue;
		if (render_frame(&app)) { status = 1; break; }
	}
	destroy_application(&app);
	return status;
}
