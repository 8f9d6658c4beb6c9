/* embedding.c -- flat entry points for embedding the host layer in another process (the Python
 * test / bench drivers, the multi-GPU launcher). They only compose the reference-shaped functions
 * of risltc_host.h; nothing here computes. */
#include "risltc_host.h"
#include "risltc_cuda.h"
#include <stdlib.h>
#include <string.h>

static char* dup_string(const char* s) {
	size_t n = strlen(s) + 1;
	char* r = (char*) malloc(n);
	memcpy(r, s, n);
	return r;
}

/* startup_application without an experiment; LTC tables come from <RISLTC_DATA_DIR>/ggx_ltc_fit. */
application_t* risltc_app_create(int cuda_ordinal, uint32_t stripe_height, uint32_t stripe_index, uint32_t stripe_count) {
	application_t* app = (application_t*) calloc(1, sizeof(application_t));
	app->stripe_height = stripe_height; app->stripe_index = stripe_index; app->stripe_count = stripe_count;
	if (startup_application(app, -1, bool_override_false, cuda_ordinal)) { free(app); return NULL; }
	return app;
}

void risltc_app_destroy(application_t* app) {
	if (!app) return;
	destroy_application(app);
	free(app);
}

/* Point the scene specification at a scene / texture directory / quicksave and (re)load everything. */
int risltc_app_load(application_t* app, const char* vks_path, const char* texture_dir, const char* quick_save_path, uint32_t width, uint32_t height) {
	scene_specification_t* s = &app->scene_specification;
	free(s->file_path); free(s->texture_path); free(s->quick_save_path);
	s->file_path = dup_string(vks_path); s->texture_path = dup_string(texture_dir); s->quick_save_path = dup_string(quick_save_path);
	app->swapchain.extent.width = width; app->swapchain.extent.height = height;
	application_updates_t updates;
	memset(&updates, 0, sizeof(updates));
	updates.reload_scene = updates.quick_load = updates.recreate_swapchain = updates.change_shading = VK_TRUE;
	return update_application(app, &updates);
}

int risltc_app_set_render_settings(application_t* app, const render_settings_t* settings) {
	app->render_settings = *settings;
	application_updates_t updates;
	memset(&updates, 0, sizeof(updates));
	updates.change_shading = VK_TRUE;
	return update_application(app, &updates);
}

void risltc_app_get_render_settings(const application_t* app, render_settings_t* settings) { *settings = app->render_settings; }
void risltc_app_reset(application_t* app, uint32_t random_seed) { app->accum_num = 0; app->noise_table.random_seed = random_seed; }
uint32_t risltc_app_accum_num(const application_t* app) { return app->accum_num; }
struct risltc_device_s* risltc_app_device(application_t* app) { return app->device.cuda; }
uint32_t risltc_app_light_count(const application_t* app) { return app->scene_specification.polygonal_light_count; }
size_t risltc_app_light_buffer_size(const application_t* app) { return get_light_buffer_size(&app->scene_specification); }

/* `frame_count` frames with accumulation, submitted as one batch: write_constants per frame on the
 * host, then a single asynchronous launch sequence (the reference submits and waits per frame because
 * it reads its timestamps blocking, main.c:3008-3012; batching changes no result). Lights are
 * re-uploaded first, like write_lights on a light-set change (main.c:456-490). */
int risltc_app_render_frames(application_t* app, uint32_t frame_count, int upload_lights) {
	if (upload_lights) {
		size_t size = get_light_buffer_size(&app->scene_specification);
		void* records = malloc(size ? size : 1);
		write_lights(records, app);
		int result = risltc_cuda_upload_lights(app->device.cuda, records, app->scene_specification.polygonal_light_count,
			get_max_polygonal_light_vertex_count(&app->scene_specification));
		free(records);
		if (result) return 1;
	}
	per_frame_constants_t* blocks = (per_frame_constants_t*) malloc(sizeof(per_frame_constants_t) * (frame_count ? frame_count : 1));
	for (uint32_t i = 0; i != frame_count; ++i) write_constants(&blocks[i], app);
	int result = risltc_cuda_render_frames(app->device.cuda, blocks, frame_count, app->accum_num);
	free(blocks);
	if (result) return 1;
	app->accum_num += frame_count;
	return 0;
}

float risltc_app_wait(application_t* app) {
	app->last_frame_ms = risltc_cuda_last_frame_ms(app->device.cuda);
	return app->last_frame_ms;
}

/* implement_screenshot of the frame in the accumulation buffer (screenshot.c); NULL skips a format */
int risltc_app_screenshot(application_t* app, const char* path_png, const char* path_hdr) {
	return take_screenshot(app, path_png, path_hdr);
}
