/* experiment_list.c -- builds the table of offline experiments from environment variables,
 * with the reference's names, settings, paths and ordering (experiment_list.c:34-470), written as
 * a table of (name, samples, light sampling, polygon technique) rows per experiment family.
 *
 * Environment: EXP_TEASER, EXP_FIG1, EXP_LO_ROUGH, EXP_MED_ROUGH, EXP_HI_ROUGH, EXP_DIFFUSE select
 * families; EXP_COMPARE, EXP_TIMINGS, COMPUTE_GT select rows inside the roughness families;
 * EXP_ENSURE_CORRECT adds the cross-estimator check; NUM_SAMPLES (default 10000) and
 * SCENE ("bistro_inside", anything else = bistro_exterior) parameterise them. */
#include "risltc_host.h"
#include <stdlib.h>
#include <string.h>

static char* dup_string(const char* s) {
	if (!s) return NULL;
	size_t n = strlen(s) + 1;
	char* r = (char*) malloc(n);
	memcpy(r, s, n);
	return r;
}

static char* join3(const char* a, const char* b, const char* c) {
	size_t n = strlen(a) + strlen(b) + strlen(c) + 1;
	char* r = (char*) malloc(n);
	snprintf(r, n, "%s%s%s", a, b, c);
	return r;
}

typedef struct row_s {
	const char* name;
	uint32_t samples;          /* 0 = NUM_SAMPLES */
	int light_sampling;        /* -1 = keep the family's setting */
	int technique;             /* -1 = keep the family's setting */
	int ss_per_frame;          /* -1 = keep */
} row_t;

typedef struct family_s {
	scene_index_t scene;
	uint32_t width, height;
	float exposure, roughness;
	const char* quick_save_path;
	const char* base_dir;
	VkBool32 ss_per_frame;
} family_t;

static void append(experiment_list_t* list, uint32_t capacity, const family_t* f, const row_t* row, uint32_t default_samples) {
	if (list->count >= capacity) { ++list->count; return; }
	experiment_t* e = &list->experiments[list->count++];
	memset(e, 0, sizeof(*e));
	e->width = f->width; e->height = f->height;
	e->scene_index = f->scene;
	e->quick_save_path = dup_string(f->quick_save_path);
	e->use_hdr = VK_TRUE;
	e->base_dir = dup_string(f->base_dir);
	e->ext = dup_string("hdr");
	e->ss_per_frame = (row->ss_per_frame >= 0) ? (VkBool32) row->ss_per_frame : f->ss_per_frame;
	e->num_samples = row->samples ? row->samples : default_samples;
	render_settings_t* s = &e->render_settings;
	s->exposure_factor = f->exposure; s->roughness_factor = f->roughness;
	s->sample_count = 1; s->sample_count_light = 1;
	s->mis_heuristic = mis_heuristic_optimal_clamped; s->mis_visibility_estimate = 0.5f;
	s->animate_noise = VK_TRUE; s->show_polygonal_lights = VK_FALSE; s->accum = VK_TRUE;
	s->light_sampling = light_reservoir; s->fast_atan = VK_FALSE;
	/* the families leave polygon_sampling_technique zero-initialised, i.e. sample_polygon_baseline */
	s->polygon_sampling_technique = sample_polygon_baseline;
	if (row->light_sampling >= 0) s->light_sampling = (light_sampling_strategies_t) row->light_sampling;
	if (row->technique >= 0) s->polygon_sampling_technique = (sample_polygon_technique_t) row->technique;
	e->exp_name = dup_string(row->name);
	/* fill_path_info, experiment_list.c:22-32 */
	char* dir = join3(e->base_dir, e->exp_name, "");
	e->screenshots_dir = dir;
	char* stem = join3(e->base_dir, e->exp_name, "/00000");
	e->screenshot_path = join3(stem, ".", e->ext);
	free(stem);
	e->timings_path = join3(e->base_dir, e->exp_name, "/timings.txt");
}

void create_experiment_list(experiment_list_t* list) {
	memset(list, 0, sizeof(*list));
	const uint32_t capacity = 1000;
	list->experiments = (experiment_t*) calloc(capacity, sizeof(experiment_t));
	uint32_t sample_count = 0;
	const char* sample_str = getenv("NUM_SAMPLES");
	if (sample_str) sample_count = (uint32_t) atoi(sample_str);
	if (sample_count == 0) sample_count = 10000;
	const char* scene_name = getenv("SCENE");
	int inside = scene_name && strcmp(scene_name, "bistro_inside") == 0;
	scene_index_t scene = inside ? scene_bistro_inside : scene_bistro_outside;
	const char* quick_save_path = inside ? "data/quicksaves/Bistro_interior.save" : "data/quicksaves/Bistro_exterior.save";
	const char* scene_dir = inside ? "bistro_inside/" : "bistro_exterior/";
	const VkBool32 compute_gt = getenv("COMPUTE_GT") != NULL;
	printf("Requested %d samples per experiment\n", sample_count);

	const int U = light_uniform, TURK = sample_polygon_area_turk, PSA = sample_polygon_projected_solid_angle, LTC = sample_polygon_ltc_cp;

	if (getenv("EXP_TEASER")) {
		const family_t f = { scene_bistro_outside, 1920, 1080, 1.5f, 0.1f, "data/quicksaves/teaser.save", "data/experiments/teaser/", VK_TRUE };
		const row_t rows[] = { { "uniform", 10000, U, TURK, -1 }, { "ris", 10000, -1, TURK, -1 }, { "ours", 10000, -1, LTC, -1 }, { "ris_projltc", 10000, -1, PSA, -1 } };
		for (uint32_t i = 0; i != sizeof(rows) / sizeof(rows[0]); ++i) append(list, capacity, &f, &rows[i], sample_count);
		const row_t gt = { "gt", 1000000, U, PSA, 0 };
		if (compute_gt) append(list, capacity, &f, &gt, sample_count);
	}
	if (getenv("EXP_FIG1")) {
		const family_t f = { scene_bistro_inside, 1920, 1080, 1.5f, 0.1f, "data/quicksaves/fig1.save", "data/experiments/fig1/", VK_TRUE };
		const row_t rows[] = { { "ours", 100, -1, LTC, -1 }, { "ris_projltc", 100, -1, PSA, -1 } };
		for (uint32_t i = 0; i != sizeof(rows) / sizeof(rows[0]); ++i) append(list, capacity, &f, &rows[i], sample_count);
		const row_t gt = { "gt", 1000000, U, PSA, 0 };
		if (compute_gt) append(list, capacity, &f, &gt, sample_count);
	}
	/* roughness families: same scene, quicksave and rows, different roughness factor and output root */
	const struct { const char* env; const char* root; float roughness; } rough[4] = {
		{ "EXP_LO_ROUGH", "data/experiments/lo_rough/", 0.05f },
		{ "EXP_MED_ROUGH", "E:/renders/med_rough/", 0.1f },
		{ "EXP_HI_ROUGH", "data/experiments/hi_rough/", 0.3f },
		{ "EXP_DIFFUSE", "E:/renders/diffuse/", 1.0f },
	};
	for (uint32_t k = 0; k != 4; ++k) {
		if (!getenv(rough[k].env)) continue;
		char* base_dir = join3(rough[k].root, scene_dir, "");
		const family_t f = { scene, 1920, 1080, 1.5f, rough[k].roughness, quick_save_path, base_dir, VK_TRUE };
		if (getenv("EXP_COMPARE")) {
			const row_t rows[] = { { "uniform_uniform", 0, U, TURK, -1 }, { "uniform_cp", 0, U, PSA, -1 }, { "uniform_area", 0, -1, TURK, -1 },
			                       { "ltc_cp", 0, -1, -1, -1 }, { "cp_cp", 0, -1, PSA, -1 } };
			for (uint32_t i = 0; i != sizeof(rows) / sizeof(rows[0]); ++i) append(list, capacity, &f, &rows[i], sample_count);
		}
		if (getenv("EXP_TIMINGS")) {
			const row_t rows[] = { { "uniform_uniform_time", 1000, U, TURK, 0 }, { "uniform_cp_time", 1000, U, PSA, 0 }, { "uniform_area_time", 1000, -1, TURK, 0 },
			                       { "cp_cp_time", 1000, -1, PSA, 0 }, { "ltc_cp_time", 1000, -1, LTC, 0 } };
			for (uint32_t i = 0; i != sizeof(rows) / sizeof(rows[0]); ++i) append(list, capacity, &f, &rows[i], sample_count);
		}
		const row_t gt = { "gt", 100000, U, PSA, 0 };
		if (compute_gt) append(list, capacity, &f, &gt, sample_count);
		free(base_dir);
	}
	if (getenv("EXP_ENSURE_CORRECT")) {
		const family_t f = { scene_bistro_outside, 1280, 720, 2.0f, 0.1f, "data/quicksaves/Bistro_exterior.save", "data/experiments/ensure_correct/", VK_FALSE };
		const row_t rows[] = { { "uniform_area", 100000, -1, TURK, -1 }, { "cp_cp", 30000, -1, PSA, -1 }, { "ltc_cp", 30000, -1, LTC, -1 } };
		for (uint32_t i = 0; i != sizeof(rows) / sizeof(rows[0]); ++i) append(list, capacity, &f, &rows[i], sample_count);
	}
	if (list->count > capacity) {
		printf("WARNING: Insufficient space allocated for %d experiments.\n", list->count);
		list->count = capacity;
	}
	else
		printf("Defined %d experiments to reproduce.\n", list->count);
	list->next = list->count + 1;
}

void destroy_experiment_list(experiment_list_t* list) {
	for (uint32_t i = 0; i != list->count; ++i) {
		experiment_t* e = &list->experiments[i];
		free(e->quick_save_path); free(e->screenshot_path); free(e->base_dir); free(e->timings_path);
		free(e->screenshots_dir); free(e->ext); free(e->exp_name);
	}
	free(list->experiments);
	memset(list, 0, sizeof(*list));
}
