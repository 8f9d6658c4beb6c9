/* light_camera_noise.c -- host arithmetic that feeds the per-frame constants and the light
 * buffer: polygonal lights (polygonal_light.c:27-118), camera matrices (camera.c:24-83),
 * 4x4 inverse and Wang hash (math_utilities.h:24-57), noise words (noise_table.c:24-28).
 * Everything is fp32 with the reference's operation order, so that the 256-byte constant
 * block and the light records come out bit-identical (checked against the reference's own
 * polygonal_light.c / camera.c / math_utilities.h compiled as-is, tests/test_host_layer.py). */
#include "risltc_host.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static char* duplicate_string(const char* s) {
	if (!s) return NULL;
	size_t n = strlen(s) + 1;
	char* r = (char*) malloc(n);
	memcpy(r, s, n);
	return r;
}

/* ---- polygonal lights ---- */

int set_polygonal_light_vertex_count(polygonal_light_t* light, uint32_t vertex_count) {
	if (vertex_count == light->vertex_count && light->vertices_plane_space && light->vertices_world_space)
		return 0;
	float* plane_space = (float*) calloc(vertex_count ? vertex_count : 1, sizeof(float) * 4);
	if (light->vertices_plane_space) {
		uint32_t keep = (vertex_count < light->vertex_count) ? vertex_count : light->vertex_count;
		memcpy(plane_space, light->vertices_plane_space, sizeof(float) * 4 * keep);
	}
	free(light->vertices_plane_space);
	free(light->vertices_world_space);
	light->vertices_plane_space = plane_space;
	light->vertices_world_space = (float*) calloc(vertex_count ? vertex_count : 1, sizeof(float) * 4);
	/* The reference compares after assigning, so it always reports "unchanged" (polygonal_light.c:39-40) */
	light->vertex_count = vertex_count;
	return 0;
}

void update_polygonal_light(polygonal_light_t* light) {
	light->inv_scaling_x = 1.0f / light->scaling_x;
	light->inv_scaling_y = 1.0f / light->scaling_y;
	/* Euler angles -> rotation: first about z, then y, then x */
	const float cx = cosf(light->rotation_angles[0]), sx = sinf(light->rotation_angles[0]);
	const float cy = cosf(light->rotation_angles[1]), sy = sinf(light->rotation_angles[1]);
	const float cz = cosf(light->rotation_angles[2]), sz = sinf(light->rotation_angles[2]);
	const float cxsy = cx * sy, sxsy = sx * sy;
	float r[3][4];
	r[0][0] = cy * cz;                r[0][1] = -cy * sz;               r[0][2] = -sy;      r[0][3] = 0.0f;
	r[1][0] = -sxsy * cz + cx * sz;   r[1][1] = sxsy * sz + cx * cz;    r[1][2] = -sx * cy; r[1][3] = 0.0f;
	r[2][0] = cxsy * cz + sx * sz;    r[2][1] = -cxsy * sz + sx * cz;   r[2][2] = cx * cy;  r[2][3] = 0.0f;
	memcpy(light->rotation, r, sizeof(r));
	/* plane space -> world space */
	const float scale[2] = { light->scaling_x, light->scaling_y };
	for (uint32_t v = 0; v != light->vertex_count; ++v) {
		const float* p = light->vertices_plane_space + 4 * v;
		float* w = light->vertices_world_space + 4 * v;
		for (uint32_t axis = 0; axis != 3; ++axis) {
			w[axis] = light->translation[axis];
			w[axis] += scale[0] * r[axis][0] * p[0];
			w[axis] += scale[1] * r[axis][1] * p[1];
		}
	}
	/* plane through the translation with the rotated z axis as normal */
	light->plane[0] = r[0][2];
	light->plane[1] = r[1][2];
	light->plane[2] = r[2][2];
	light->plane[3] = -(r[0][2] * light->translation[0] + r[1][2] * light->translation[1] + r[2][2] * light->translation[2]);
	/* signed area of the triangle fan around vertex 0 */
	float signed_area = 0.0f;
	const float* p0 = light->vertices_plane_space;
	for (uint32_t i = 0; i + 2 < light->vertex_count; ++i) {
		const float* pa = light->vertices_plane_space + 4 * (i + 2);
		const float* pb = light->vertices_plane_space + 4 * (i + 1);
		float ax = pa[0] - p0[0], bx = pb[0] - p0[0];
		float ay = pa[1] - p0[1], by = pb[1] - p0[1];
		signed_area += 0.5f * (ax * by - bx * ay);
	}
	signed_area *= scale[0] * scale[1];
	float abs_area = (signed_area < 0.0f) ? -signed_area : signed_area;
	light->area = abs_area;
	light->rcp_area = 1.0f / abs_area;
	/* this fork exports the flux verbatim as surface radiance (polygonal_light.c:92-94) */
	for (uint32_t i = 0; i != 3; ++i) light->surface_radiance[i] = light->radiant_flux[i];
	/* orient the plane so that the plane-space winding is positive */
	if (!(signed_area > 0.0f))
		for (uint32_t i = 0; i != 4; ++i) light->plane[i] = -light->plane[i];
}

polygonal_light_t duplicate_polygonal_light(const polygonal_light_t* light) {
	polygonal_light_t copy = *light;
	copy.texture_file_path = duplicate_string(light->texture_file_path);
	copy.vertex_count = 0;
	copy.vertices_plane_space = NULL;
	copy.vertices_world_space = NULL;
	set_polygonal_light_vertex_count(&copy, light->vertex_count);
	memcpy(copy.vertices_plane_space, light->vertices_plane_space, sizeof(float) * 4 * light->vertex_count);
	return copy;
}

void destroy_polygonal_light(polygonal_light_t* light) {
	free(light->vertices_plane_space);
	free(light->vertices_world_space);
	free(light->texture_file_path);
	memset(light, 0, sizeof(*light));
}

/* ---- camera ---- */

void get_world_to_view_space(float world_to_view_space[4][4], const first_person_camera_t* camera) {
	const float cos_x = cosf(camera->rotation_x), sin_x = sinf(camera->rotation_x);
	const float cos_z = cosf(camera->rotation_z), sin_z = sinf(camera->rotation_z);
	const float about_x[3][3] = { { 1.0f, 0.0f, 0.0f }, { 0.0f, cos_x, sin_x }, { 0.0f, -sin_x, cos_x } };
	const float about_z[3][3] = { { cos_z, sin_z, 0.0f }, { -sin_z, cos_z, 0.0f }, { 0.0f, 0.0f, 1.0f } };
	float view_to_world[3][3];
	memset(view_to_world, 0, sizeof(view_to_world));
	for (uint32_t i = 0; i != 3; ++i)
		for (uint32_t j = 0; j != 3; ++j)
			for (uint32_t k = 0; k != 3; ++k)
				view_to_world[i][j] += about_z[i][k] * about_x[k][j];
	float origin[3] = { 0.0f, 0.0f, 0.0f };
	for (uint32_t i = 0; i != 3; ++i)
		for (uint32_t j = 0; j != 3; ++j)
			origin[i] -= view_to_world[j][i] * camera->position_world_space[j];
	for (uint32_t row = 0; row != 3; ++row) {
		for (uint32_t col = 0; col != 3; ++col) world_to_view_space[row][col] = view_to_world[col][row];
		world_to_view_space[row][3] = origin[row];
	}
	world_to_view_space[3][0] = world_to_view_space[3][1] = world_to_view_space[3][2] = 0.0f;
	world_to_view_space[3][3] = 1.0f;
}

void get_view_to_projection_space(float view_to_projection_space[4][4], const first_person_camera_t* camera, float aspect_ratio) {
	const float near = camera->near, far = camera->far;
	const float top = tanf(0.5f * camera->vertical_fov);
	const float right = aspect_ratio * top;
	memset(view_to_projection_space, 0, sizeof(float) * 16);
	view_to_projection_space[0][0] = -1.0f / right;
	view_to_projection_space[1][1] = 1.0f / top;
	view_to_projection_space[2][2] = -(far + near) / (far - near);
	view_to_projection_space[2][3] = -2.0f * far * near / (far - near);
	view_to_projection_space[3][2] = -1.0f;
}

void get_world_to_projection_space(float world_to_projection_space[4][4], const first_person_camera_t* camera, float aspect_ratio) {
	float world_to_view[4][4], view_to_projection[4][4];
	get_world_to_view_space(world_to_view, camera);
	get_view_to_projection_space(view_to_projection, camera, aspect_ratio);
	memset(world_to_projection_space, 0, sizeof(float) * 16);
	for (uint32_t i = 0; i != 4; ++i)
		for (uint32_t j = 0; j != 4; ++j)
			for (uint32_t k = 0; k != 4; ++k)
				world_to_projection_space[i][j] += view_to_projection[i][k] * world_to_view[k][j];
}

/* ---- math utilities ---- */

/* Cofactor expansion with the reference's term order (math_utilities.h:24-47): every entry of the
 * adjugate is a sum of six signed triple products, listed here as {sign, a, b, c} over the flat
 * row-major index of the input; the order of the six terms fixes the fp32 rounding. */
void matrix_inverse(float inverse[4][4], const float matrix[4][4]) {
	static const signed char terms[16][6][4] = {
		{ { 1, 5, 10, 15 }, { -1, 5, 11, 14 }, { -1, 9, 6, 15 }, { 1, 9, 7, 14 }, { 1, 13, 6, 11 }, { -1, 13, 7, 10 } },
		{ { -1, 1, 10, 15 }, { 1, 1, 11, 14 }, { 1, 9, 2, 15 }, { -1, 9, 3, 14 }, { -1, 13, 2, 11 }, { 1, 13, 3, 10 } },
		{ { 1, 1, 6, 15 }, { -1, 1, 7, 14 }, { -1, 5, 2, 15 }, { 1, 5, 3, 14 }, { 1, 13, 2, 7 }, { -1, 13, 3, 6 } },
		{ { -1, 1, 6, 11 }, { 1, 1, 7, 10 }, { 1, 5, 2, 11 }, { -1, 5, 3, 10 }, { -1, 9, 2, 7 }, { 1, 9, 3, 6 } },
		{ { -1, 4, 10, 15 }, { 1, 4, 11, 14 }, { 1, 8, 6, 15 }, { -1, 8, 7, 14 }, { -1, 12, 6, 11 }, { 1, 12, 7, 10 } },
		{ { 1, 0, 10, 15 }, { -1, 0, 11, 14 }, { -1, 8, 2, 15 }, { 1, 8, 3, 14 }, { 1, 12, 2, 11 }, { -1, 12, 3, 10 } },
		{ { -1, 0, 6, 15 }, { 1, 0, 7, 14 }, { 1, 4, 2, 15 }, { -1, 4, 3, 14 }, { -1, 12, 2, 7 }, { 1, 12, 3, 6 } },
		{ { 1, 0, 6, 11 }, { -1, 0, 7, 10 }, { -1, 4, 2, 11 }, { 1, 4, 3, 10 }, { 1, 8, 2, 7 }, { -1, 8, 3, 6 } },
		{ { 1, 4, 9, 15 }, { -1, 4, 11, 13 }, { -1, 8, 5, 15 }, { 1, 8, 7, 13 }, { 1, 12, 5, 11 }, { -1, 12, 7, 9 } },
		{ { -1, 0, 9, 15 }, { 1, 0, 11, 13 }, { 1, 8, 1, 15 }, { -1, 8, 3, 13 }, { -1, 12, 1, 11 }, { 1, 12, 3, 9 } },
		{ { 1, 0, 5, 15 }, { -1, 0, 7, 13 }, { -1, 4, 1, 15 }, { 1, 4, 3, 13 }, { 1, 12, 1, 7 }, { -1, 12, 3, 5 } },
		{ { -1, 0, 5, 11 }, { 1, 0, 7, 9 }, { 1, 4, 1, 11 }, { -1, 4, 3, 9 }, { -1, 8, 1, 7 }, { 1, 8, 3, 5 } },
		{ { -1, 4, 9, 14 }, { 1, 4, 10, 13 }, { 1, 8, 5, 14 }, { -1, 8, 6, 13 }, { -1, 12, 5, 10 }, { 1, 12, 6, 9 } },
		{ { 1, 0, 9, 14 }, { -1, 0, 10, 13 }, { -1, 8, 1, 14 }, { 1, 8, 2, 13 }, { 1, 12, 1, 10 }, { -1, 12, 2, 9 } },
		{ { -1, 0, 5, 14 }, { 1, 0, 6, 13 }, { 1, 4, 1, 14 }, { -1, 4, 2, 13 }, { -1, 12, 1, 6 }, { 1, 12, 2, 5 } },
		{ { 1, 0, 5, 10 }, { -1, 0, 6, 9 }, { -1, 4, 1, 10 }, { 1, 4, 2, 9 }, { 1, 8, 1, 6 }, { -1, 8, 2, 5 } },
	};
	/* the table above is indexed by the flat OUTPUT index in the order 0,1,2,3 (first row), 4.. etc. */
	float* inv = &inverse[0][0];
	const float* m = &matrix[0][0];
	for (int e = 0; e != 16; ++e) {
		float acc = 0.0f;
		for (int t = 0; t != 6; ++t) {
			float a = m[terms[e][t][1]];
			if (terms[e][t][0] < 0) a = -a;
			float product = a * m[terms[e][t][2]] * m[terms[e][t][3]];
			acc = (t == 0) ? product : acc + product;
		}
		inv[e] = acc;
	}
	float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
	float rcp_det = 1.0f / det;
	for (int i = 0; i != 16; ++i) inv[i] = inv[i] * rcp_det;
}

uint32_t wang_random_number(uint32_t seed) {
	seed = (seed ^ 61u) ^ (seed >> 16);
	seed *= 9u;
	seed = seed ^ (seed >> 4);
	seed *= 0x27d4eb2du;
	seed = seed ^ (seed >> 15);
	return seed;
}

/* ---- noise ---- */

void set_noise_constants(uint32_t resolution_mask[2], uint32_t* texture_index_mask, uint32_t random_numbers[4], noise_table_t* noise, VkBool32 animate_noise) {
	(void) resolution_mask; (void) texture_index_mask;   /* this fork has no noise textures (noise_table.h:25-29) */
	for (uint32_t i = 0; i != 4; ++i)
		random_numbers[i] = animate_noise ? wang_random_number(noise->random_seed * 4 + i) : (i * 0x123456u);
	++noise->random_seed;
}
