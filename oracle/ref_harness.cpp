// ref_harness.cpp -- runs the reference's OWN shader source on the CPU. TEST INFRASTRUCTURE.
//
// oracle/build_ref.py turns /root/reference/src/shaders/shading_pass.frag.glsl (with its includes
// inlined) into one translation unit by mechanical token rewrites only (strip #version/#extension and
// [[unroll]] attributes, "inout T x" -> "T& x", append f to float literals, turn the uniform / buffer
// blocks into namespaces, thread_local for per-invocation in/out variables, main -> shader_main) and
// passes its path here as REF_SHADER_TU. Nothing of that text is stored in the repository; the only
// output is oracle/_ref/libref_shading_<variant>.so.
//
// The -D switches of main.c:962-991 are given on the compiler command line, so one library is one
// shader variant. Driver-side behaviour that is not in the tree (texel fetches, LTC sampler, ray
// queries, the visibility buffer) is provided by glsl_shim.hpp and forwarded to the caller.
#include "glsl_shim.hpp"
#include <cstdlib>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace glsl {
any_hit_fn g_any_hit = nullptr;
sample_texture_fn g_sample_texture = nullptr;
const void* g_any_hit_scene = nullptr;
thread_local unsigned long long g_ray_count = 0;
int g_rt_light_count = 1;

#include REF_SHADER_TU

}  // namespace glsl

using namespace glsl;

extern "C" {

// Scene binding: mesh texel buffers (bindings 1-3), flat material "textures" (binding 5), LTC arrays
// (binding 6), light SSBO (binding 8, records in the write_lights layout), any-hit callback (binding 9).
void ref_bind_scene(const uint32_t* positions, const uint16_t* normals_uv, const uint8_t* material_indices,
	const float* material_constants, uint32_t material_count,
	const float* light_records, uint32_t light_count, uint32_t record_floats,
	const uint16_t* ltc_rgba16, const uint16_t* ltc_rg16, uint32_t ltc_res, uint32_t ltc_layers,
	any_hit_fn any_hit, const void* any_hit_scene)
{
	g_quantized_vertex_positions.data = positions; g_quantized_vertex_positions.bytes_per_texel = 8;
	g_packed_normals_and_tex_coords.data = normals_uv;
	g_material_indices.data = material_indices; g_material_indices.bytes_per_texel = 1;
	for (uint32_t i = 0; i != material_count && i != MATERIAL_COUNT; ++i) {
		const float* m = material_constants + 8 * i;
		sampler2D base = { { m[0], m[1], m[2], 1.0f }, nullptr }, spec = { { m[3], m[4], m[5], 1.0f }, nullptr }, nrm = { { m[6], m[7], 1.0f, 1.0f }, nullptr };
		g_material_textures[3 * i + 0] = base; g_material_textures[3 * i + 1] = spec; g_material_textures[3 * i + 2] = nrm;
	}
	g_rt_light_count = (int) light_count;
	for (uint32_t i = 0; i <= light_count && i < POLYGONAL_LIGHT_ARRAY_SIZE; ++i) {
		// entry light_count repeats the last light: float(seed) * 2^-32 rounds to 1.0 for the top 128 seeds and
		// the shader then indexes one past the end (shading_pass.frag.glsl:712,730); defined as a clamp.
		const float* rec = light_records + (size_t) (i < light_count ? i : light_count - 1) * record_floats;
		polygonal_light_t& l = g_polygonal_lights[i];
		l.surface_radiance = vec3(rec[0], rec[1], rec[2]);
		l.plane = vec4(rec[4], rec[5], rec[6], rec[7]);
		std::memcpy(&l.vertex_count, rec + 8, 4);
		for (int v = 0; v != MAX_POLYGONAL_LIGHT_VERTEX_COUNT; ++v) l.vertices_world_space[v] = vec3(rec[12 + 4 * v], rec[13 + 4 * v], rec[14 + 4 * v]);
	}
	g_ltc_tables[0].data = ltc_rgba16; g_ltc_tables[0].channels = 4; g_ltc_tables[0].res = (int) ltc_res; g_ltc_tables[0].layers = (int) ltc_layers;
	g_ltc_tables[1].data = ltc_rg16; g_ltc_tables[1].channels = 2; g_ltc_tables[1].res = (int) ltc_res; g_ltc_tables[1].layers = (int) ltc_layers;
	g_any_hit = any_hit; g_any_hit_scene = any_hit_scene;
}

// Material textures (binding 5) as texture objects of the caller, 3 per material, sampled through the caller's textureGrad
// (null: back to the flat texels of ref_bind_scene)
void ref_bind_textures(const void* textures, uint32_t stride_bytes, uint32_t texture_count, sample_texture_fn sample) {
	g_sample_texture = sample;
	for (uint32_t i = 0; i != 3 * MATERIAL_COUNT; ++i)
		g_material_textures[i].texture = (textures && i < texture_count) ? (const char*) textures + (size_t) stride_bytes * i : nullptr;
}

// The 256-byte per_frame_constants_t block (main.h:537-553): std140 + row_major, so matrices arrive as rows.
void ref_set_constants(const void* block) {
	const float* f = (const float*) block;
	const uint32_t* u = (const uint32_t*) block;
	g_mesh_dequantization_factor = vec3(f[0], f[1], f[2]);
	g_mesh_dequantization_summand = vec3(f[4], f[5], f[6]);
	g_error_factor = f[7];
	for (int row = 0; row != 4; ++row) for (int col = 0; col != 4; ++col) g_world_to_projection_space[col][row] = f[8 + 4 * row + col];
	for (int row = 0; row != 3; ++row) for (int col = 0; col != 3; ++col) g_pixel_to_ray_direction_world_space[col][row] = f[24 + 4 * row + col];
	g_camera_position_world_space = vec3(f[36], f[37], f[38]);
	g_mis_visibility_estimate = f[39];
	g_viewport_size = uvec2(u[40], u[41]);
	g_cursor_position = ivec2((int) u[42], (int) u[43]);
	g_exposure_factor = f[44];
	g_roughness_factor = f[45];
	g_noise_random_numbers = uvec4(u[52], u[53], u[54], u[55]);
	g_ltc_constants.fresnel_index_factor = f[56]; g_ltc_constants.fresnel_index_summand = f[57];
	g_ltc_constants.roughness_factor = f[58]; g_ltc_constants.roughness_summand = f[59];
	g_ltc_constants.inclination_factor = f[60]; g_ltc_constants.inclination_summand = f[61];
}

// The shading subpass for rows [row_begin, row_end): one shader_main() per pixel. Emitter pixels are
// answered without running the shader, which would index the vertex buffers with a negative primitive
// index first (shading_pass.frag.glsl:692 before :696); their defined result is (1,1,1) * exposure.
unsigned long long ref_shade_rows(const uint32_t* visibility, float* out_rgba, uint32_t row_begin, uint32_t row_end) {
	const uint32_t W = g_viewport_size.x;
	unsigned long long rays = 0;
	#pragma omp parallel for schedule(dynamic, 4) reduction(+:rays)
	for (long long y = row_begin; y < (long long) row_end; ++y) {
		g_ray_count = 0;
		for (uint32_t x = 0; x != W; ++x) {
			uint32_t prim = visibility[(size_t) y * W + x];
			float* o = out_rgba + 4 * ((size_t) y * W + x);
			if (prim != 0xFFFFFFFFu && (prim >> 31)) { o[0] = o[1] = o[2] = 1.0f * g_exposure_factor; o[3] = 1.0f; continue; }
			gl_FragCoord = vec4((float) x + 0.5f, (float) y + 0.5f, 0.0f, 1.0f);
			g_visibility_buffer.value = prim;
			g_out_color = vec4(0.0f, 0.0f, 0.0f, 0.0f);
			shader_main();
			o[0] = g_out_color.x; o[1] = g_out_color.y; o[2] = g_out_color.z; o[3] = g_out_color.w;
		}
		rays += g_ray_count;
	}
	return rays;
}

// ---- function-level entry points (same signatures as the orc_* ones, so tests can diff them)
uint32_t ref_clip_polygon(uint32_t vertex_count, float v[8][3]) {
	vec3 p[MAX_POLYGON_VERTEX_COUNT];
	for (int i = 0; i != MAX_POLYGON_VERTEX_COUNT; ++i) p[i] = vec3(v[i][0], v[i][1], v[i][2]);
	uint32_t vc = clip_polygon(vertex_count, p);
	for (int i = 0; i != MAX_POLYGON_VERTEX_COUNT; ++i) { v[i][0] = p[i].x; v[i][1] = p[i].y; v[i][2] = p[i].z; }
	return vc;
}

float ref_calculate_ltc(uint32_t vertex_count, const float v[8][3]) {
	vec3 p[MAX_POLYGON_VERTEX_COUNT];
	for (int i = 0; i != MAX_POLYGON_VERTEX_COUNT; ++i) p[i] = vec3(v[i][0], v[i][1], v[i][2]);
	return calculate_ltc(vertex_count, p);
}

// out: {vertex_count, vertices[8][2], ellipses[8][2], inner_ellipse_0[2], sectors[8], total} = 44 floats
void ref_psa(uint32_t vertex_count, const float v[8][3], float u0, float u1, float* out_polygon, float* out_dir) {
	vec3 p[MAX_POLYGON_VERTEX_COUNT];
	for (int i = 0; i != MAX_POLYGON_VERTEX_COUNT; ++i) p[i] = vec3(v[i][0], v[i][1], v[i][2]);
	projected_solid_angle_polygon_t poly = prepare_projected_solid_angle_polygon_sampling(vertex_count, p);
	for (int k = 0; k != 44; ++k) out_polygon[k] = 0.0f;
	out_polygon[0] = (float) poly.vertex_count;
	for (uint32_t k = 0; k != vertex_count && k != MAX_POLYGON_VERTEX_COUNT; ++k) {
		out_polygon[1 + 2 * k] = poly.vertices[k].x; out_polygon[2 + 2 * k] = poly.vertices[k].y;
		out_polygon[17 + 2 * k] = poly.ellipses[k].x; out_polygon[18 + 2 * k] = poly.ellipses[k].y;
		out_polygon[35 + k] = poly.sector_projected_solid_angles[k];
	}
	out_polygon[33] = poly.inner_ellipse_0.x; out_polygon[34] = poly.inner_ellipse_0.y;
	out_polygon[43] = poly.projected_solid_angle;
	vec3 d = sample_projected_solid_angle_polygon(poly, vec2(u0, u1));
	out_dir[0] = d.x; out_dir[1] = d.y; out_dir[2] = d.z;
}

// out: world_to_shading[12] (column-major), shading_to_cosine[9], cosine_to_shading[9], albedo, determinant
void ref_ltc_coefficients(float fresnel_0, float roughness, const float* pos, const float* normal, const float* outgoing, const float* c, float* out) {
	ltc_constants_t k;
	k.fresnel_index_factor = c[0]; k.fresnel_index_summand = c[1]; k.roughness_factor = c[2]; k.roughness_summand = c[3];
	k.inclination_factor = c[4]; k.inclination_summand = c[5];
	ltc_coefficients_t l = get_ltc_coefficients(fresnel_0, roughness, vec3(pos[0], pos[1], pos[2]), vec3(normal[0], normal[1], normal[2]), vec3(outgoing[0], outgoing[1], outgoing[2]), k);
	for (int j = 0; j != 4; ++j) for (int i = 0; i != 3; ++i) out[3 * j + i] = l.world_to_shading_space[j][i];
	for (int j = 0; j != 3; ++j) for (int i = 0; i != 3; ++i) { out[12 + 3 * j + i] = l.shading_to_cosine_space[j][i]; out[21 + 3 * j + i] = l.cosine_to_shading_space[j][i]; }
	out[30] = l.albedo; out[31] = l.shading_to_cosine_space_determinant;
}

void ref_noise(uint32_t px, uint32_t py, uint32_t width, uint32_t frame_word, uint32_t draws, float* out) {
	noise_accessor_t a = get_noise_accessor(uvec2(px, py), uvec2(width, 0), uvec4(frame_word, 0, 0, 0));
	for (uint32_t i = 0; i != draws; ++i) out[i] = get_noise_1(a);
}

int ref_thread_count(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

uint32_t ref_max_polygon_vertex_count(void) { return MAX_POLYGON_VERTEX_COUNT; }

}  // extern "C"

// get_shading_data for one pixel; same output layout as orc_shading_data. Call ref_set_constants first.
extern "C" void ref_shading_data(uint32_t x, uint32_t y, uint32_t primitive_index, float* out) {
	ivec2 pixel((int) x, (int) y);
	vec3 dir = g_pixel_to_ray_direction_world_space * vec3(pixel, 1.0f);
	shading_data_t s = get_shading_data(pixel, (int) primitive_index, dir);
	const float v[17] = { s.position.x, s.position.y, s.position.z, s.normal.x, s.normal.y, s.normal.z, s.outgoing.x, s.outgoing.y, s.outgoing.z,
		s.lambert_outgoing, s.diffuse_albedo.x, s.diffuse_albedo.y, s.diffuse_albedo.z, s.fresnel_0.x, s.fresnel_0.y, s.fresnel_0.z, s.roughness };
	std::memcpy(out, v, sizeof(v));
}
