/* risltc_oracle.c -- see risltc_oracle.h. TEST INFRASTRUCTURE, not product code.
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off matters: a*b+c is two roundings here; only the places where
 * the GLSL says fma() use fmaf(). GLSL built-ins are given these definitions
 * (shared with oracle/glsl_shim.hpp so that the compiled-GLSL reference build
 * and this file can be compared bit for bit):
 *   dot(a,b)      = ((a.x*b.x + a.y*b.y) + a.z*b.z) [+ a.w*b.w]
 *   cross(a,b)    = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
 *   length(v)     = sqrtf(dot(v,v));  inversesqrt(x) = 1.0f / sqrtf(x)
 *   normalize(v)  = v * inversesqrt(dot(v,v))
 *   M*v           = sum over columns j (in order) of column_j * v_j
 *   mix(a,b,t)    = a*(1-t) + b*t;  clamp(x,lo,hi) = min(max(x,lo),hi)
 *   atan/acos/sin/cos = libm float versions; x/y = IEEE division
 *   float(uint)   = round to nearest even
 */
#include "risltc_oracle.h"
#include "clip_rotation_table.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846f
#define ORC_HALF_PI 1.57079632679489661923f
#define ORC_INV_PI 0.318309886183790671538f

typedef struct { float x, y; } v2;
typedef struct { float x, y, z; } v3;

static inline v2 mk2(float x, float y) { v2 r = { x, y }; return r; }
static inline v3 mk3(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
static inline float dot2(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 add3(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline v3 mul3v(v3 a, v3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v2 add2(v2 a, v2 b) { return mk2(a.x + b.x, a.y + b.y); }
static inline v2 sub2(v2 a, v2 b) { return mk2(a.x - b.x, a.y - b.y); }
static inline v2 mul2(v2 a, float s) { return mk2(a.x * s, a.y * s); }
static inline v3 cross3(v3 a, v3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float inversesqrt(float x) { return 1.0f / sqrtf(x); }
static inline float length3(v3 a) { return sqrtf(dot3(a, a)); }
static inline v3 normalize3(v3 a) { return mul3(a, inversesqrt(dot3(a, a))); }
static inline v2 normalize2(v2 a) { return mul2(a, inversesqrt(dot2(a, a))); }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
/* mat3 [col][row] times vec3 */
static inline v3 m3_mul(const float m[3][3], v3 v) {
	return mk3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
	           m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
	           m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}
/* mat4x3 [col][row] times vec4(v, 1) */
static inline v3 m43_mul_point(const float m[4][3], v3 v) {
	return mk3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z + m[3][0] * 1.0f,
	           m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z + m[3][1] * 1.0f,
	           m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z + m[3][2] * 1.0f);
}

/* ------------------------------------------------------------------ noise */

/* math_utilities.h:50-57 */
uint32_t orc_wang_random_number(uint32_t seed) {
	seed = (seed ^ 61u) ^ (seed >> 16);
	seed *= 9u;
	seed ^= seed >> 4;
	seed *= 0x27d4eb2du;
	seed ^= seed >> 15;
	return seed;
}

/* noise_utility.glsl:26-42 */
static uint32_t murmur3_mix(uint32_t hash, uint32_t k) {
	k *= 0xcc9e2d51u;
	k = (k << 15) | (k >> 17);
	k *= 0x1b873593u;
	hash ^= k;
	hash = ((hash << 13) | (hash >> 19)) * 5u + 0xe6546b64u;
	return hash;
}

/* noise_utility.glsl:44-52 */
static uint32_t murmur3_finalize(uint32_t hash) {
	hash ^= hash >> 16;
	hash *= 0x85ebca6bu;
	hash ^= hash >> 13;
	hash *= 0xc2b2ae35u;
	hash ^= hash >> 16;
	return hash;
}

/* get_noise_accessor, noise_utility.glsl:77-84 (only random_numbers.x is used) */
uint32_t orc_noise_seed(uint32_t px, uint32_t py, uint32_t width, uint32_t frame_word) {
	uint32_t index = murmur3_mix(0u, px + py * width);
	return murmur3_finalize(murmur3_mix(index, frame_word));
}

/* rand_lcg + get_noise_gen, noise_utility.glsl:63-72 */
float orc_noise_next(uint32_t* seed) {
	*seed = 1664525u * (*seed) + 1013904223u;
	return (float) (*seed) * (1.0f / 4294967296.0f);
}

/* --------------------------------------------------------- polygon helpers */

/* GLSL prescribes no rounding for atan / acos / sin / cos (every driver has its own), so the oracle has to DEFINE them:
 * the correctly rounded fp32 value, obtained as the double-precision libm function rounded once (the double result is
 * within an ulp of the true value; the chance that rounding it to fp32 differs from rounding the true value is ~2^-29 per
 * call). The definition is independent of the libm version and can be met bit for bit by the CUDA kernels' exact mode
 * (csrc/common.cuh, RL_CR_LIBM). oracle/glsl_shim.hpp gives the compiled reference GLSL the same definition. The host
 * arithmetic below (light rotation, camera matrices) keeps cosf / sinf / tanf: it is C in the reference too. */
static inline float glsl_atan(float x) { return (float) atan((double) x); }
static inline float glsl_acos(float x) { return (float) acos((double) x); }
static inline float glsl_sin(float x) { return (float) sin((double) x); }
static inline float glsl_cos(float x) { return (float) cos((double) x); }

/* polygon_sampling.glsl:84-98 */
static float fast_positive_atan(float y) {
	float rx, ry, rz;
	rx = (fabsf(y) > 1.0f) ? (1.0f / fabsf(y)) : fabsf(y);
	ry = rx * rx;
	rz = fmaf(ry, 0.02083509974181652f, -0.08513300120830536f);
	rz = fmaf(ry, rz, 0.18014100193977356f);
	rz = fmaf(ry, rz, -0.3302994966506958f);
	ry = fmaf(ry, rz, 0.9998660087585449f);
	rz = fmaf(-2.0f * ry, rx, ORC_HALF_PI);
	rz = (fabsf(y) > 1.0f) ? rz : 0.0f;
	rx = fmaf(rx, ry, rz);
	return (y < 0.0f) ? (ORC_PI - rx) : rx;
}

/* polygon_sampling.glsl:105-112 */
static float positive_atan(float tangent, int fast_atan) {
	if (fast_atan) return fast_positive_atan(tangent);
	float offset = (tangent < 0.0f) ? ORC_PI : 0.0f;
	return glsl_atan(tangent) + offset;
}

/* polygon_sampling.glsl:184-186 */
static inline float mix_fma(float x, float y, float a) { return fmaf(a, y, fmaf(-a, x, x)); }

/* polygon_sampling.glsl:262-270 */
static inline float kahan(float a, float b, float c, float d) {
	float cd = c * d;
	float error = fmaf(c, d, -cd);
	float result = fmaf(a, b, -cd);
	return result - error;
}

/* polygon_sampling.glsl:275-281 */
static inline v3 cross_stable(v3 l, v3 r) {
	return mk3(kahan(l.y, r.z, l.z, r.y), kahan(l.z, r.x, l.x, r.z), kahan(l.x, r.y, l.y, r.x));
}

static inline v2 rotate_90(v2 v) { return mk2(-v.y, v.x); }

/* polygon_sampling.glsl:293-300: the sign BIT, so -0 counts as inner */
static inline int is_inner_ellipse(v2 e) { return (f2u(e.x) & 0x80000000u) != 0; }

/* polygon_sampling.glsl:320-329 */
static v2 ellipse_from_edge(v3 vertex_0, v3 vertex_1) {
	v3 normal = cross_stable(vertex_0, vertex_1);
	float scaling = 1.0f / normal.z;
	scaling = is_inner_ellipse(mk2(normal.x, normal.y)) ? -scaling : scaling;
	v2 ellipse = mk2(normal.x * scaling, normal.y * scaling);
	ellipse.x = (normal.z != 0.0f) ? ellipse.x : INFINITY;
	return ellipse;
}

/* polygon_sampling.glsl:335-337 */
static inline v2 ellipse_transform(v2 e, v2 p) {
	float d = dot2(e, p);
	return mk2(fmaf(d, e.x, p.x), fmaf(d, e.y, p.y));
}
/* polygon_sampling.glsl:343-351 */
static inline float get_ellipse_det(v2 e) { return fmaf(e.x, e.x, fmaf(e.y, e.y, 1.0f)); }
static inline float get_ellipse_rsqrt_det(v2 e) { return inversesqrt(get_ellipse_det(e)); }
/* polygon_sampling.glsl:354-358 */
static inline float get_ellipse_direction_factor_rsq(v2 e, v2 dir) {
	float ed = dot2(e, dir);
	float dd = dot2(dir, dir);
	return fmaf(ed, ed, dd);
}
/* polygon_sampling.glsl:367-369 */
static inline float get_ellipse_direction_factor(v2 e, v2 dir) { return inversesqrt(get_ellipse_direction_factor_rsq(e, dir)); }
/* polygon_sampling.glsl:373-376 */
static inline float get_ellipse_normalized_direction_factor(v2 e, v2 ndir) {
	float ed = dot2(e, ndir);
	return inversesqrt(fmaf(ed, ed, 1.0f));
}

/* polygon_sampling.glsl:381-386 */
static float area_between_ellipses_from_tangents(float inner_rsqrt_det, float inner_tangent, float outer_rsqrt_det, float outer_tangent, int fast_atan) {
	float inner_area = inner_rsqrt_det * positive_atan(inner_tangent, fast_atan);
	float result = fmaf(outer_rsqrt_det, positive_atan(outer_tangent, fast_atan), -inner_area);
	return (result > 0.0f) ? (0.5f * result) : 0.0f;
}

/* polygon_sampling.glsl:394-401 */
static float area_between_ellipses_in_sector(v2 inner, float inner_rsqrt_det, v2 outer, float outer_rsqrt_det, v2 dir_0, v2 dir_1, int fast_atan) {
	float det_dirs = fmaxf(+0.0f, dot2(dir_1, rotate_90(dir_0)));
	float inner_dot = inner_rsqrt_det * dot2(dir_0, ellipse_transform(inner, dir_1));
	float outer_dot = outer_rsqrt_det * dot2(dir_0, ellipse_transform(outer, dir_1));
	return area_between_ellipses_from_tangents(inner_rsqrt_det, det_dirs / inner_dot, outer_rsqrt_det, det_dirs / outer_dot, fast_atan);
}

/* polygon_sampling.glsl:409-416 */
static float ellipse_area_in_sector(v2 ellipse, v2 dir_0, v2 dir_1, int fast_atan) {
	float rsqrt_det = get_ellipse_rsqrt_det(ellipse);
	float det_dirs = fmaxf(+0.0f, dot2(dir_1, rotate_90(dir_0)));
	float ellipse_dot = rsqrt_det * dot2(dir_0, ellipse_transform(ellipse, dir_1));
	float area = 0.5f * rsqrt_det * positive_atan(det_dirs / ellipse_dot, fast_atan);
	return (rsqrt_det > 0.0f) ? area : 0.0f;
}

/* polygon_sampling.glsl:425-439 */
static void compare_and_swap(orc_psa_polygon_t* p, uint32_t lhs, uint32_t rhs) {
	v2 l = mk2(p->vertices[lhs][0], p->vertices[lhs][1]);
	v2 r = mk2(p->vertices[rhs][0], p->vertices[rhs][1]);
	float normal_z = kahan(l.x, -r.y, l.y, -r.x);
	int swap = (normal_z == 0.0f) ? (isinf(p->ellipses[rhs][0]) != 0) : (normal_z > 0.0f);
	if (swap) {
		for (int k = 0; k != 2; ++k) {
			float t = p->vertices[lhs][k]; p->vertices[lhs][k] = p->vertices[rhs][k]; p->vertices[rhs][k] = t;
			t = p->ellipses[lhs][k]; p->ellipses[lhs][k] = p->ellipses[rhs][k]; p->ellipses[rhs][k] = t;
		}
	}
}

/* polygon_sampling.glsl:444-506: the comparator lists per vertex count */
static void sort_convex_polygon_vertices(orc_psa_polygon_t* p) {
	static const unsigned char net5[][2] = { {2,4},{1,3},{1,2},{0,3},{3,4} };
	static const unsigned char net6[][2] = { {3,5},{2,4},{1,5},{0,4},{4,5},{1,3} };
	static const unsigned char net7[][2] = { {2,5},{1,6},{5,6},{3,4},{0,4},{4,6},{1,3},{3,5},{4,5} };
	static const unsigned char net8[][2] = { {2,6},{3,7},{1,5},{0,4},{4,6},{5,7},{6,7},{4,5},{1,3} };
	uint32_t n = p->vertex_count;
	if (n == 3) compare_and_swap(p, 1, 2);
	else if (n == 4) compare_and_swap(p, 1, 3);
	else if (n == 5) for (int i = 0; i != 5; ++i) compare_and_swap(p, net5[i][0], net5[i][1]);
	else if (n == 6) for (int i = 0; i != 6; ++i) compare_and_swap(p, net6[i][0], net6[i][1]);
	else if (n == 7) for (int i = 0; i != 9; ++i) compare_and_swap(p, net7[i][0], net7[i][1]);
	else if (n == 8) for (int i = 0; i != 9; ++i) compare_and_swap(p, net8[i][0], net8[i][1]);
	compare_and_swap(p, 0, 2);
	if (n >= 4) compare_and_swap(p, 2, 3);
	compare_and_swap(p, 0, 1);
}

/* polygon_sampling.glsl:508-521 */
static float integrate_edge_vec(v3 v1, v3 v2_) {
	v1 = normalize3(v1);
	v2_ = normalize3(v2_);
	float x = dot3(v1, v2_);
	float y = fabsf(x);
	float a = 0.8543985f + (0.4965155f + 0.0145206f * y) * y;
	float b = 3.4175940f + (4.1616724f + y) * y;
	float v = a / b;
	float theta_sintheta = (x > 0.0f) ? v : 0.5f * (1.0f / sqrtf(fmaxf(1.0f - x * x, 1e-7f))) - v;
	return cross3(v1, v2_).z * theta_sintheta;
}

/* polygon_sampling.glsl:523-530. Hazard: the reference loops over
 * MAX_POLYGON_VERTEX_COUNT edges and reads slots beyond vertex_count that are
 * only defined when vc >= MAX-1; the defined behaviour, used here, is the closed
 * loop over the vc clipped vertices (identical wherever the reference is defined:
 * a repeated closing vertex contributes cross(v,v).z = 0). */
float orc_calculate_ltc(uint32_t vertex_count, const float v[ORC_MAX_P][3]) {
	float result = 0.0f;
	for (uint32_t i = 0; i != vertex_count; ++i)
		result += integrate_edge_vec(ld3(v[i]), ld3(v[(i + 1) % vertex_count]));
	return fabsf(result);
}

/* polygon_clipping.glsl:19-25 */
static v3 iz0(v3 lhs, v3 rhs) {
	float lerp_factor = lhs.z / (lhs.z - rhs.z);
	return mk3(fmaf(lerp_factor, rhs.x, fmaf(-lerp_factor, lhs.x, lhs.x)),
	           fmaf(lerp_factor, rhs.y, fmaf(-lerp_factor, lhs.y, lhs.y)), 0.0f);
}

/* polygon_clipping.glsl:35-225, with the per-case slot order taken from
 * clip_rotation_table.h (derived from the reference's case list). */
uint32_t orc_clip_polygon(uint32_t vertex_count, float v[ORC_MAX_P][3], uint32_t min_vertices, uint32_t max_polygon_vertices) {
	uint32_t n = vertex_count;
	uint32_t mask = 0;
	for (uint32_t i = 0; i + 1 < max_polygon_vertices; ++i)
		if (v[i][2] > 0.0f && (i < min_vertices || i < vertex_count)) mask |= 1u << i;
	if (n < 3 || n > 7 || mask >= 128u) return 0;
	uint32_t rot = orc_clip_rotation[n - 3][mask];
	if (rot == 0xFF) return 0;
	v3 walk[ORC_MAX_P];
	uint32_t vc = 0;
	for (uint32_t i = 0; i != n; ++i) {
		uint32_t j = (i + 1) % n;
		uint32_t a = (mask >> i) & 1u, b = (mask >> j) & 1u;
		if (a) walk[vc++] = ld3(v[i]);
		if (a != b) walk[vc++] = iz0(ld3(v[i]), ld3(v[j]));
	}
	for (uint32_t j = 0; j != vc; ++j) {
		v3 s = walk[(j + rot) % vc];
		v[j][0] = s.x; v[j][1] = s.y; v[j][2] = s.z;
	}
	if (vc < ORC_MAX_P) { v[vc][0] = v[0][0]; v[vc][1] = v[0][1]; v[vc][2] = v[0][2]; }
	return vc;
}

/* polygon_sampling.glsl:545-613 */
void orc_prepare_psa(orc_psa_polygon_t* polygon, uint32_t vertex_count, const float vertices[ORC_MAX_P][3], uint32_t max_polygon_vertices, uint32_t fast_atan) {
	(void) max_polygon_vertices;
	memset(polygon, 0, sizeof(*polygon));
	uint32_t vc = vertex_count;
	polygon->vertex_count = vc;
	v2 inner_ellipse_0 = mk2(1.0f, 0.0f);
	v2 ellipses[ORC_MAX_P];
	polygon->vertices[0][0] = vertices[0][0]; polygon->vertices[0][1] = vertices[0][1];
	ellipses[0] = ellipse_from_edge(ld3(vertices[0]), ld3(vertices[1]));
	v2 previous_ellipse = ellipses[0];
	for (uint32_t i = 1; i != vc; ++i) {
		polygon->vertices[i][0] = vertices[i][0]; polygon->vertices[i][1] = vertices[i][1];
		v2 ellipse = ellipse_from_edge(ld3(vertices[i]), ld3(vertices[(i + 1) % vc]));
		int ellipse_inner = is_inner_ellipse(ellipse);
		ellipses[i] = ellipse_inner ? previous_ellipse : ellipse;
		inner_ellipse_0 = (is_inner_ellipse(previous_ellipse) && !ellipse_inner) ? previous_ellipse : inner_ellipse_0;
		previous_ellipse = ellipse;
	}
	{
		v2 ellipse = ellipses[0];
		int ellipse_inner = is_inner_ellipse(ellipse);
		ellipses[0] = ellipse_inner ? previous_ellipse : ellipse;
		inner_ellipse_0 = (is_inner_ellipse(previous_ellipse) && !ellipse_inner) ? previous_ellipse : inner_ellipse_0;
	}
	for (uint32_t i = 0; i != vc; ++i) { polygon->ellipses[i][0] = ellipses[i].x; polygon->ellipses[i][1] = ellipses[i].y; }
	polygon->inner_ellipse_0[0] = inner_ellipse_0.x; polygon->inner_ellipse_0[1] = inner_ellipse_0.y;
	polygon->projected_solid_angle = 0.0f;
	if (inner_ellipse_0.x > 0.0f) {
		/* central case: vc sectors, each bounded by one ellipse */
		for (uint32_t i = 0; i != vc; ++i) {
			uint32_t j = (i + 1) % vc;
			polygon->sector_projected_solid_angles[i] = ellipse_area_in_sector(
				mk2(polygon->ellipses[i][0], polygon->ellipses[i][1]),
				mk2(polygon->vertices[i][0], polygon->vertices[i][1]),
				mk2(polygon->vertices[j][0], polygon->vertices[j][1]), (int) fast_atan);
			polygon->projected_solid_angle += polygon->sector_projected_solid_angles[i];
		}
	}
	else {
		sort_convex_polygon_vertices(polygon);
		v2 inner_ellipse = inner_ellipse_0;
		float inner_rsqrt_det = get_ellipse_rsqrt_det(inner_ellipse);
		v2 outer_ellipse = mk2(0.0f, 0.0f);
		float outer_rsqrt_det = 0.0f;
		for (uint32_t i = 0; i + 1 != vc; ++i) {
			v2 vertex_ellipse = mk2(polygon->ellipses[i][0], polygon->ellipses[i][1]);
			int vertex_inner = is_inner_ellipse(vertex_ellipse);
			float vertex_rsqrt_det = get_ellipse_rsqrt_det(vertex_ellipse);
			if (i == 0) {
				outer_ellipse = vertex_ellipse;
				outer_rsqrt_det = vertex_rsqrt_det;
			}
			else {
				inner_ellipse = vertex_inner ? vertex_ellipse : inner_ellipse;
				inner_rsqrt_det = vertex_inner ? vertex_rsqrt_det : inner_rsqrt_det;
				outer_ellipse = vertex_inner ? outer_ellipse : vertex_ellipse;
				outer_rsqrt_det = vertex_inner ? outer_rsqrt_det : vertex_rsqrt_det;
			}
			polygon->sector_projected_solid_angles[i] = area_between_ellipses_in_sector(
				inner_ellipse, inner_rsqrt_det, outer_ellipse, outer_rsqrt_det,
				mk2(polygon->vertices[i][0], polygon->vertices[i][1]),
				mk2(polygon->vertices[i + 1][0], polygon->vertices[i + 1][1]), (int) fast_atan);
			polygon->projected_solid_angle += polygon->sector_projected_solid_angles[i];
		}
	}
}

/* polygon_sampling.glsl:622-634 */
static v2 normalize_approx_and_flip(v2 rhs, v2 semi_circle) {
	float scaling = fabsf(rhs.x) + fabsf(rhs.y);
	scaling = u2f(f2u(scaling) ^ 0x7F800000u);
	scaling = (dot2(rhs, semi_circle) >= 0.0f) ? scaling : -scaling;
	return mul2(rhs, scaling);
}

/* polygon_sampling.glsl:649-654. q is mat2 [col][row]. */
static v2 solve_homogeneous_quadratic(const float q[2][2]) {
	float coeff_xy = 0.5f * (q[0][1] + q[1][0]);
	float sqrt_discriminant = sqrtf(fmaxf(0.0f, coeff_xy * coeff_xy - q[0][0] * q[1][1]));
	float scaled_root = fabsf(coeff_xy) + sqrt_discriminant;
	return (coeff_xy >= 0.0f) ? mk2(scaled_root, -q[0][0]) : mk2(q[1][1], scaled_root);
}

/* outerProduct(c, r)[col j][row i] = c[i] * r[j] */
static inline void outer2(float out[2][2], v2 c, v2 r) {
	out[0][0] = c.x * r.x; out[0][1] = c.y * r.x;
	out[1][0] = c.x * r.y; out[1][1] = c.y * r.y;
}

/* polygon_sampling.glsl:668-762 */
static v2 sample_sector_between_ellipses(v2 random_numbers, float target_area, v2 inner_ellipse, v2 outer_ellipse, v2 dir_0, v2 dir_1, uint32_t iteration_count, int fast_atan, int biased) {
	v2 quad_dirs[3];
	quad_dirs[0] = normalize2(dir_0);
	quad_dirs[2] = normalize2(dir_1);
	quad_dirs[1] = add2(quad_dirs[0], quad_dirs[2]);
	float nf[2][3] = {
		{ get_ellipse_normalized_direction_factor(inner_ellipse, quad_dirs[0]),
		  get_ellipse_direction_factor(inner_ellipse, quad_dirs[1]),
		  get_ellipse_normalized_direction_factor(inner_ellipse, quad_dirs[2]) },
		{ get_ellipse_normalized_direction_factor(outer_ellipse, quad_dirs[0]),
		  get_ellipse_direction_factor(outer_ellipse, quad_dirs[1]),
		  get_ellipse_normalized_direction_factor(outer_ellipse, quad_dirs[2]) }
	};
	float sector_areas[2] = {
		nf[1][0] * nf[1][1] - nf[0][0] * nf[0][1],
		nf[1][1] * nf[1][2] - nf[0][1] * nf[0][2]
	};
	float target_quad_area = mix_fma(-sector_areas[0], sector_areas[1], random_numbers.x);
	int first = (target_quad_area <= 0.0f);
	quad_dirs[2] = first ? quad_dirs[0] : quad_dirs[2];
	nf[0][2] = first ? nf[0][0] : nf[0][2];
	nf[1][2] = first ? nf[1][0] : nf[1][2];
	target_quad_area += first ? sector_areas[0] : -sector_areas[1];
	/* determinant(mat2(c0, c1)) = c0.x * c1.y - c1.x * c0.y */
	target_quad_area *= fabsf(quad_dirs[1].x * quad_dirs[2].y - quad_dirs[2].x * quad_dirs[1].y);
	v2 quad_normals[2] = {
		add2(mul2(quad_dirs[1], nf[0][1]), mul2(quad_dirs[2], nf[0][2])),
		add2(mul2(quad_dirs[1], nf[1][1]), mul2(quad_dirs[2], nf[1][2]))
	};
	quad_normals[0] = ellipse_transform(inner_ellipse, quad_normals[0]);
	quad_normals[1] = ellipse_transform(outer_ellipse, quad_normals[1]);
	float quad_offsets[2] = {
		dot2(quad_normals[0], quad_dirs[1]) * nf[0][1],
		dot2(quad_normals[1], quad_dirs[1]) * nf[1][1]
	};
	float quadratic[2][2], tmp[2][2];
	outer2(quadratic, mul2(rotate_90(quad_dirs[2]), quad_offsets[1] * nf[1][2]), quad_normals[0]);
	outer2(tmp, add2(mul2(rotate_90(quad_dirs[2]), quad_offsets[0] * nf[0][2]), mul2(quad_normals[0], target_quad_area)), quad_normals[1]);
	for (int c = 0; c != 2; ++c) for (int r = 0; r != 2; ++r) quadratic[c][r] -= tmp[c][r];
	v2 current_dir = solve_homogeneous_quadratic(quadratic);
	if (!biased) {
		float acceptable_error = 1.0e-5f;
		iteration_count = (fabsf(random_numbers.x - 0.5f) <= 0.5f - acceptable_error) ? iteration_count : 0;
		float inner_rsqrt_det = get_ellipse_rsqrt_det(inner_ellipse);
		float outer_rsqrt_det = get_ellipse_rsqrt_det(outer_ellipse);
		for (uint32_t i = 0; i != iteration_count; ++i) {
			current_dir = normalize_approx_and_flip(current_dir, quad_dirs[1]);
			v2 inner_dir = ellipse_transform(inner_ellipse, current_dir);
			v2 outer_dir = ellipse_transform(outer_ellipse, current_dir);
			float det_dirs = fmaxf(+0.0f, dot2(current_dir, rotate_90(quad_dirs[0])));
			float error = target_area - area_between_ellipses_from_tangents(
				inner_rsqrt_det, det_dirs / (inner_rsqrt_det * dot2(quad_dirs[0], inner_dir)),
				outer_rsqrt_det, det_dirs / (outer_rsqrt_det * dot2(quad_dirs[0], outer_dir)), fast_atan);
			outer2(quadratic, sub2(inner_dir, outer_dir), rotate_90(current_dir));
			outer2(tmp, mul2(inner_dir, 2.0f * error), outer_dir);
			for (int c = 0; c != 2; ++c) for (int r = 0; r != 2; ++r) quadratic[c][r] -= tmp[c][r];
			current_dir = solve_homogeneous_quadratic(quadratic);
		}
	}
	current_dir = (dot2(current_dir, quad_dirs[1]) >= 0.0f) ? current_dir : mk2(-current_dir.x, -current_dir.y);
	float inner_factor = 1.0f / get_ellipse_direction_factor_rsq(inner_ellipse, current_dir);
	float outer_factor = 1.0f / get_ellipse_direction_factor_rsq(outer_ellipse, current_dir);
	return mul2(current_dir, sqrtf(mix_fma(inner_factor, outer_factor, random_numbers.y)));
}

/* polygon_sampling.glsl:772-828 */
void orc_sample_psa(float out_dir[3], const orc_psa_polygon_t* polygon, float u0, float u1, uint32_t max_polygon_vertices, uint32_t fast_atan, uint32_t biased) {
	uint32_t P = max_polygon_vertices;
	uint32_t vc = polygon->vertex_count;
	float target = u0 * polygon->projected_solid_angle;
	v2 sampled = mk2(0.0f, 0.0f);
	v2 outer_ellipse = mk2(0.0f, 0.0f), dir_0 = mk2(0.0f, 0.0f);
	if (polygon->inner_ellipse_0[0] > 0.0f) {
		for (uint32_t i = 0; i != P; ++i) {
			if (i > 0) target -= polygon->sector_projected_solid_angles[i - 1];
			outer_ellipse = mk2(polygon->ellipses[i][0], polygon->ellipses[i][1]);
			dir_0 = mk2(polygon->vertices[i][0], polygon->vertices[i][1]);
			if ((i >= 2 && i + 1 == vc) || target < polygon->sector_projected_solid_angles[i]) break;
		}
		float sqrt_det = sqrtf(get_ellipse_det(outer_ellipse));
		float angle = 2.0f * target * sqrt_det;
		v2 t = rotate_90(ellipse_transform(outer_ellipse, dir_0));
		float ca = glsl_cos(angle) * sqrt_det, sa = glsl_sin(angle);
		sampled = mk2(ca * dir_0.x + sa * t.x, ca * dir_0.y + sa * t.y);
		float s = sqrtf(u1 / get_ellipse_direction_factor_rsq(outer_ellipse, sampled));
		sampled = mul2(sampled, s);
	}
	else {
		float sector_psa = 0.0f;
		v2 inner_ellipse = mk2(polygon->inner_ellipse_0[0], polygon->inner_ellipse_0[1]);
		v2 dir_1 = mk2(0.0f, 0.0f);
		for (uint32_t i = 0; i + 1 != P; ++i) {
			v2 vertex_ellipse = mk2(polygon->ellipses[i][0], polygon->ellipses[i][1]);
			if (i == 0) outer_ellipse = vertex_ellipse;
			else {
				target -= polygon->sector_projected_solid_angles[i - 1];
				int vertex_inner = is_inner_ellipse(vertex_ellipse);
				inner_ellipse = vertex_inner ? vertex_ellipse : inner_ellipse;
				outer_ellipse = vertex_inner ? outer_ellipse : vertex_ellipse;
			}
			dir_0 = mk2(polygon->vertices[i][0], polygon->vertices[i][1]);
			dir_1 = mk2(polygon->vertices[i + 1][0], polygon->vertices[i + 1][1]);
			sector_psa = polygon->sector_projected_solid_angles[i];
			if ((i >= 1 && i + 2 == vc) || target < sector_psa) break;
		}
		v2 rn = mk2(target / sector_psa, u1);
		sampled = sample_sector_between_ellipses(rn, target, inner_ellipse, outer_ellipse, dir_0, dir_1, 2, (int) fast_atan, (int) biased);
	}
	out_dir[0] = sampled.x;
	out_dir[1] = sampled.y;
	out_dir[2] = sqrtf(fmaxf(0.0f, fmaf(-sampled.x, sampled.x, fmaf(-sampled.y, sampled.y, 1.0f))));
}

/* ------------------------------------------------------------------- LTC */

/* ltc_table.c:82-116 */
void orc_quantize_ltc_fit(const float d[5], uint16_t rgba[4], uint16_t rg[2]) {
	float inverse[3][3] = {
		{ d[2], 0.0f, -d[1] * d[2] },
		{ 0.0f, d[0] - d[1] * d[3], 0.0f },
		{ -d[2] * d[3], 0.0f, d[0] * d[2] }
	};
	float max_magnitude = fabsf(inverse[0][0]);
	for (int k = 0; k != 3; ++k) for (int l = 0; l != 3; ++l)
		if (max_magnitude < fabsf(inverse[k][l])) max_magnitude = fabsf(inverse[k][l]);
	for (int k = 0; k != 3; ++k) for (int l = 0; l != 3; ++l) inverse[k][l] /= max_magnitude;
	float processed[6] = { inverse[0][0], inverse[0][2], inverse[1][1], inverse[2][0], inverse[2][2], d[4] };
	for (int i = 0; i != 6; ++i) {
		float x = processed[i];
		x *= (i == 1) ? -1.0f : 1.0f;
		if (x < 0.0f) x = 0.0f;
		if (x > 1.0f) x = 1.0f;
		uint16_t q = (uint16_t) (x * 65535.0f + 0.5f);
		if (i < 4) rgba[i] = q; else rg[i - 4] = q;
	}
}

/* ltc_table.c:184-191 */
void orc_ltc_constants(float out[8], uint32_t roughness_count, uint32_t inclination_count, uint32_t fresnel_count) {
	out[0] = (float) (fresnel_count - 1);
	out[1] = 0.0f;
	out[2] = (float) (roughness_count - 1) / (float) roughness_count;
	out[3] = 0.5f / (float) roughness_count;
	out[4] = (float) (inclination_count - 1) / (0.5f * ORC_PI * inclination_count);
	out[5] = 0.5f / (float) inclination_count;
	out[6] = out[7] = 0.0f;
}

/* Texture-unit stand-in for textureLod(sampler2DArray, ...) with the sampler of
 * ltc_table.c:170-177 (linear, clamp to edge, one mip): layer = nearest, exact
 * fp32 bilinear weights. Driver behaviour, defined here ("parity unpinned"). */
static void ltc_fetch(const orc_scene_t* scene, float u, float v, float layer_coord, float out[6]) {
	int res = (int) scene->ltc_res;
	int layer = (int) floorf(layer_coord + 0.5f);
	if (layer < 0) layer = 0;
	if (layer > (int) scene->ltc_layers - 1) layer = (int) scene->ltc_layers - 1;
	float x = u * (float) res - 0.5f;
	float y = v * (float) res - 0.5f;
	float fx0 = floorf(x), fy0 = floorf(y);
	float fx = x - fx0, fy = y - fy0;
	int x0 = (int) fx0, y0 = (int) fy0, x1 = x0 + 1, y1 = y0 + 1;
	if (x0 < 0) x0 = 0; if (x0 > res - 1) x0 = res - 1;
	if (x1 < 0) x1 = 0; if (x1 > res - 1) x1 = res - 1;
	if (y0 < 0) y0 = 0; if (y0 > res - 1) y0 = res - 1;
	if (y1 < 0) y1 = 0; if (y1 > res - 1) y1 = res - 1;
	float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
	size_t base = (size_t) layer * res * res;
	size_t i00 = base + (size_t) y0 * res + x0, i10 = base + (size_t) y0 * res + x1;
	size_t i01 = base + (size_t) y1 * res + x0, i11 = base + (size_t) y1 * res + x1;
	for (int c = 0; c != 4; ++c)
		out[c] = w00 * ((float) scene->ltc_rgba16[i00 * 4 + c] / 65535.0f) + w10 * ((float) scene->ltc_rgba16[i10 * 4 + c] / 65535.0f)
		       + w01 * ((float) scene->ltc_rgba16[i01 * 4 + c] / 65535.0f) + w11 * ((float) scene->ltc_rgba16[i11 * 4 + c] / 65535.0f);
	for (int c = 0; c != 2; ++c)
		out[4 + c] = w00 * ((float) scene->ltc_rg16[i00 * 2 + c] / 65535.0f) + w10 * ((float) scene->ltc_rg16[i10 * 2 + c] / 65535.0f)
		           + w01 * ((float) scene->ltc_rg16[i01 * 2 + c] / 65535.0f) + w11 * ((float) scene->ltc_rg16[i11 * 2 + c] / 65535.0f);
}

/* ltc_utility.glsl:56-88 */
void orc_get_ltc_coefficients(orc_ltc_t* ltc, const orc_scene_t* scene, float fresnel_0, float roughness,
	const float position[3], const float normal_[3], const float outgoing_[3], const float c[6])
{
	v3 normal = ld3(normal_), outgoing = ld3(outgoing_), pos = ld3(position);
	float normal_dot_outgoing = dot3(normal, outgoing);
	float inclination = glsl_acos(clampf(normal_dot_outgoing, 0.0f, 1.0f));
	float tu = fmaf(sqrtf(clampf(roughness, 0.0f, 1.0f)), c[2], c[3]);
	float tv = fmaf(inclination, c[4], c[5]);
	float tl = fmaf(clampf(fresnel_0, 0.0f, 1.0f), c[0], c[1]);
	float d[6];
	ltc_fetch(scene, tu, tv, tl, d);
	float d0x = d[0], d0y = d[1], d0z = d[2], d0w = d[3], d1x = d[4], d1y = d[5];
	/* mat3(a,b,c, d,e,f, g,h,i) fills columns */
	float s2c[3][3] = { { d0x, 0.0f, -d0y }, { 0.0f, d0z, 0.0f }, { d0w, 0.0f, d1x } };
	memcpy(ltc->shading_to_cosine, s2c, sizeof(s2c));
	ltc->albedo = d1y;
	float determinant_2x2 = d0x * d1x + d0y * d0w;
	ltc->determinant = d0z * determinant_2x2;
	float inv_determinant_2x2 = 1.0f / determinant_2x2;
	float c2s[3][3] = {
		{ d1x * inv_determinant_2x2, 0.0f, d0y * inv_determinant_2x2 },
		{ 0.0f, 1.0f / d0z, 0.0f },
		{ -d0w * inv_determinant_2x2, 0.0f, d0x * inv_determinant_2x2 } };
	memcpy(ltc->cosine_to_shading, c2s, sizeof(c2s));
	v3 x_axis = normalize3(mk3(fmaf(-normal_dot_outgoing, normal.x, outgoing.x), fmaf(-normal_dot_outgoing, normal.y, outgoing.y), fmaf(-normal_dot_outgoing, normal.z, outgoing.z)));
	v3 y_axis = cross3(normal, x_axis);
	/* rotation = transpose(mat3(x_axis, y_axis, normal)): column j = (x[j], y[j], n[j]) */
	float rot[3][3] = { { x_axis.x, y_axis.x, normal.x }, { x_axis.y, y_axis.y, normal.y }, { x_axis.z, y_axis.z, normal.z } };
	for (int j = 0; j != 3; ++j) for (int i = 0; i != 3; ++i) ltc->world_to_shading[j][i] = rot[j][i];
	/* -rotation * position */
	float neg[3][3];
	for (int j = 0; j != 3; ++j) for (int i = 0; i != 3; ++i) neg[j][i] = -rot[j][i];
	v3 t = m3_mul(neg, pos);
	ltc->world_to_shading[3][0] = t.x; ltc->world_to_shading[3][1] = t.y; ltc->world_to_shading[3][2] = t.z;
}

/* ltc_utility.glsl:100-105 */
static float evaluate_ltc_density(const orc_ltc_t* ltc, v3 dir_shading_space, float rcp_projected_solid_angle) {
	v3 dc = m3_mul(ltc->shading_to_cosine, dir_shading_space);
	float l2 = dot3(dc, dc);
	float density = fmaxf(0.0f, dc.z) * ltc->determinant / (l2 * l2);
	return density * rcp_projected_solid_angle;
}

/* ---------------------------------------------------------- host arithmetic */

/* polygonal_light.c:44-98 */
void orc_update_polygonal_light(const float angles[3], float scaling_x, float scaling_y, const float translation[3],
	const float radiant_flux[3], uint32_t vertex_count, const float* plane_space,
	float* world, float plane[4], float surface_radiance[3], float* area, float rotation_out[3][4])
{
	float cx = cosf(angles[0]), sx = sinf(angles[0]);
	float cy = cosf(angles[1]), sy = sinf(angles[1]);
	float cz = cosf(angles[2]), sz = sinf(angles[2]);
	float cxsy = cx * sy, sxsy = sx * sy;
	float rotation[3][4] = {
		{ cy * cz, -cy * sz, -sy, 0.0f },
		{ -sxsy * cz + cx * sz, sxsy * sz + cx * cz, -sx * cy, 0.0f },
		{ cxsy * cz + sx * sz, -cxsy * sz + sx * cz, cx * cy, 0.0f },
	};
	if (rotation_out) memcpy(rotation_out, rotation, sizeof(rotation));
	float scalings[2] = { scaling_x, scaling_y };
	for (uint32_t i = 0; i != vertex_count; ++i)
		for (uint32_t j = 0; j != 3; ++j) {
			world[i * 4 + j] = translation[j];
			for (uint32_t k = 0; k != 2; ++k)
				world[i * 4 + j] += scalings[k] * rotation[j][k] * plane_space[i * 4 + k];
		}
	plane[0] = rotation[0][2]; plane[1] = rotation[1][2]; plane[2] = rotation[2][2];
	plane[3] = -(rotation[0][2] * translation[0] + rotation[1][2] * translation[1] + rotation[2][2] * translation[2]);
	float signed_area = 0.0f;
	for (uint32_t i = 0; i + 2 != vertex_count; ++i) {
		float m00 = plane_space[(i + 2) * 4 + 0] - plane_space[0], m01 = plane_space[(i + 1) * 4 + 0] - plane_space[0];
		float m10 = plane_space[(i + 2) * 4 + 1] - plane_space[1], m11 = plane_space[(i + 1) * 4 + 1] - plane_space[1];
		signed_area += 0.5f * (m00 * m11 - m01 * m10);
	}
	signed_area *= scalings[0] * scalings[1];
	*area = (signed_area < 0.0f) ? -signed_area : signed_area;
	for (int i = 0; i != 3; ++i) surface_radiance[i] = radiant_flux[i];
	for (int i = 0; i != 4; ++i) plane[i] = (signed_area > 0.0f) ? plane[i] : (-plane[i]);
}

/* camera.c:24-83 */
void orc_world_to_projection(float out[4][4], const float position[3], float rotation_x_angle, float rotation_z_angle, float vertical_fov, float near_plane, float far_plane, float aspect) {
	float cos_x = cosf(rotation_x_angle), sin_x = sinf(rotation_x_angle);
	float cos_z = cosf(rotation_z_angle), sin_z = sinf(rotation_z_angle);
	float rx[3][3] = { { 1.0f, 0.0f, 0.0f }, { 0.0f, cos_x, sin_x }, { 0.0f, -sin_x, cos_x } };
	float rz[3][3] = { { cos_z, sin_z, 0.0f }, { -sin_z, cos_z, 0.0f }, { 0.0f, 0.0f, 1.0f } };
	float rotation[3][3] = { { 0.0f } };
	for (int i = 0; i != 3; ++i) for (int j = 0; j != 3; ++j) for (int l = 0; l != 3; ++l)
		rotation[i][j] += rz[i][l] * rx[l][j];
	float origin_view[3] = { 0.0f, 0.0f, 0.0f };
	for (int i = 0; i != 3; ++i) for (int j = 0; j != 3; ++j)
		origin_view[i] -= rotation[j][i] * position[j];
	float w2v[4][4] = {
		{ rotation[0][0], rotation[1][0], rotation[2][0], origin_view[0] },
		{ rotation[0][1], rotation[1][1], rotation[2][1], origin_view[1] },
		{ rotation[0][2], rotation[1][2], rotation[2][2], origin_view[2] },
		{ 0.0f, 0.0f, 0.0f, 1.0f } };
	float top = tanf(0.5f * vertical_fov);
	float right = aspect * top;
	float v2p[4][4] = {
		{ -1.0f / right, 0.0f, 0.0f, 0.0f },
		{ 0.0f, 1.0f / top, 0.0f, 0.0f },
		{ 0.0f, 0.0f, -(far_plane + near_plane) / (far_plane - near_plane), -2.0f * far_plane * near_plane / (far_plane - near_plane) },
		{ 0.0f, 0.0f, -1.0f, 0.0f } };
	memset(out, 0, sizeof(float) * 16);
	for (int i = 0; i != 4; ++i) for (int j = 0; j != 4; ++j) for (int l = 0; l != 4; ++l)
		out[i][j] += v2p[i][l] * w2v[l][j];
}

/* math_utilities.h:24-47 -- cofactor expansion, term order as in the reference
 * (the order of the six products per entry fixes the fp32 result). */
static void matrix_inverse(float inverse[4][4], const float matrix[4][4]) {
	float* inv = &inverse[0][0];
	const float* m = &matrix[0][0];
	/* rows of the adjugate, each "a*b*c" triple written as index triples with signs */
	static const signed char T[16][6][4] = {
		/* inv[0] */ { {1,5,10,15},{-1,5,11,14},{-1,9,6,15},{1,9,7,14},{1,13,6,11},{-1,13,7,10} },
		/* inv[1] */ { {-1,1,10,15},{1,1,11,14},{1,9,2,15},{-1,9,3,14},{-1,13,2,11},{1,13,3,10} },
		/* inv[2] */ { {1,1,6,15},{-1,1,7,14},{-1,5,2,15},{1,5,3,14},{1,13,2,7},{-1,13,3,6} },
		/* inv[3] */ { {-1,1,6,11},{1,1,7,10},{1,5,2,11},{-1,5,3,10},{-1,9,2,7},{1,9,3,6} },
		/* inv[4] */ { {-1,4,10,15},{1,4,11,14},{1,8,6,15},{-1,8,7,14},{-1,12,6,11},{1,12,7,10} },
		/* inv[5] */ { {1,0,10,15},{-1,0,11,14},{-1,8,2,15},{1,8,3,14},{1,12,2,11},{-1,12,3,10} },
		/* inv[6] */ { {-1,0,6,15},{1,0,7,14},{1,4,2,15},{-1,4,3,14},{-1,12,2,7},{1,12,3,6} },
		/* inv[7] */ { {1,0,6,11},{-1,0,7,10},{-1,4,2,11},{1,4,3,10},{1,8,2,7},{-1,8,3,6} },
		/* inv[8] */ { {1,4,9,15},{-1,4,11,13},{-1,8,5,15},{1,8,7,13},{1,12,5,11},{-1,12,7,9} },
		/* inv[9] */ { {-1,0,9,15},{1,0,11,13},{1,8,1,15},{-1,8,3,13},{-1,12,1,11},{1,12,3,9} },
		/* inv[10]*/ { {1,0,5,15},{-1,0,7,13},{-1,4,1,15},{1,4,3,13},{1,12,1,7},{-1,12,3,5} },
		/* inv[11]*/ { {-1,0,5,11},{1,0,7,9},{1,4,1,11},{-1,4,3,9},{-1,8,1,7},{1,8,3,5} },
		/* inv[12]*/ { {-1,4,9,14},{1,4,10,13},{1,8,5,14},{-1,8,6,13},{-1,12,5,10},{1,12,6,9} },
		/* inv[13]*/ { {1,0,9,14},{-1,0,10,13},{-1,8,1,14},{1,8,2,13},{1,12,1,10},{-1,12,2,9} },
		/* inv[14]*/ { {-1,0,5,14},{1,0,6,13},{1,4,1,14},{-1,4,2,13},{-1,12,1,6},{1,12,2,5} },
		/* inv[15]*/ { {1,0,5,10},{-1,0,6,9},{-1,4,1,10},{1,4,2,9},{1,8,1,6},{-1,8,2,5} },
	};
	for (int e = 0; e != 16; ++e) {
		float acc = 0.0f;
		for (int t = 0; t != 6; ++t) {
			float a = m[T[e][t][1]];
			if (T[e][t][0] < 0) a = -a;
			float term = a * m[T[e][t][2]] * m[T[e][t][3]];
			acc = (t == 0) ? term : acc + term;
		}
		inv[e] = acc;
	}
	float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
	float rcp_det = 1.0f / det;
	for (int i = 0; i != 16; ++i) inv[i] = inv[i] * rcp_det;
}

/* main.c:2920-2944 */
void orc_pixel_to_ray(float out[3][4], const float world_to_projection[4][4], uint32_t width, uint32_t height) {
	float vt[4];
	vt[0] = 2.0f / width;
	vt[1] = 2.0f / height;
	vt[2] = 0.5f * vt[0] - 1.0f;
	vt[3] = 0.5f * vt[1] - 1.0f;
	float p2w[4][4], w2p[4][4];
	memcpy(w2p, world_to_projection, sizeof(w2p));
	w2p[0][3] = 0.0f; w2p[1][3] = 0.0f; w2p[2][3] = 0.0f;
	matrix_inverse(p2w, w2p);
	float p2r[4][3] = { { vt[0], 0.0f, vt[2] }, { 0.0f, vt[1], vt[3] }, { 0.0f, 0.0f, 1.0f }, { 0.0f, 0.0f, 1.0f } };
	memset(out, 0, sizeof(float) * 12);
	for (int i = 0; i != 3; ++i) for (int j = 0; j != 3; ++j) for (int k = 0; k != 4; ++k)
		out[i][j] += p2w[i][k] * p2r[k][j];
}

/* ------------------------------------------------------------ mesh decode */

/* mesh_quantization.glsl:38-45 (fma form, used by the shaders) */
static v3 decode_position_64_bit(uint32_t q0, uint32_t q1, const float factor[3], const float summand[3]) {
	float px = (float) (q0 & 0x1FFFFFu);
	float py = (float) (((q0 & 0xFFE00000u) >> 21) | ((q1 & 0x3FFu) << 11));
	float pz = (float) ((q1 & 0x7FFFFC00u) >> 10);
	return mk3(fmaf(px, factor[0], summand[0]), fmaf(py, factor[1], summand[1]), fmaf(pz, factor[2], summand[2]));
}

/* scene.c:176-187 (mul + add form, used for the acceleration structure) */
static v3 dequantize_for_bvh(uint32_t q0, uint32_t q1, const float factor[3], const float summand[3]) {
	float px = (float) (q0 & 0x1FFFFFu);
	float py = (float) (((q0 & 0xFFE00000u) >> 21) | ((q1 & 0x3FFu) << 11));
	float pz = (float) ((q1 & 0x7FFFFC00u) >> 10);
	return mk3(px * factor[0] + summand[0], py * factor[1] + summand[1], pz * factor[2] + summand[2]);
}

/* mesh_quantization.glsl:19-33 */
static v3 decode_normal_32_bit(float ox, float oy) {
	const float factor = 2.0f * (65534.0f / 65535.0f);
	const float summand = -(32768.0f / 65535.0f) * factor;
	ox = fmaf(ox, factor, summand);
	oy = fmaf(oy, factor, summand);
	v3 normal = mk3(ox, oy, 1.0f - fabsf(ox) - fabsf(oy));
	float sx = (ox >= 0.0f) ? 1.0f : -1.0f, sy = (oy >= 0.0f) ? 1.0f : -1.0f;
	if (normal.z < 0.0f) {
		float nx = (1.0f - fabsf(normal.y)) * sx;
		float ny = (1.0f - fabsf(normal.x)) * sy;
		normal.x = nx; normal.y = ny;
	}
	return normalize3(normal);
}


/* BVH stand-in, shading-point reconstruction and the three passes (unity build) */
#include "risltc_oracle_frame.inc"

/* ----------------------------------------------------------------- copy pass */

/* packHalf2x16 for one value: fp32 -> fp16 bits, round to nearest even */
static uint16_t half_bits_of(float value) {
	uint32_t x; memcpy(&x, &value, 4);
	uint32_t sign = (x >> 16) & 0x8000u, mantissa = x & 0x7FFFFFu, biased = (x >> 23) & 0xFFu;
	int32_t exponent = (int32_t) biased - 127 + 15;
	if (biased == 0xFFu) return (uint16_t) (sign | 0x7C00u | (mantissa ? 0x200u : 0u));
	if (exponent >= 31) return (uint16_t) (sign | 0x7C00u);
	if (exponent <= 0) {
		if (exponent < -10) return (uint16_t) sign;
		mantissa |= 0x800000u;
		uint32_t shift = (uint32_t) (14 - exponent);
		uint32_t half = mantissa >> shift, rest = mantissa & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
		if (rest > halfway || (rest == halfway && (half & 1u))) ++half;
		return (uint16_t) (sign | half);
	}
	uint32_t half = ((uint32_t) exponent << 10) | (mantissa >> 13), rest = mantissa & 0x1FFFu;
	if (rest > 0x1000u || (rest == 0x1000u && (half & 1u))) ++half;   /* may carry into the exponent: correct */
	return (uint16_t) (sign | half);
}

/* copy_pass.frag.glsl:28-58 + srgb_utility.glsl:20-34 + the UNORM8 write of the swapchain image (round to nearest).
 * frame_bits 0: display (linear -> sRGB), 1 / 2: low / high byte of the half bits of r, g, b. pow() as glsl_atan above. */
void orc_copy_pass(const float* rgba, uint64_t pixel_count, uint32_t frame_bits, uint8_t* rgb8) {
	for (uint64_t i = 0; i != pixel_count; ++i)
		for (int k = 0; k != 3; ++k) {
			float v = rgba[4 * i + k];
			uint32_t byte;
			if (frame_bits != 0u) {
				uint32_t h = half_bits_of(v);
				byte = (frame_bits == 1u) ? (h & 0xFFu) : (h >> 8);
			}
			else {
				float linear = clampf(v, 0.0f, 1.0f);
				float srgb = (linear <= 0.0031308f) ? (12.92f * linear) : (1.055f * (float) pow((double) linear, (double) (1.0f / 2.4f)) - 0.055f);
				byte = (uint32_t) (clampf(srgb, 0.0f, 1.0f) * 255.0f + 0.5f);
			}
			rgb8[3 * i + k] = (uint8_t) byte;
		}
}
