// common.cuh -- shared types and small math helpers of the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

// One translation unit, compiled WITHOUT a*b+c contraction and with IEEE division / square root
// (risltc_b200/build.py), so that code written with plain operators rounds exactly like the C oracle.
// Code that may round differently says so explicitly: fmaf(), approx_rsqrt(), approx_rcp(), approx_sqrt().
#ifndef RL_NS
#define RL_NS exact
#endif

#define RL_MAX_P 8            // largest MAX_POLYGON_VERTEX_COUNT (V_max 7 + 1, main.c:191-204)
#define RL_PI 3.14159265358979323846f
#define RL_HALF_PI 1.57079632679489661923f
#define RL_INV_PI 0.318309886183790671538f

// ---- per-frame uniforms: the fields of per_frame_constants_t that the hot path reads
// (main.h:537-553 / shared_constants.glsl:21-60), repacked for a kernel parameter.
struct FrameUniforms {
	float dequant_factor[3], dequant_summand[3];
	float world_to_projection[4][4];
	float pixel_to_ray[3][4];
	float camera[3];
	float mis_visibility_estimate;
	uint32_t width, height;
	float exposure, roughness_factor;
	uint32_t frame_word;      // g_noise_random_numbers.x, the only word the shader reads (noise_utility.glsl:82)
	float ltc_constants[6];
	uint32_t accum_num;
};

// ---- shader variant (the -D table of main.c:962-991)
struct Variant {
	uint32_t light_sampling, polygon_technique, mis_heuristic, sample_count, light_samples, fast_atan;
	uint32_t min_light_vertices, max_light_vertices;
};
enum { TECH_BASELINE = 0, TECH_TURK = 1, TECH_PSA = 2, TECH_PSA_BIASED = 3, TECH_LTC_CP = 4 };
enum { MIS_BALANCE = 0, MIS_POWER = 1, MIS_WEIGHTED = 2, MIS_OPTIMAL_CLAMPED = 3, MIS_OPTIMAL = 4 };

// ---- image partition: this device owns the rows y with (y / stripe_h) % stripe_count == stripe_index
struct Stripes {
	uint32_t stripe_h, stripe_index, stripe_count, owned_rows;
	uint32_t h_shift;   // log2(stripe_h) if it is a power of two, 0xFFFFFFFF otherwise (set by risltc_cuda_resize)
	// Called per pixel by the rasteriser and the resolve kernel: the two integer divisions by run-time values were 26 % of
	// raster_tiles_kernel's instructions (ncu source view, C3, round 2) -- for the identity when one device owns every row.
	__host__ __device__ uint32_t global_row(uint32_t local_row) const {
		if (stripe_count == 1u) return local_row;
		if (h_shift != 0xFFFFFFFFu) return ((((local_row >> h_shift) * stripe_count + stripe_index) << h_shift) | (local_row & (stripe_h - 1u)));
		return ((local_row / stripe_h) * stripe_count + stripe_index) * stripe_h + local_row % stripe_h;
	}
};

// ---- BVH: binary tree, one 64-byte node holds the boxes of BOTH children (4 x 16-byte loads)
struct __align__(16) BvhNode {
	float4 a;   // left.lo.xyz, left.hi.x
	float4 b;   // left.hi.yz, right.lo.xy
	float4 c;   // right.lo.z, right.hi.xyz
	int4 d;     // left child, right child (>= 0 inner node, < 0 leaf: ~((first << 4) | (count - 1))), unused x2
};
// ---- triangle: v0, e1 = v1 - v0, e2 = v2 - v0 (fp32 differences, identical to computing them per test);
// .w of the first vector carries the primitive id (bit 31 = emitter flag).
struct __align__(16) BvhTri { float4 v0, e1, e2; };

// ---- 4-wide BVH for the any-hit kernel: 64 bytes hold FOUR children. Boxes are quantised to 8 bits per plane relative
// to the node's own box (origin + q * 2^e per axis, rounded outwards at build time), which halves the bytes -- and the
// L1 tag lookups, the measured limiter of the binary layout -- per child and halves the node visits per ray.
struct __align__(16) Qbvh4Node {
	uint4 a;    // origin x, y, z; w = 2^15 * step of the x planes (all float bits; steps are powers of two)
	uint4 b;    // lo.x, lo.y, lo.z, hi.x: one byte per child each
	uint4 c;    // hi.y, hi.z; 2^15 * step of the y planes, of the z planes (float bits)
	int4 refs;  // child references (>= 0 node, < 0 leaf as in BvhNode, RL_Q4_EMPTY = no child)
};
#define RL_Q4_EMPTY 0x7FFFFFFF

// ---- material textures (scene.h:104-118: base colour, specular, normal per material), every mip level decoded to RGBA texels
#define RL_TEXEL_RGBA32F 0u
#define RL_TEXEL_RGBA8_UNORM 1u
#define RL_TEXEL_RGBA8_SRGB 2u
struct __align__(16) TextureDesc {
	unsigned long long offset;   // of level 0 in SceneView::texels, in bytes; the smaller levels follow tightly packed
	uint32_t format, width, height, levels;
	uint32_t pad[2];
};

struct SceneView {
	const uint2* positions;          // T*3 quantised positions (mesh_t.positions)
	const ushort4* normals_uv;       // T*3
	const uint8_t* material_indices; // T
	const float4* materials;         // 2 float4 per material: {base rgb, occlusion}, {roughness_lin, metalicity, normal.r, normal.g}
	const float4* lights;            // light records, (3 + V) float4 each
	const float4* lights_tri;        // triangle lights repacked as 3 float4 {v0 | Le.r, v1 | Le.g, v2 | Le.b}, or null
	uint32_t light_count, light_stride4;
	const ushort4* ltc_rgba; const ushort2* ltc_rg; uint32_t ltc_res, ltc_layers;
	const BvhNode* nodes; const BvhTri* tris; uint32_t triangle_count;
	const Qbvh4Node* nodes4;         // the same tree collapsed to four children per node (shadow rays)
	const TextureDesc* textures;     // 3 per material, or null: flat materials (the `materials` constants)
	const unsigned char* texels;
	const float* srgb_table;         // 256 entries: sRGB byte -> linear
};

struct PixelBuffers {
	uint32_t* visibility;   // [owned_rows * W] primitive id | emitter << 31, 0xFFFFFFFF = background
	float4* origin;         // [pixels] shading position, .w = bits of (number of light-sample groups)
	float4* base;           // [pixels] colour that needs no ray (background, emitters, inline variants)
	float4* group;          // [L][pixels] {carry rgb, scale}: sum of terms that need no ray, factor W (or N) of the group
	float4* ray_a;          // [L*S*2][pixels] {dir xyz, t_max}
	float4* ray_b;          // [L*S*2][pixels] {term rgb, valid}
	uint4* pick;            // [pixels] RIS winner of the specialised path: light index, W (float bits), RNG state, unused
	float4* shade;          // [6][pixels] shading point and LTC coefficients of the specialised path, written by the RIS kernel for the
	                        // winner kernel: {position, roughness} {normal, d0} {diffuse albedo, d1} {F0, d2} {outgoing, d3} {d4, d5, -, -}
	float4* accum;          // [pixels] RGBA32F running mean
	unsigned long long* counters;   // [0] shaded pixels, [1] rays traced, [3] candidates; [4..7] rays, node visits, triangle tests, occluded rays of the counting shadow-ray kernel
	unsigned int* ticket;           // [0] next unclaimed ray of the trace kernel, [1] next tile of ris_ltc3_kernel, [2] of winner_kernel (zero between frames)
	uint32_t pixel_count;
};

namespace RL_NS {
// ---- tiny vector helpers (no operator overloading on purpose: every rounding is visible)
__device__ __forceinline__ float3 mk3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float2 mk2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float dot2(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float3 add3(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 scale3(float3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 mul3(float3 a, float3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return mk2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return mk2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 scale2(float2 a, float s) { return mk2(a.x * s, a.y * s); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// GLSL inversesqrt / normalize as the oracle defines them: 1/sqrt (both correctly rounded), then multiply
__device__ __noinline__ float inversesqrt_ieee(float x) { return 1.0f / sqrtf(x); }
// The same value, RN(1 / RN(sqrt(x))), without the two IEEE sequences and their slow-path branches: one MUFU.RSQ seeds
// both the square root (two residual corrections) and its reciprocal (two Newton steps), all in FMAs. Bit-identical to
// inversesqrt_ieee for EVERY float in [2^-64, 2^64) -- verified exhaustively on the device (risltc_cuda_kat_exact_math,
// tests/test_gpu_kat.py) -- and handed to it outside that range.
__device__ __forceinline__ float inversesqrt(float x) {
	if ((__float_as_uint(x) - 0x1F800000u) >= 0x40000000u) return inversesqrt_ieee(x);
	float y;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	const float h = __fmul_rn(0.5f, y);
	float s = __fmul_rn(x, y);
	s = __fmaf_rn(__fmaf_rn(-s, s, x), h, s);
	s = __fmaf_rn(__fmaf_rn(-s, s, x), h, s);
	float r = __fmaf_rn(__fmaf_rn(-s, y, 1.0f), y, y);
	r = __fmaf_rn(__fmaf_rn(-s, r, 1.0f), r, r);
	// the one mantissa where Newton's last sum is an exact tie and rounds to even, wrongly: s = 2^k (2 - 2^-23), whose
	// reciprocal 2^-(k+1) (1 + 2^-24 + ...) rounds UP to 2^-(k+1) (1 + 2^-23)
	const uint32_t sb = __float_as_uint(s);
	if ((sb & 0x7FFFFFu) == 0x7FFFFFu) r = __uint_as_float(0x7E800001u - (sb & 0x7F800000u));
	return r;
}
// x / 65535 for an integer 0 <= x <= 65535, correctly rounded (checked for all 65536 values by the same device test):
// product with the rounded reciprocal plus one residual correction instead of an IEEE division
__device__ __forceinline__ float unorm16(uint32_t q) {
	const float x = (float) q, c = 1.0f / 65535.0f;
	const float q0 = __fmul_rn(x, c);
	return __fmaf_rn(__fmaf_rn(-q0, 65535.0f, x), c, q0);
}
// x / 255 for an integer 0 <= x <= 255, correctly rounded, by the same sequence (all 256 values checked against exact
// rational arithmetic in tests/test_oracle_identities.py): the texel decode of the material textures
__device__ __forceinline__ float unorm8(uint32_t q) {
	const float x = (float) q, c = 1.0f / 255.0f;
	const float q0 = __fmul_rn(x, c);
	return __fmaf_rn(__fmaf_rn(-q0, 255.0f, x), c, q0);
}
// single MUFU instructions (about 1 ulp, denormals flushed) for the paths where the rounding is free
__device__ __forceinline__ float approx_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float approx_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float approx_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// GLSL's atan / acos / sin / cos have no prescribed rounding; the oracle (oracle/risltc_oracle.c, oracle/glsl_shim.hpp)
// defines them as the CORRECTLY ROUNDED fp32 value, computed as the double-precision function rounded once. A translation
// unit compiled with RL_CR_LIBM (generic.cu, kat.cu, winner_cr.cu) does the same -- its transcendental results are then
// bit-identical to the oracle's (up to the ~2^-29 chance per call that the double result straddles a rounding boundary) --
// and one without it (the production kernels of api.cu) uses CUDA's fp32 functions (1-2 ulp).
#ifdef RL_CR_LIBM
__device__ __forceinline__ float rl_atan(float x) { return (float) atan((double) x); }
__device__ __forceinline__ float rl_acos(float x) { return (float) acos((double) x); }
__device__ __forceinline__ float rl_sin(float x) { return (float) sin((double) x); }
__device__ __forceinline__ float rl_cos(float x) { return (float) cos((double) x); }
#else
__device__ __forceinline__ float rl_atan(float x) { return atanf(x); }
__device__ __forceinline__ float rl_acos(float x) { return acosf(x); }
__device__ __forceinline__ float rl_sin(float x) { return sinf(x); }
__device__ __forceinline__ float rl_cos(float x) { return cosf(x); }
#endif
__device__ __forceinline__ float3 normalize3(float3 a) { return scale3(a, inversesqrt(dot3(a, a))); }
__device__ __forceinline__ float2 normalize2(float2 a) { return scale2(a, inversesqrt(dot2(a, a))); }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float2 rotate_90(float2 v) { return mk2(-v.y, v.x); }
}  // namespace RL_NS
