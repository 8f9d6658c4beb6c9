// shading.cuh -- device functions of the fused RIS + shading kernel.
//
// What is computed is what risltc's shading pass computes (shading_pass.frag.glsl,
// polygon_sampling.glsl, polygon_clipping.glsl, ltc_utility.glsl, brdfs.glsl,
// noise_utility.glsl, reservoir.glsl, mesh_quantization.glsl); how it is organised
// is specific to this kernel: sparse LTC matrices kept as five scalars, the
// world->cosine matrix hoisted out of the 32-candidate loop, horizon clipping as
// "walk + rotation table" instead of a 115-way switch, shadow rays returned to the
// caller as records so that they can be traced by a separate traversal kernel.
#pragma once
#include "common.cuh"

// psa_prepare / psa_sample are called from many places of the generic kernel (out of line there: one copy of the code); the
// phase-structured winner kernel calls each exactly once per loop body and wants them inline, with the polygon in registers
#ifndef RL_PSA_ATTR
#define RL_PSA_ATTR __noinline__
#endif

namespace RL_NS {

// rot[n - 3][above_mask]: slot order of the clipped polygon (tools/gen_clip_table.py); statically initialised, so every
// translation unit that includes this header carries its own copy and nothing is uploaded at run time
__constant__ unsigned char c_clip_rotation[5][128] =
#include "clip_rotation_table.inc"
;

// ------------------------------------------------------------------ noise
// noise_utility.glsl:26-95. The image is defined by the ORDER of these draws.
__device__ __forceinline__ uint32_t murmur3_mix(uint32_t hash, uint32_t k) {
	k *= 0xcc9e2d51u; k = __funnelshift_l(k, k, 15); k *= 0x1b873593u;
	hash ^= k;
	return __funnelshift_l(hash, hash, 13) * 5u + 0xe6546b64u;
}
__device__ __forceinline__ uint32_t murmur3_finalize(uint32_t h) {
	h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
	return h;
}
__device__ __forceinline__ uint32_t noise_seed(uint32_t px, uint32_t py, uint32_t width, uint32_t frame_word) {
	return murmur3_finalize(murmur3_mix(murmur3_mix(0u, px + py * width), frame_word));
}
__device__ __forceinline__ float noise_next(uint32_t& seed) {
	seed = 1664525u * seed + 1013904223u;
	return __uint2float_rn(seed) * 2.3283064365386962890625e-10f;
}

// ------------------------------------------------------- mesh attributes
// mesh_quantization.glsl:38-45
__device__ __forceinline__ float3 decode_position(uint2 q, const float* factor, const float* summand) {
	float px = (float) (q.x & 0x1FFFFFu);
	float py = (float) (((q.x & 0xFFE00000u) >> 21) | ((q.y & 0x3FFu) << 11));
	float pz = (float) ((q.y & 0x7FFFFC00u) >> 10);
	return mk3(fmaf(px, factor[0], summand[0]), fmaf(py, factor[1], summand[1]), fmaf(pz, factor[2], summand[2]));
}
// mesh_quantization.glsl:19-33
__device__ __forceinline__ float3 decode_normal(float ox, float oy) {
	const float factor = 2.0f * (65534.0f / 65535.0f);
	const float summand = -(32768.0f / 65535.0f) * factor;
	ox = fmaf(ox, factor, summand);
	oy = fmaf(oy, factor, summand);
	float3 n = mk3(ox, oy, 1.0f - fabsf(ox) - fabsf(oy));
	if (n.z < 0.0f) {
		float nx = (1.0f - fabsf(n.y)) * ((ox >= 0.0f) ? 1.0f : -1.0f);
		float ny = (1.0f - fabsf(n.x)) * ((oy >= 0.0f) ? 1.0f : -1.0f);
		n.x = nx; n.y = ny;
	}
	return normalize3(n);
}

// ------------------------------------------------------- material textures
// textureGrad with the sampler of scene.c:546-552 under the definition stated in DESIGN.md (texture filtering is driver
// code; the CPU oracle implements the same definition): isotropic level of detail from the larger of the two screen-space footprints, trilinear, repeat
// addressing, taps as c00 + fx (c10 - c00). Same operations in the same order as the oracle: bit-identical results.
__device__ __forceinline__ float4 texture_texel(const SceneView& s, const TextureDesc& t, size_t level_offset, uint32_t w, int x, int y) {
	const size_t index = level_offset + (size_t) y * w + (size_t) x;
	if (t.format == RL_TEXEL_RGBA32F) return __ldg((const float4*) (s.texels + t.offset) + index);
	const uchar4 p = __ldg((const uchar4*) (s.texels + t.offset) + index);
	// unorm8(x) == (float) x / 255.0f for all 256 values (common.cuh)
	if (t.format == RL_TEXEL_RGBA8_SRGB) return make_float4(__ldg(&s.srgb_table[p.x]), __ldg(&s.srgb_table[p.y]), __ldg(&s.srgb_table[p.z]), unorm8(p.w));
	return make_float4(unorm8(p.x), unorm8(p.y), unorm8(p.z), unorm8(p.w));
}
__device__ __forceinline__ int texture_wrap(float coordinate, int size) {
	// coordinate / size: for a power of two the quotient is the product with the (exact) reciprocal, bit for bit -- both are
	// the one correctly rounded value of the same real number -- and mip levels of the reference's assets always are
	const float fs = (float) size;
	const float quotient = ((size & (size - 1)) == 0) ? coordinate * __uint_as_float(0x7F000000u - __float_as_uint(fs)) : coordinate / fs;
	const float wrapped = coordinate - floorf(quotient) * fs;
	const int i = (int) wrapped;
	return (i >= size || i < 0) ? 0 : i;
}
__device__ float4 texture_bilinear(const SceneView& s, const TextureDesc& t, uint32_t level, float u, float v) {
	size_t offset = 0;
	uint32_t w = t.width, h = t.height;
	for (uint32_t l = 0; l != level; ++l) { offset += (size_t) w * h; w = (w > 1u) ? w >> 1 : 1u; h = (h > 1u) ? h >> 1 : 1u; }
	float x = u * (float) w - 0.5f, y = v * (float) h - 0.5f;
	if (!(fabsf(x) < 1.0e9f)) x = 0.0f;
	if (!(fabsf(y) < 1.0e9f)) y = 0.0f;
	const float x0 = floorf(x), y0 = floorf(y), fx = x - x0, fy = y - y0;
	const int ix0 = texture_wrap(x0, (int) w), ix1 = texture_wrap(x0 + 1.0f, (int) w), iy0 = texture_wrap(y0, (int) h), iy1 = texture_wrap(y0 + 1.0f, (int) h);
	const float4 c00 = texture_texel(s, t, offset, w, ix0, iy0), c10 = texture_texel(s, t, offset, w, ix1, iy0);
	const float4 c01 = texture_texel(s, t, offset, w, ix0, iy1), c11 = texture_texel(s, t, offset, w, ix1, iy1);
	float4 r;
	{ const float top = c00.x + fx * (c10.x - c00.x), bottom = c01.x + fx * (c11.x - c01.x); r.x = top + fy * (bottom - top); }
	{ const float top = c00.y + fx * (c10.y - c00.y), bottom = c01.y + fx * (c11.y - c01.y); r.y = top + fy * (bottom - top); }
	{ const float top = c00.z + fx * (c10.z - c00.z), bottom = c01.z + fx * (c11.z - c01.z); r.z = top + fy * (bottom - top); }
	{ const float top = c00.w + fx * (c10.w - c00.w), bottom = c01.w + fx * (c11.w - c01.w); r.w = top + fy * (bottom - top); }
	return r;
}
__device__ float4 texture_sample_grad(const SceneView& s, uint32_t texture_index, float2 uv, float2 ddx, float2 ddy) {
	const TextureDesc t = s.textures[texture_index];
	const float W = (float) t.width, H = (float) t.height;
	const float ax = ddx.x * W, ay = ddx.y * H, bx = ddy.x * W, by = ddy.y * H;
	const float rho_squared = fmaxf(ax * ax + ay * ay, bx * bx + by * by);
	float lambda = 0.0f;
	if (rho_squared > 0.0f && rho_squared < 3.0e38f) lambda = 0.5f * (float) log2((double) rho_squared);   // correctly rounded, like the oracle
	lambda = clampf(lambda, 0.0f, (float) (t.levels - 1u));
	const float level = floorf(lambda), fraction = lambda - level;
	float4 c = texture_bilinear(s, t, (uint32_t) level, uv.x, uv.y);
	if (fraction > 0.0f && (uint32_t) level + 1u < t.levels) {
		const float4 n = texture_bilinear(s, t, (uint32_t) level + 1u, uv.x, uv.y);
		c = make_float4(c.x + fraction * (n.x - c.x), c.y + fraction * (n.y - c.y), c.z + fraction * (n.z - c.z), c.w + fraction * (n.w - c.w));
	}
	return c;
}

struct ShadingPoint {  // shading_data_t, brdfs.glsl:22-39
	float3 position, normal, outgoing;
	float lambert_outgoing;
	float3 diffuse_albedo, fresnel_0;
	float roughness;
};

// get_shading_data, shading_pass.frag.glsl:571-672. Scenes with flat materials (no texture objects) skip the screen-space
// derivative block (:604-627), which only feeds the three textureGrad fetches (:629-633).
// TEXTURED = false compiles the texture path out (the production RIS kernel has one instantiation per case, so that scenes
// with flat materials carry none of the sampler's code or registers).
template <bool TEXTURED = true>
__device__ ShadingPoint reconstruct_shading_point(const SceneView& s, const FrameUniforms& f, uint32_t prim, float3 ray_dir) {
	ShadingPoint r;
	float3 p[3], n[3]; float2 uv[3];
	#pragma unroll
	for (int i = 0; i != 3; ++i) {
		p[i] = decode_position(__ldg(&s.positions[prim * 3u + i]), f.dequant_factor, f.dequant_summand);
		ushort4 a = __ldg(&s.normals_uv[prim * 3u + i]);
		n[i] = decode_normal(unorm16(a.x), unorm16(a.y));
		uv[i] = mk2(fmaf(unorm16(a.z), 8.0f, 0.0f), fmaf(unorm16(a.w), -8.0f, 1.0f));
	}
	float3 origin = mk3(f.camera[0], f.camera[1], f.camera[2]);
	float3 e0 = sub3(p[1], p[0]), e1 = sub3(p[2], p[0]);
	float3 ray_cross_e1 = cross3(ray_dir, e1);
	float rcp_det = 1.0f / dot3(e0, ray_cross_e1);
	float3 to0 = sub3(origin, p[0]);
	float by = rcp_det * dot3(to0, ray_cross_e1);
	float bz = -rcp_det * dot3(ray_dir, cross3(e0, to0));
	float bx = 1.0f - (by + bz);
	r.position = mk3(fmaf(bx, p[0].x, fmaf(by, p[1].x, bz * p[2].x)), fmaf(bx, p[0].y, fmaf(by, p[1].y, bz * p[2].y)), fmaf(bx, p[0].z, fmaf(by, p[1].z, bz * p[2].z)));
	float3 ng = normalize3(mk3(fmaf(bx, n[0].x, fmaf(by, n[1].x, bz * n[2].x)), fmaf(bx, n[0].y, fmaf(by, n[1].y, bz * n[2].y)), fmaf(bx, n[0].z, fmaf(by, n[1].z, bz * n[2].z))));
	uint32_t mat = __ldg(&s.material_indices[prim]);
	float4 m0, m1;   // {base colour rgb, -}, {linear roughness, metalicity, normal.r, normal.g}
	if (TEXTURED && s.textures) {
		const float3 to0_ = sub3(origin, p[0]);
		const float det_0_dir_edge_1 = dot3(to0_, ray_cross_e1);
		const float3 edge_0_cross_0 = cross3(e0, to0_);
		const float det_dir_edge_0_0 = dot3(ray_dir, edge_0_cross_0);
		float3 bd[2];   // screen-space derivatives of the barycentrics, :604-617
		#pragma unroll
		for (int i = 0; i != 2; ++i) {
			const float3 dd = mk3(f.pixel_to_ray[0][i], f.pixel_to_ray[1][i], f.pixel_to_ray[2][i]);
			const float3 rc = cross3(dd, e1);
			const float rcp_det_deriv = -dot3(e0, rc) * rcp_det * rcp_det;
			const float d01 = dot3(to0_, rc);
			bd[i].y = rcp_det_deriv * det_0_dir_edge_1 + rcp_det * d01;
			const float d00 = dot3(dd, edge_0_cross_0);
			bd[i].z = -rcp_det_deriv * det_dir_edge_0_0 - rcp_det * d00;
			bd[i].x = -(bd[i].y + bd[i].z);
		}
		const float2 tc = mk2(fmaf(bx, uv[0].x, fmaf(by, uv[1].x, bz * uv[2].x)), fmaf(bx, uv[0].y, fmaf(by, uv[1].y, bz * uv[2].y)));
		float2 dt[2];   // :622-627, accumulated from zero in the order j = 0, 1, 2
		#pragma unroll
		for (int i = 0; i != 2; ++i) {
			dt[i] = mk2(0.0f, 0.0f);
			dt[i].x += bd[i].x * uv[0].x; dt[i].y += bd[i].x * uv[0].y;
			dt[i].x += bd[i].y * uv[1].x; dt[i].y += bd[i].y * uv[1].y;
			dt[i].x += bd[i].z * uv[2].x; dt[i].y += bd[i].z * uv[2].y;
		}
		const float4 tb = texture_sample_grad(s, 3u * mat + 0u, tc, dt[0], dt[1]);
		const float4 ts = texture_sample_grad(s, 3u * mat + 1u, tc, dt[0], dt[1]);
		const float4 tn = texture_sample_grad(s, 3u * mat + 2u, tc, dt[0], dt[1]);
		m0 = make_float4(tb.x, tb.y, tb.z, ts.x);
		m1 = make_float4(ts.y, ts.z, tn.x, tn.y);
	}
	else { m0 = __ldg(&s.materials[2 * mat]); m1 = __ldg(&s.materials[2 * mat + 1]); }
	float3 base = mk3(m0.x, m0.y, m0.z);
	float3 nts;
	nts.x = fmaf(m1.z, 2.0f, -1.0f);
	nts.y = fmaf(m1.w, 2.0f, -1.0f);
	nts.z = sqrtf(fmaxf(0.0f, fmaf(-nts.x, nts.x, fmaf(-nts.y, nts.y, 1.0f))));
	float metal = m1.y;
	r.diffuse_albedo = mk3(fmaf(base.x, -metal, base.x), fmaf(base.y, -metal, base.y), fmaf(base.z, -metal, base.z));
	float dielectric = 0.02f * (1.0f - metal);
	r.fresnel_0 = mk3(dielectric + base.x * metal, dielectric + base.y * metal, dielectric + base.z * metal);
	r.roughness = clampf((m1.x * m1.x) * f.roughness_factor, 0.0064f, 1.0f);
	float2 t0 = sub2(uv[1], uv[0]), t1 = sub2(uv[2], uv[0]);
	float3 n_x_e0 = cross3(ng, e0), e1_x_n = cross3(e1, ng);
	float3 tangent = add3(scale3(e1_x_n, t0.x), scale3(n_x_e0, t1.x));
	float3 bitangent = add3(scale3(e1_x_n, t0.y), scale3(n_x_e0, t1.y));
	float mean_len = sqrtf(0.5f * (dot3(tangent, tangent) + dot3(bitangent, bitangent)));
	nts.z *= fmaxf(1.0e-10f, mean_len);
	r.normal = normalize3(mk3(tangent.x * nts.x + bitangent.x * nts.y + ng.x * nts.z,
	                          tangent.y * nts.x + bitangent.y * nts.y + ng.y * nts.z,
	                          tangent.z * nts.x + bitangent.z * nts.y + ng.z * nts.z));
	r.outgoing = normalize3(sub3(origin, r.position));
	float offset = fmaxf(0.0f, 1.0e-3f - dot3(r.normal, r.outgoing));
	r.normal = normalize3(mk3(fmaf(offset, r.outgoing.x, r.normal.x), fmaf(offset, r.outgoing.y, r.normal.y), fmaf(offset, r.outgoing.z, r.normal.z)));
	r.lambert_outgoing = dot3(r.normal, r.outgoing);
	return r;
}

// -------------------------------------------------------------------- LTC
// ltc_coefficients_t (ltc_utility.glsl:29-46) with the sparsity of the 3x3 matrices made explicit:
// shading_to_cosine S = [[s00,0,s02],[0,s11,0],[s20,0,s22]] (row, column), likewise its inverse C.
struct LtcFrame {
	float3 rx, ry, rz;   // rows of the world->shading rotation (x axis, y axis, normal)
	float3 t;            // world->shading translation
	float s00, s02, s11, s20, s22;
	float c00, c02, c11, c20, c22;
	float albedo, det;
	// rows of shading_to_cosine * world_to_shading (mat3 * mat4x3), hoisted out of the candidate loop
	float3 qx, qy, qz, qt;
};

// textureLod on the two LTC arrays with the sampler of ltc_table.c:170-177, as exact fp32 bilinear
// (layer = nearest); same definition as the oracle's ltc_fetch.
__device__ void ltc_fetch(const SceneView& s, float u, float v, float layer_coord, float out[6]) {
	int res = (int) s.ltc_res;
	int layer = min(max((int) floorf(layer_coord + 0.5f), 0), (int) s.ltc_layers - 1);
	float x = u * (float) res - 0.5f, y = v * (float) res - 0.5f;
	float fx0 = floorf(x), fy0 = floorf(y);
	float fx = x - fx0, fy = y - fy0;
	int x0 = min(max((int) fx0, 0), res - 1), x1 = min(max((int) fx0 + 1, 0), res - 1);
	int y0 = min(max((int) fy0, 0), res - 1), y1 = min(max((int) fy0 + 1, 0), res - 1);
	float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
	uint32_t base = (uint32_t) layer * res * res;
	uint32_t i00 = base + y0 * res + x0, i10 = base + y0 * res + x1, i01 = base + y1 * res + x0, i11 = base + y1 * res + x1;
	ushort4 a00 = __ldg(&s.ltc_rgba[i00]), a10 = __ldg(&s.ltc_rgba[i10]), a01 = __ldg(&s.ltc_rgba[i01]), a11 = __ldg(&s.ltc_rgba[i11]);
	ushort2 b00 = __ldg(&s.ltc_rg[i00]), b10 = __ldg(&s.ltc_rg[i10]), b01 = __ldg(&s.ltc_rg[i01]), b11 = __ldg(&s.ltc_rg[i11]);
	#define RL_BILERP(c00_, c10_, c01_, c11_) (w00 * unorm16(c00_) + w10 * unorm16(c10_) + w01 * unorm16(c01_) + w11 * unorm16(c11_))
	out[0] = RL_BILERP(a00.x, a10.x, a01.x, a11.x);
	out[1] = RL_BILERP(a00.y, a10.y, a01.y, a11.y);
	out[2] = RL_BILERP(a00.z, a10.z, a01.z, a11.z);
	out[3] = RL_BILERP(a00.w, a10.w, a01.w, a11.w);
	out[4] = RL_BILERP(b00.x, b10.x, b01.x, b11.x);
	out[5] = RL_BILERP(b00.y, b10.y, b01.y, b11.y);
	#undef RL_BILERP
}

// get_ltc_coefficients, ltc_utility.glsl:56-88
__device__ LtcFrame make_ltc_frame(const SceneView& s, float fresnel_0, float roughness, float3 pos, float3 normal, float3 outgoing, const float* c) {
	LtcFrame l;
	float n_dot_o = dot3(normal, outgoing);
	float inclination = rl_acos(clampf(n_dot_o, 0.0f, 1.0f));
	float d[6];
	ltc_fetch(s, fmaf(sqrtf(clampf(roughness, 0.0f, 1.0f)), c[2], c[3]), fmaf(inclination, c[4], c[5]), fmaf(clampf(fresnel_0, 0.0f, 1.0f), c[0], c[1]), d);
	// mat3(d0.x,0,-d0.y, 0,d0.z,0, d0.w,0,d1.x) is column-major: S[0][0]=d0.x, S[2][0]=-d0.y, S[1][1]=d0.z, S[0][2]=d0.w, S[2][2]=d1.x
	l.s00 = d[0]; l.s20 = -d[1]; l.s11 = d[2]; l.s02 = d[3]; l.s22 = d[4];
	l.albedo = d[5];
	float det2 = d[0] * d[4] + d[1] * d[3];
	l.det = d[2] * det2;
	float inv2 = 1.0f / det2;
	// inverse: C[0][0]=d1.x*inv, C[2][0]=d0.y*inv, C[1][1]=1/d0.z, C[0][2]=-d0.w*inv, C[2][2]=d0.x*inv
	l.c00 = d[4] * inv2; l.c20 = d[1] * inv2; l.c11 = 1.0f / d[2]; l.c02 = -d[3] * inv2; l.c22 = d[0] * inv2;
	float3 x_axis = normalize3(mk3(fmaf(-n_dot_o, normal.x, outgoing.x), fmaf(-n_dot_o, normal.y, outgoing.y), fmaf(-n_dot_o, normal.z, outgoing.z)));
	float3 y_axis = cross3(normal, x_axis);
	l.rx = x_axis; l.ry = y_axis; l.rz = normal;
	// -rotation * position, summed column by column like a GLSL mat3 * vec3
	l.t = mk3((-x_axis.x) * pos.x + (-x_axis.y) * pos.y + (-x_axis.z) * pos.z,
	          (-y_axis.x) * pos.x + (-y_axis.y) * pos.y + (-y_axis.z) * pos.z,
	          (-normal.x) * pos.x + (-normal.y) * pos.y + (-normal.z) * pos.z);
	// shading_to_cosine * world_to_shading, element (row i, column j) = S[i][0] W[0][j] + S[i][1] W[1][j] + S[i][2] W[2][j];
	// the structurally-zero products add exact zeros, so only the non-zero ones are written.
	l.qx = mk3(l.s00 * x_axis.x + l.s02 * normal.x, l.s00 * x_axis.y + l.s02 * normal.y, l.s00 * x_axis.z + l.s02 * normal.z);
	l.qy = mk3(l.s11 * y_axis.x, l.s11 * y_axis.y, l.s11 * y_axis.z);
	l.qz = mk3(l.s20 * x_axis.x + l.s22 * normal.x, l.s20 * x_axis.y + l.s22 * normal.y, l.s20 * x_axis.z + l.s22 * normal.z);
	l.qt = mk3(l.s00 * l.t.x + l.s02 * l.t.z, l.s11 * l.t.y, l.s20 * l.t.x + l.s22 * l.t.z);
	return l;
}

// The same frame from the six fetched table values (d) and the shading point, for a kernel that receives them from another
// kernel instead of repeating acos + the bilinear fetch: the identical operations in the identical order.
__device__ LtcFrame ltc_frame_from_fetch(const float d[6], float3 pos, float3 normal, float3 outgoing) {
	LtcFrame l;
	const float n_dot_o = dot3(normal, outgoing);
	l.s00 = d[0]; l.s20 = -d[1]; l.s11 = d[2]; l.s02 = d[3]; l.s22 = d[4];
	l.albedo = d[5];
	const float det2 = d[0] * d[4] + d[1] * d[3];
	l.det = d[2] * det2;
	const float inv2 = 1.0f / det2;
	l.c00 = d[4] * inv2; l.c20 = d[1] * inv2; l.c11 = 1.0f / d[2]; l.c02 = -d[3] * inv2; l.c22 = d[0] * inv2;
	const float3 x_axis = normalize3(mk3(fmaf(-n_dot_o, normal.x, outgoing.x), fmaf(-n_dot_o, normal.y, outgoing.y), fmaf(-n_dot_o, normal.z, outgoing.z)));
	const float3 y_axis = cross3(normal, x_axis);
	l.rx = x_axis; l.ry = y_axis; l.rz = normal;
	l.t = mk3((-x_axis.x) * pos.x + (-x_axis.y) * pos.y + (-x_axis.z) * pos.z,
	          (-y_axis.x) * pos.x + (-y_axis.y) * pos.y + (-y_axis.z) * pos.z,
	          (-normal.x) * pos.x + (-normal.y) * pos.y + (-normal.z) * pos.z);
	l.qx = mk3(l.s00 * x_axis.x + l.s02 * normal.x, l.s00 * x_axis.y + l.s02 * normal.y, l.s00 * x_axis.z + l.s02 * normal.z);
	l.qy = mk3(l.s11 * y_axis.x, l.s11 * y_axis.y, l.s11 * y_axis.z);
	l.qz = mk3(l.s20 * x_axis.x + l.s22 * normal.x, l.s20 * x_axis.y + l.s22 * normal.y, l.s20 * x_axis.z + l.s22 * normal.z);
	l.qt = mk3(l.s00 * l.t.x + l.s02 * l.t.z, l.s11 * l.t.y, l.s20 * l.t.x + l.s22 * l.t.z);
	return l;
}

// world -> shading space (mat4x3 * vec4(v, 1)); `flip` negates the y row (shading_pass.frag.glsl:300-304)
__device__ __forceinline__ float3 to_shading_space(const LtcFrame& l, float3 v, bool flip) {
	float y = l.ry.x * v.x + l.ry.y * v.y + l.ry.z * v.z + l.t.y * 1.0f;
	return mk3(l.rx.x * v.x + l.rx.y * v.y + l.rx.z * v.z + l.t.x * 1.0f, flip ? -y : y,
	           l.rz.x * v.x + l.rz.y * v.y + l.rz.z * v.z + l.t.z * 1.0f);
}
__device__ __forceinline__ float3 to_cosine_space(const LtcFrame& l, float3 v, bool flip) {
	float y = l.qy.x * v.x + l.qy.y * v.y + l.qy.z * v.z + l.qt.y * 1.0f;
	return mk3(l.qx.x * v.x + l.qx.y * v.y + l.qx.z * v.z + l.qt.x * 1.0f, flip ? -y : y,
	           l.qz.x * v.x + l.qz.y * v.y + l.qz.z * v.z + l.qt.z * 1.0f);
}
// evaluate_ltc_density, ltc_utility.glsl:100-105
__device__ __forceinline__ float ltc_density(const LtcFrame& l, float3 dir, float rcp_psa) {
	float3 dc = mk3(l.s00 * dir.x + 0.0f * dir.y + l.s02 * dir.z, 0.0f * dir.x + l.s11 * dir.y + 0.0f * dir.z, l.s20 * dir.x + 0.0f * dir.y + l.s22 * dir.z);
	float l2 = dot3(dc, dc);
	return (fmaxf(0.0f, dc.z) * l.det / (l2 * l2)) * rcp_psa;
}

// ---------------------------------------------------------------- polygons
// iz0, polygon_clipping.glsl:19-25
__device__ __forceinline__ float3 horizon_crossing(float3 a, float3 b) {
	float t = a.z / (a.z - b.z);
	return mk3(fmaf(t, b.x, fmaf(-t, a.x, a.x)), fmaf(t, b.y, fmaf(-t, a.y, a.y)), 0.0f);
}

// clip_polygon, polygon_clipping.glsl:35-225: clip against z >= 0, output order per rotation table.
template <int P>
__device__ uint32_t clip_to_horizon(uint32_t n, float3 (&v)[P], uint32_t min_vertices) {
	uint32_t mask = 0;
	#pragma unroll
	for (int i = 0; i != P - 1; ++i)
		if (v[i].z > 0.0f && ((uint32_t) i < min_vertices || (uint32_t) i < n)) mask |= 1u << i;
	if (n < 3u || n > 7u) return 0;
	uint32_t rot = c_clip_rotation[n - 3][mask];
	if (rot == 0xFFu) return 0;
	float3 walk[P];
	uint32_t vc = 0;
	#pragma unroll
	for (int i = 0; i != P - 1; ++i) {
		if ((uint32_t) i >= n) break;
		uint32_t j = ((uint32_t) i + 1u == n) ? 0u : (uint32_t) i + 1u;
		uint32_t a = (mask >> i) & 1u, b = (mask >> j) & 1u;
		if (a) walk[vc++] = v[i];
		if (a != b) walk[vc++] = horizon_crossing(v[i], v[j]);
	}
	for (uint32_t j = 0; j != vc; ++j) {
		uint32_t k = j + rot; if (k >= vc) k -= vc;
		v[j] = walk[k];
	}
	return vc;
}

// integrateEdgeVec + calculate_ltc, polygon_sampling.glsl:508-530 (closed loop over the vc clipped vertices)
__device__ __forceinline__ float edge_form_factor(float3 a, float3 b) {
	a = normalize3(a); b = normalize3(b);
	float x = dot3(a, b), y = fabsf(x);
	float num = 0.8543985f + (0.4965155f + 0.0145206f * y) * y;
	float den = 3.4175940f + (4.1616724f + y) * y;
	float v = num / den;
	float theta_sintheta = (x > 0.0f) ? v : 0.5f * (1.0f / sqrtf(fmaxf(1.0f - x * x, 1e-7f))) - v;
	return cross3(a, b).z * theta_sintheta;
}
template <int P>
__device__ float polygon_form_factor(uint32_t vc, const float3 (&v)[P]) {
	float sum = 0.0f;
	for (uint32_t i = 0; i != vc; ++i) sum += edge_form_factor(v[i], v[(i + 1u == vc) ? 0u : i + 1u]);
	return fabsf(sum);
}

// ---- projected solid angle sampling (polygon_sampling.glsl:231-828)
__device__ __forceinline__ float kahan(float a, float b, float c, float d) {   // :262-270
	float cd = __fmul_rn(c, d);
	float err = __fmaf_rn(c, d, -cd);
	float res = __fmaf_rn(a, b, -cd);
	return __fsub_rn(res, err);
}
__device__ __forceinline__ bool sign_bit(float x) { return (__float_as_uint(x) & 0x80000000u) != 0u; }   // is_inner_ellipse :293-300
__device__ __forceinline__ float2 ellipse_from_edge(float3 a, float3 b) {   // :320-329
	float nx = kahan(a.y, b.z, a.z, b.y), ny = kahan(a.z, b.x, a.x, b.z), nz = kahan(a.x, b.y, a.y, b.x);
	float scaling = 1.0f / nz;
	scaling = sign_bit(nx) ? -scaling : scaling;
	float2 e = mk2(nx * scaling, ny * scaling);
	e.x = (nz != 0.0f) ? e.x : INFINITY;
	return e;
}
__device__ __forceinline__ float2 ellipse_transform(float2 e, float2 p) { float d = dot2(e, p); return mk2(fmaf(d, e.x, p.x), fmaf(d, e.y, p.y)); }
__device__ __forceinline__ float ellipse_det(float2 e) { return fmaf(e.x, e.x, fmaf(e.y, e.y, 1.0f)); }
__device__ __forceinline__ float ellipse_rsqrt_det(float2 e) { return inversesqrt(ellipse_det(e)); }
__device__ __forceinline__ float ellipse_dir_factor_rsq(float2 e, float2 d) { float ed = dot2(e, d); return fmaf(ed, ed, dot2(d, d)); }

__device__ __forceinline__ float fast_positive_atan(float y) {   // :84-98
	float rx = (fabsf(y) > 1.0f) ? (1.0f / fabsf(y)) : fabsf(y);
	float ry = rx * rx;
	float rz = fmaf(ry, 0.02083509974181652f, -0.08513300120830536f);
	rz = fmaf(ry, rz, 0.18014100193977356f);
	rz = fmaf(ry, rz, -0.3302994966506958f);
	ry = fmaf(ry, rz, 0.9998660087585449f);
	rz = fmaf(-2.0f * ry, rx, RL_HALF_PI);
	rz = (fabsf(y) > 1.0f) ? rz : 0.0f;
	rx = fmaf(rx, ry, rz);
	return (y < 0.0f) ? (RL_PI - rx) : rx;
}
// USE_FAST_ATAN and the biased sampler are compile-time switches here like in the reference (main.c:962-991). They used to be
// run-time bools handed down to the out-of-line PSA functions; ptxas 12.9 keeps such a warp-uniform argument in a uniform
// register, reuses that register for the double-precision atan's constants in one divergent branch (decentral polygons) and
// reads it afterwards in the other (central polygons of the same warp): whole warps silently took the fast atan. Found by
// the known-answer test of the PSA preparation once its transcendental functions were correctly rounded.
template <bool FAST>
__device__ __forceinline__ float positive_atan(float tangent) {   // :105-112
	if (FAST) return fast_positive_atan(tangent);
	return rl_atan(tangent) + ((tangent < 0.0f) ? RL_PI : 0.0f);
}
template <bool FAST>
__device__ __forceinline__ float area_from_tangents(float inner_rsqrt, float inner_tan, float outer_rsqrt, float outer_tan) {   // :381-386
	float inner_area = inner_rsqrt * positive_atan<FAST>(inner_tan);
	float r = fmaf(outer_rsqrt, positive_atan<FAST>(outer_tan), -inner_area);
	return (r > 0.0f) ? (0.5f * r) : 0.0f;
}
__device__ __forceinline__ float mix_fma(float x, float y, float a) { return fmaf(a, y, fmaf(-a, x, x)); }   // :184-186

template <int P>
struct PsaPolygon {   // projected_solid_angle_polygon_t, :231-253
	uint32_t vc;
	float2 v[P], e[P];
	float2 inner0;
	float sector[P];
	float total;
};

template <int P>
__device__ __forceinline__ void psa_compare_swap(PsaPolygon<P>& p, int l, int r) {   // :425-439
	float2 a = p.v[l], b = p.v[r];
	float nz = kahan(a.x, -b.y, a.y, -b.x);
	bool swap = (nz == 0.0f) ? (isinf(p.e[r].x) != 0) : (nz > 0.0f);
	if (swap) { p.v[l] = b; p.v[r] = a; float2 t = p.e[l]; p.e[l] = p.e[r]; p.e[r] = t; }
}

template <int P>
__device__ void psa_sort(PsaPolygon<P>& p) {   // :444-506
	uint32_t n = p.vc;
	if (n == 3) psa_compare_swap(p, 1, 2);
	if (P >= 4 && n == 4) psa_compare_swap(p, 1, 3);
	if (P >= 5 && n == 5) { psa_compare_swap(p, 2, 4); psa_compare_swap(p, 1, 3); psa_compare_swap(p, 1, 2); psa_compare_swap(p, 0, 3); psa_compare_swap(p, 3, 4); }
	if (P >= 6 && n == 6) { psa_compare_swap(p, 3, 5); psa_compare_swap(p, 2, 4); psa_compare_swap(p, 1, 5); psa_compare_swap(p, 0, 4); psa_compare_swap(p, 4, 5); psa_compare_swap(p, 1, 3); }
	if (P >= 7 && n == 7) { psa_compare_swap(p, 2, 5); psa_compare_swap(p, 1, 6); psa_compare_swap(p, 5, 6); psa_compare_swap(p, 3, 4); psa_compare_swap(p, 0, 4); psa_compare_swap(p, 4, 6); psa_compare_swap(p, 1, 3); psa_compare_swap(p, 3, 5); psa_compare_swap(p, 4, 5); }
	if (P >= 8 && n == 8) { psa_compare_swap(p, 2, 6); psa_compare_swap(p, 3, 7); psa_compare_swap(p, 1, 5); psa_compare_swap(p, 0, 4); psa_compare_swap(p, 4, 6); psa_compare_swap(p, 5, 7); psa_compare_swap(p, 6, 7); psa_compare_swap(p, 4, 5); psa_compare_swap(p, 1, 3); }
	psa_compare_swap(p, 0, 2);
	if (P >= 4 && n >= 4) psa_compare_swap(p, 2, 3);
	psa_compare_swap(p, 0, 1);
}

// prepare_projected_solid_angle_polygon_sampling, :545-613
template <int P, bool FAST>
__device__ RL_PSA_ATTR void psa_prepare(PsaPolygon<P>& p, uint32_t vc, const float3 (&v)[P]) {
	p.vc = vc;
	float2 inner0 = mk2(1.0f, 0.0f);
	p.v[0] = mk2(v[0].x, v[0].y);
	p.e[0] = ellipse_from_edge(v[0], v[1]);
	float2 prev = p.e[0];
	#pragma unroll
	for (int i = 1; i != P; ++i) {
		if ((uint32_t) i >= vc) break;
		p.v[i] = mk2(v[i].x, v[i].y);
		// the successor of the last vertex is vertex 0: selected by VALUE between two statically indexed slots (a selected
		// index would force the array into local memory)
		const float3 successor = ((uint32_t) i + 1u == vc || i + 1 == P) ? v[0] : v[(i + 1 < P) ? i + 1 : 0];
		float2 e = ellipse_from_edge(v[i], successor);
		bool inner = sign_bit(e.x);
		p.e[i] = inner ? prev : e;
		inner0 = (sign_bit(prev.x) && !inner) ? prev : inner0;
		prev = e;
	}
	{
		float2 e = p.e[0];
		bool inner = sign_bit(e.x);
		p.e[0] = inner ? prev : e;
		inner0 = (sign_bit(prev.x) && !inner) ? prev : inner0;
	}
	p.inner0 = inner0;
	p.total = 0.0f;
	if (inner0.x > 0.0f) {   // central case: one ellipse per sector
		#pragma unroll
		for (int i = 0; i != P; ++i) {
			if ((uint32_t) i >= vc) break;
			float2 d0 = p.v[i], d1 = ((uint32_t) i + 1u == vc || i + 1 == P) ? p.v[0] : p.v[(i + 1 < P) ? i + 1 : 0];
			float rs = ellipse_rsqrt_det(p.e[i]);
			float det_dirs = fmaxf(+0.0f, dot2(d1, rotate_90(d0)));
			float edot = rs * dot2(d0, ellipse_transform(p.e[i], d1));
			float area = 0.5f * rs * positive_atan<FAST>(det_dirs / edot);
			p.sector[i] = (rs > 0.0f) ? area : 0.0f;
			p.total += p.sector[i];
		}
	}
	else {
		psa_sort(p);
		float2 inner = inner0, outer = mk2(0.0f, 0.0f);
		float inner_rs = ellipse_rsqrt_det(inner), outer_rs = 0.0f;
		#pragma unroll
		for (int i = 0; i != P - 1; ++i) {
			if ((uint32_t) i + 1u >= vc) break;
			float2 ve = p.e[i];
			bool vin = sign_bit(ve.x);
			float vrs = ellipse_rsqrt_det(ve);
			if (i == 0) { outer = ve; outer_rs = vrs; }
			else {
				inner = vin ? ve : inner; inner_rs = vin ? vrs : inner_rs;
				outer = vin ? outer : ve; outer_rs = vin ? outer_rs : vrs;
			}
			float2 d0 = p.v[i], d1 = p.v[i + 1];
			float det_dirs = fmaxf(+0.0f, dot2(d1, rotate_90(d0)));
			float idot = inner_rs * dot2(d0, ellipse_transform(inner, d1));
			float odot = outer_rs * dot2(d0, ellipse_transform(outer, d1));
			p.sector[i] = area_from_tangents<FAST>(inner_rs, det_dirs / idot, outer_rs, det_dirs / odot);
			p.total += p.sector[i];
		}
	}
}

__device__ __forceinline__ float2 solve_homogeneous_quadratic(float q00, float q01, float q10, float q11) {   // :649-654, q[col][row]
	float cxy = 0.5f * (q01 + q10);
	float sd = sqrtf(fmaxf(0.0f, cxy * cxy - q00 * q11));
	float root = fabsf(cxy) + sd;
	return (cxy >= 0.0f) ? mk2(root, -q00) : mk2(q11, root);
}

// sample_sector_between_ellipses, :668-762
template <bool FAST, bool BIASED>
__device__ float2 sample_between_ellipses(float2 rn, float target_area, float2 inner, float2 outer, float2 dir_0, float2 dir_1) {
	float2 q0 = normalize2(dir_0), q2 = normalize2(dir_1);
	float2 q1 = add2(q0, q2);
	float i0 = inversesqrt(fmaf(dot2(inner, q0), dot2(inner, q0), 1.0f)), i1 = inversesqrt(ellipse_dir_factor_rsq(inner, q1)), i2 = inversesqrt(fmaf(dot2(inner, q2), dot2(inner, q2), 1.0f));
	float o0 = inversesqrt(fmaf(dot2(outer, q0), dot2(outer, q0), 1.0f)), o1 = inversesqrt(ellipse_dir_factor_rsq(outer, q1)), o2 = inversesqrt(fmaf(dot2(outer, q2), dot2(outer, q2), 1.0f));
	float area0 = o0 * o1 - i0 * i1, area1 = o1 * o2 - i1 * i2;
	float tq = mix_fma(-area0, area1, rn.x);
	bool first = (tq <= 0.0f);
	q2 = first ? q0 : q2; i2 = first ? i0 : i2; o2 = first ? o0 : o2;
	tq += first ? area0 : -area1;
	tq *= fabsf(q1.x * q2.y - q2.x * q1.y);
	float2 n0 = ellipse_transform(inner, add2(scale2(q1, i1), scale2(q2, i2)));
	float2 n1 = ellipse_transform(outer, add2(scale2(q1, o1), scale2(q2, o2)));
	float off0 = dot2(n0, q1) * i1, off1 = dot2(n1, q1) * o1;
	float2 r2 = rotate_90(q2);
	// quadratic = outerProduct(a, n0) - outerProduct(b, n1), [col j][row i] = a[i] n0[j] - b[i] n1[j]
	float2 a = scale2(r2, off1 * o2);
	float2 b = add2(scale2(r2, off0 * i2), scale2(n0, tq));
	float2 cur = solve_homogeneous_quadratic(a.x * n0.x - b.x * n1.x, a.y * n0.x - b.y * n1.x, a.x * n0.y - b.x * n1.y, a.y * n0.y - b.y * n1.y);
	if (!BIASED) {
		int iterations = (fabsf(rn.x - 0.5f) <= 0.5f - 1.0e-5f) ? 2 : 0;
		float inner_rs = ellipse_rsqrt_det(inner), outer_rs = ellipse_rsqrt_det(outer);
		for (int it = 0; it != iterations; ++it) {
			// normalize_approx_and_flip, :622-634
			float sc = __uint_as_float(__float_as_uint(fabsf(cur.x) + fabsf(cur.y)) ^ 0x7F800000u);
			sc = (dot2(cur, q1) >= 0.0f) ? sc : -sc;
			cur = scale2(cur, sc);
			float2 id = ellipse_transform(inner, cur), od = ellipse_transform(outer, cur);
			float det_dirs = fmaxf(+0.0f, dot2(cur, rotate_90(q0)));
			float err = target_area - area_from_tangents<FAST>(inner_rs, det_dirs / (inner_rs * dot2(q0, id)), outer_rs, det_dirs / (outer_rs * dot2(q0, od)));
			float2 c = sub2(id, od), rc = rotate_90(cur), e2 = scale2(id, 2.0f * err);
			cur = solve_homogeneous_quadratic(c.x * rc.x - e2.x * od.x, c.y * rc.x - e2.y * od.x, c.x * rc.y - e2.x * od.y, c.y * rc.y - e2.y * od.y);
		}
	}
	cur = (dot2(cur, q1) >= 0.0f) ? cur : mk2(-cur.x, -cur.y);
	float fi = 1.0f / ellipse_dir_factor_rsq(inner, cur), fo = 1.0f / ellipse_dir_factor_rsq(outer, cur);
	return scale2(cur, sqrtf(mix_fma(fi, fo, rn.y)));
}

// sample_projected_solid_angle_polygon, :772-828
template <int P, bool FAST, bool BIASED>
__device__ RL_PSA_ATTR float3 psa_sample(const PsaPolygon<P>& p, float u0, float u1) {
	float target = u0 * p.total;
	float2 s, outer = mk2(0.0f, 0.0f), d0 = mk2(0.0f, 0.0f);
	if (p.inner0.x > 0.0f) {
		#pragma unroll
		for (int i = 0; i != P; ++i) {
			if (i > 0) target -= p.sector[i - 1];
			outer = p.e[i]; d0 = p.v[i];
			if ((i >= 2 && (uint32_t) i + 1u == p.vc) || target < p.sector[i]) break;
		}
		float sqrt_det = sqrtf(ellipse_det(outer));
		float angle = 2.0f * target * sqrt_det;
		float2 t = rotate_90(ellipse_transform(outer, d0));
		float ca = rl_cos(angle) * sqrt_det, sa = rl_sin(angle);
		s = mk2(ca * d0.x + sa * t.x, ca * d0.y + sa * t.y);
		s = scale2(s, sqrtf(u1 / ellipse_dir_factor_rsq(outer, s)));
	}
	else {
		float sector_psa = 0.0f;
		float2 inner = p.inner0, d1 = mk2(0.0f, 0.0f);
		#pragma unroll
		for (int i = 0; i != P - 1; ++i) {
			float2 ve = p.e[i];
			if (i == 0) outer = ve;
			else {
				target -= p.sector[i - 1];
				bool vin = sign_bit(ve.x);
				inner = vin ? ve : inner; outer = vin ? outer : ve;
			}
			d0 = p.v[i]; d1 = p.v[i + 1]; sector_psa = p.sector[i];
			if ((i >= 1 && (uint32_t) i + 2u == p.vc) || target < sector_psa) break;
		}
		s = sample_between_ellipses<FAST, BIASED>(mk2(target / sector_psa, u1), target, inner, outer, d0, d1);
	}
	return mk3(s.x, s.y, sqrtf(fmaxf(0.0f, fmaf(-s.x, s.x, fmaf(-s.y, s.y, 1.0f)))));
}

// run-time selection between the compiled variants, at the call site (the flags come straight from kernel parameters there)
template <int P>
__device__ __forceinline__ void psa_prepare_rt(PsaPolygon<P>& p, uint32_t vc, const float3 (&v)[P], bool fast) {
	if (fast) psa_prepare<P, true>(p, vc, v); else psa_prepare<P, false>(p, vc, v);
}
template <int P>
__device__ __forceinline__ float3 psa_sample_rt(const PsaPolygon<P>& p, float u0, float u1, bool fast, bool biased) {
	if (fast) return biased ? psa_sample<P, true, true>(p, u0, u1) : psa_sample<P, true, false>(p, u0, u1);
	return biased ? psa_sample<P, false, true>(p, u0, u1) : psa_sample<P, false, false>(p, u0, u1);
}

// ------------------------------------------------------------------- BRDF
// evaluate_brdf (diffuse + specular), brdfs.glsl:58-93: Lambert + GGX / Smith with greyscale Schlick Fresnel
__device__ float3 evaluate_brdf(const ShadingPoint& sp, float3 incoming) {
	float3 h = normalize3(add3(incoming, sp.outgoing));
	float lambert_in = dot3(sp.normal, incoming);
	float o_dot_h = dot3(sp.outgoing, h);
	float n_dot_h = dot3(sp.normal, h);
	float r2 = sp.roughness * sp.roughness;
	float ggx = fmaf(fmaf(n_dot_h, r2, -n_dot_h), n_dot_h, 1.0f);
	ggx = r2 / (ggx * ggx);
	float masking = lambert_in * sqrtf(fmaf(fmaf(-sp.lambert_outgoing, r2, sp.lambert_outgoing), sp.lambert_outgoing, r2));
	float shadowing = sp.lambert_outgoing * sqrtf(fmaf(fmaf(-lambert_in, r2, lambert_in), lambert_in, r2));
	float smith = 0.5f / (masking + shadowing);
	float flipped = 1.0f - clampf(o_dot_h, 0.0f, 1.0f);
	float f2 = flipped * flipped;
	float w = f2 * flipped * f2;
	float3 fr = mk3(sp.fresnel_0.x + (1.0f - sp.fresnel_0.x) * w, sp.fresnel_0.y + (1.0f - sp.fresnel_0.y) * w, sp.fresnel_0.z + (1.0f - sp.fresnel_0.z) * w);
	float spec = ggx * smith * dot3(fr, mk3(0.21263901f, 0.71516868f, 0.07219232f));
	return mk3((sp.diffuse_albedo.x + spec) * RL_INV_PI, (sp.diffuse_albedo.y + spec) * RL_INV_PI, (sp.diffuse_albedo.z + spec) * RL_INV_PI);
}

// get_mis_estimate, shading_pass.frag.glsl:219-269
__device__ float3 mis_estimate(uint32_t heuristic, float3 integrand, float3 sw, float sd, float3 ow, float od, float ve) {
	if (heuristic == MIS_WEIGHTED) {
		float3 ws = add3(scale3(sw, sd), scale3(ow, od));
		float3 num = mul3(sw, integrand);
		return mk3(num.x / ws.x, num.y / ws.y, num.z / ws.z);
	}
	if (heuristic == MIS_OPTIMAL_CLAMPED || heuristic == MIS_OPTIMAL) {
		float balance = 1.0f / (sd + od);
		float3 ws = add3(scale3(sw, sd), scale3(ow, od));
		if (heuristic == MIS_OPTIMAL_CLAMPED) {
			float m = fmaf(-ve, balance, balance);
			return mk3(fmaf(ve, sw.x / ws.x, m) * integrand.x, fmaf(ve, sw.y / ws.y, m) * integrand.y, fmaf(ve, sw.z / ws.z, m) * integrand.z);
		}
		return mk3(ve * sw.x + balance * (integrand.x - ve * ws.x), ve * sw.y + balance * (integrand.y - ve * ws.y), ve * sw.z + balance * (integrand.z - ve * ws.z));
	}
	float w = (heuristic == MIS_BALANCE) ? 1.0f / (sd + od) : sd / (sd * sd + od * od);
	return scale3(integrand, w);
}

// ------------------------------------------------------------------ lights
template <int V>
struct Light {   // polygonal_light_t, polygonal_light_utility.glsl:27-43
	float3 radiance;
	float4 plane;
	uint32_t count;
	float3 v[V];
};
template <int V>
__device__ __forceinline__ Light<V> load_light(const SceneView& s, uint32_t index) {
	const float4* rec = s.lights + (size_t) index * s.light_stride4;
	Light<V> l;
	float4 a = __ldg(rec), b = __ldg(rec + 1), c = __ldg(rec + 2);
	l.radiance = mk3(a.x, a.y, a.z); l.plane = b; l.count = __float_as_uint(c.x);
	#pragma unroll
	for (int i = 0; i != V; ++i) { float4 p = __ldg(rec + 3 + i); l.v[i] = mk3(p.x, p.y, p.z); }
	return l;
}
__device__ __forceinline__ float plane_side(float3 pos, float4 plane) { return pos.x * plane.x + pos.y * plane.y + pos.z * plane.z + 1.0f * plane.w; }

// One shadow ray that the estimator asks for, with what it contributes (see kernels.cu)
struct RayRequest {
	float3 dir;        // world space, normalised
	float t_max;
	float3 if_visible; // term added to the light sample's sum when the ray is unoccluded
};

// Candidate target function in LTC_CP mode: the LTC integral of the unshadowed diffuse + specular
// lobes (evaluate_polygonal_light_shading, shading_pass.frag.glsl:430-456). No random numbers, no rays.
template <int V>
__device__ float3 ltc_target(const ShadingPoint& sp, const LtcFrame& ltc, const Light<V>& light, uint32_t min_vertices) {
	bool flip = plane_side(sp.position, light.plane) < 0.0f;
	float3 pv[V + 1];
	#pragma unroll
	for (int i = 0; i != V; ++i) pv[i] = to_shading_space(ltc, light.v[i], flip);
	pv[V] = mk3(0.0f, 0.0f, 0.0f);
	float3 result = mk3(0.0f, 0.0f, 0.0f);
	uint32_t vc = clip_to_horizon<V + 1>(light.count, pv, min_vertices);
	if (vc > 0) {
		float f = polygon_form_factor<V + 1>(vc, pv);
		result = mul3(scale3(sp.diffuse_albedo, f), light.radiance);
	}
	#pragma unroll
	for (int i = 0; i != V; ++i) pv[i] = to_cosine_space(ltc, light.v[i], flip);
	vc = clip_to_horizon<V + 1>(light.count, pv, min_vertices);
	if (vc > 0) {
		float f = polygon_form_factor<V + 1>(vc, pv) * ltc.albedo;
		result = add3(result, scale3(light.radiance, f));
	}
	return result;
}

// The "prepare both techniques" block shared by shading_pass.frag.glsl:296-363 and :462-516
// what the per-sample body needs of the two prepared techniques (the polygons themselves are only needed to draw the samples)
struct TechniqueTerms {
	float3 diffuse_weight, specular_weight;
	float rcp_diffuse, rcp_specular;
	float specular_total;   // projected solid angle of the specular polygon (0: single-technique sample)
	bool flip;
};
template <int V>
struct Techniques : TechniqueTerms {
	PsaPolygon<V + 1> diffuse, specular;
	bool valid;
};

// diffuse / specular weights and reciprocal solid angles from the two totals (shading_pass.frag.glsl:318-336, :347-352)
__device__ __forceinline__ void technique_weights(TechniqueTerms& t, const ShadingPoint& sp, float ltc_albedo, float diffuse_total, float specular_total, float3 radiance, bool optimal) {
	float sw = ltc_albedo * specular_total;
	float3 da = mk3(fmaxf(sp.diffuse_albedo.x, 0.01f), fmaxf(sp.diffuse_albedo.y, 0.01f), fmaxf(sp.diffuse_albedo.z, 0.01f));
	t.diffuse_weight = scale3(da, diffuse_total);
	t.rcp_diffuse = 1.0f / diffuse_total;
	t.rcp_specular = 1.0f / specular_total;
	t.specular_weight = mk3(sw, sw, sw);
	t.specular_total = specular_total;
	if (optimal) {
		float3 rop = scale3(radiance, RL_INV_PI);
		t.diffuse_weight = mul3(t.diffuse_weight, rop);
		t.specular_weight = mul3(t.specular_weight, rop);
	}
}

template <int V>
__device__ void prepare_techniques(Techniques<V>& t, const ShadingPoint& sp, const LtcFrame& ltc, const Light<V>& light, const Variant& var) {
	t.valid = false;
	t.flip = plane_side(sp.position, light.plane) < 0.0f;
	t.specular.total = 0.0f;
	t.specular_total = 0.0f;
	bool fast = var.fast_atan != 0;
	float3 pv[V + 1];
	#pragma unroll
	for (int i = 0; i != V; ++i) pv[i] = to_shading_space(ltc, light.v[i], t.flip);
	pv[V] = mk3(0.0f, 0.0f, 0.0f);
	uint32_t vc = clip_to_horizon<V + 1>(light.count, pv, var.min_light_vertices);
	if (vc == 0) return;
	psa_prepare_rt<V + 1>(t.diffuse, vc, pv, fast);
	#pragma unroll
	for (int i = 0; i != V; ++i) pv[i] = to_cosine_space(ltc, light.v[i], t.flip);
	vc = clip_to_horizon<V + 1>(light.count, pv, var.min_light_vertices);
	if (vc != 0) psa_prepare_rt<V + 1>(t.specular, vc, pv, fast);
	if (t.diffuse.total == 0.0f) return;
	technique_weights(t, sp, ltc.albedo, t.diffuse.total, t.specular.total, light.radiance, var.mis_heuristic == MIS_OPTIMAL);
	t.valid = true;
}

// One pass of the per-sample body (shading_pass.frag.glsl:365-394 / :518-550) for technique j, up to the
// point where visibility is needed. Returns false when the sample is skipped (dir.z <= 0).
// side_visible tells whether a shadow ray has to be traced; if_visible / if_occluded are the terms to add.
template <int V>
__device__ bool technique_sample(const TechniqueTerms& t, const ShadingPoint& sp, const LtcFrame& ltc, const Light<V>& light, const Variant& var,
	int j, float3 dir, float mis_ve, bool peters, RayRequest& ray, bool& side_visible, float3& if_occluded)
{
	if (dir.z <= 0.0f) return false;
	float dd = dir.z * t.rcp_diffuse;
	float sd = ltc_density(ltc, dir, t.rcp_specular);
	// transpose(world_to_shading) * dir
	float dy = t.flip ? -dir.y : dir.y;   // the flipped frame has its y row negated
	float3 w = mk3(ltc.rx.x * dir.x + (t.flip ? -ltc.ry.x : ltc.ry.x) * dir.y + ltc.rz.x * dir.z,
	               ltc.rx.y * dir.x + (t.flip ? -ltc.ry.y : ltc.ry.y) * dir.y + ltc.rz.y * dir.z,
	               ltc.rx.z * dir.x + (t.flip ? -ltc.ry.z : ltc.ry.z) * dir.y + ltc.rz.z * dir.z);
	(void) dy;
	side_visible = dot3(sp.normal, w) > 0.0f;
	float3 full = mk3(0.0f, 0.0f, 0.0f);
	if (side_visible) {
		float3 rb = mul3(light.radiance, evaluate_brdf(sp, w));
		full = mk3(dir.z * rb.x, dir.z * rb.y, dir.z * rb.z);
		ray.dir = w;
		ray.t_max = -plane_side(sp.position, light.plane) / (w.x * light.plane.x + w.y * light.plane.y + w.z * light.plane.z) - 1e-3f;
	}
	float3 zero = mk3(0.0f, 0.0f, 0.0f);
	bool single = (t.specular_total <= 0.0f);
	if (j == 0 && single) {
		float r = 1.0f / dd;
		ray.if_visible = scale3(full, r);
		if_occluded = peters ? zero : scale3(zero, r);
	}
	else if (j == 0) {
		ray.if_visible = mis_estimate(var.mis_heuristic, full, t.diffuse_weight, dd, t.specular_weight, sd, mis_ve);
		if_occluded = mis_estimate(var.mis_heuristic, zero, t.diffuse_weight, dd, t.specular_weight, sd, mis_ve);
	}
	else {
		ray.if_visible = mis_estimate(var.mis_heuristic, full, t.specular_weight, sd, t.diffuse_weight, dd, mis_ve);
		if_occluded = mis_estimate(var.mis_heuristic, zero, t.specular_weight, sd, t.diffuse_weight, dd, mis_ve);
	}
	return true;
}

// cosine -> shading space direction of a specular sample (shading_pass.frag.glsl:371)
__device__ __forceinline__ float3 cosine_to_shading_dir(const LtcFrame& l, float3 d) {
	return normalize3(mk3(l.c00 * d.x + 0.0f * d.y + l.c02 * d.z, 0.0f * d.x + l.c11 * d.y + 0.0f * d.z, l.c20 * d.x + 0.0f * d.y + l.c22 * d.z));
}

// Turk area sampling baseline, polygon_sampling_related_work.glsl:34-64 and shading_pass.frag.glsl:148-152
template <int V>
__device__ __forceinline__ float3 turk_sample(const Light<V>& l, float u0, float u1) {
	float s = sqrtf(u0);
	float b0 = 1.0f - s, b1 = s * u1, b2 = fmaf(-s, u1, s);
	return add3(add3(scale3(l.v[0], b0), scale3(l.v[1], b1)), scale3(l.v[2], b2));
}
template <int V>
__device__ __forceinline__ float light_area_012(const Light<V>& l) {
	float3 a = sub3(l.v[1], l.v[0]), b = sub3(l.v[2], l.v[0]);
	float3 c = mk3(kahan(a.y, b.z, a.z, b.y), kahan(a.z, b.x, a.x, b.z), kahan(a.x, b.y, a.y, b.x));
	return sqrtf(dot3(c, c)) / 2.0f;
}

}  // namespace RL_NS
