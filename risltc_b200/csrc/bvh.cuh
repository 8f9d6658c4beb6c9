// bvh.cuh -- software ray traversal (B200 has no RT cores).
//
// Replaces the driver's acceleration structure behind rayQueryEXT
// (shading_pass.frag.glsl:120-126, built at scene.c:142-406) and the fixed-function
// rasteriser of the visibility pass (visibility_pass.*, main.c:715-721,751-756).
//
// The triangle test is the definition shared with the CPU oracle
// (oracle/risltc_oracle_frame.inc): Moeller-Trumbore with true divisions and NO
// fused multiply-adds, written with __f*_rn intrinsics. Hit / no-hit and primitive
// ids therefore agree bit for bit with the oracle, independently of the shape of
// the tree. The three IEEE divisions are only executed when the decision is close:
// the same numerators times one MUFU reciprocal, compared with safety margins far
// above its error, settle every other case (far misses and clear hits) identically.
//
// Node layout: 64 bytes = both children's boxes + two child references, fetched as
// four 16-byte loads; triangles are 48 bytes {v0 | id, e1, e2} = three 16-byte loads.
// Boxes are padded by 1e-5 of the scene extent at build time (bvh_build.cpp), which
// covers the rounding of the fma slab test.
#pragma once
#include "common.cuh"

namespace RL_NS {

#define RL_STACK 64

__device__ __forceinline__ float xdot3(float3 a, float3 b) {
	return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ float3 xcross3(float3 a, float3 b) {
	return mk3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)),
	           __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
	           __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float3 xsub3(float3 a, float3 b) { return mk3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }

// The exactly rounded numerators of the oracle's tri_test: u = su / det, v = sv / det, t = st / det.
struct TriTerms { float det, su, sv, st; };
__device__ __forceinline__ TriTerms tri_terms(const BvhTri& tr, float3 o, float3 d) {
	float3 v0 = mk3(tr.v0.x, tr.v0.y, tr.v0.z), e1 = mk3(tr.e1.x, tr.e1.y, tr.e1.z), e2 = mk3(tr.e2.x, tr.e2.y, tr.e2.z);
	float3 p = xcross3(d, e2);
	TriTerms r;
	r.det = xdot3(e1, p);
	float3 s = xsub3(o, v0);
	r.su = xdot3(s, p);
	float3 q = xcross3(s, e1);
	r.sv = xdot3(d, q);
	r.st = xdot3(e2, q);
	return r;
}
// The oracle's decision sequence on the exact quotients.
__device__ __noinline__ bool tri_exact(TriTerms k, float& t) {
	if (k.det == 0.0f) return false;
	float u = __fdiv_rn(k.su, k.det);
	if (!(u >= 0.0f)) return false;
	float v = __fdiv_rn(k.sv, k.det);
	if (!(v >= 0.0f) || !(__fadd_rn(u, v) <= 1.0f)) return false;
	t = __fdiv_rn(k.st, k.det);
	return true;
}
// Returns true when the ray's supporting line crosses the triangle; t and det are outputs (oracle semantics).
__device__ __forceinline__ bool tri_test(const BvhTri& tr, float3 o, float3 d, float& t, float& det) {
	TriTerms k = tri_terms(tr, o, d);
	det = k.det;
	return tri_exact(k, t);
}

// Margins of the quick classification: the MUFU reciprocal is good to ~2^-22 relative, the quotients and their sum to a
// few 2^-23; 1e-5 is 40x that. Values whose magnitude is below kTiny could round to zero and are left to the exact path.
#define RL_TRI_EPS 1.0e-5f
#define RL_TRI_TINY 1.0e-30f

// Any-hit decision for t in the open interval (t_min, t_max): identical to `tri_test(...) && t > t_min && t < t_max`.
__device__ __forceinline__ bool tri_any_hit(const BvhTri& tr, float3 o, float3 d, float t_min, float t_max) {
	TriTerms k = tri_terms(tr, o, d);
	float r = approx_rcp(k.det);
	float ua = k.su * r, va = k.sv * r, ta = k.st * r, sa = ua + va;
	float mt = RL_TRI_EPS * fmaxf(fabsf(ta), t_max) + RL_TRI_TINY;
	// certain misses (the comparisons are false for NaN, which then falls through to the exact path or the certain-hit test)
	if (ua < -RL_TRI_TINY || va < -RL_TRI_TINY || sa > 1.0f + RL_TRI_EPS || ta < t_min - mt || ta > t_max + mt) return false;
	if (ua > RL_TRI_TINY && va > RL_TRI_TINY && sa < 1.0f - RL_TRI_EPS && ta > t_min + mt && ta < t_max - mt && fabsf(k.det) > RL_TRI_TINY) return true;
	float t;
	return tri_exact(k, t) && t > t_min && t < t_max;
}

__device__ __forceinline__ bool slab(float lox, float loy, float loz, float hix, float hiy, float hiz,
	float3 o, float3 inv, float t_min, float t_max, float& t_near)
{
	float ax = (lox - o.x) * inv.x, bx = (hix - o.x) * inv.x;
	float ay = (loy - o.y) * inv.y, by = (hiy - o.y) * inv.y;
	float az = (loz - o.z) * inv.z, bz = (hiz - o.z) * inv.z;
	float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), t_min));
	float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), t_max));
	t_near = t0;
	return t0 <= t1;
}

// Any-hit query with the open interval (t_min, t_max) (Vulkan culls triangle candidates
// with t <= tmin or t >= tmax), no face culling (scene.c:325), terminate on first hit.
__device__ bool bvh_any_hit(const SceneView& s, float3 o, float3 d, float t_min, float t_max) {
	if (!(t_min < t_max)) return false;
	float3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	int stack[RL_STACK];
	int sp = 0;
	int node = 0;
	while (true) {
		const BvhNode n = s.nodes[node];
		float tl, tr_;
		bool hl = slab(n.a.x, n.a.y, n.a.z, n.a.w, n.b.x, n.b.y, o, inv, t_min, t_max, tl);
		bool hr = slab(n.b.z, n.b.w, n.c.x, n.c.y, n.c.z, n.c.w, o, inv, t_min, t_max, tr_);
		int next = -1;
		#pragma unroll
		for (int side = 0; side != 2; ++side) {
			bool h = side ? hr : hl;
			int child = side ? n.d.y : n.d.x;
			if (!h) continue;
			if (child < 0) {
				uint32_t ref = ~(uint32_t) child;
				uint32_t first = ref >> 4, count = (ref & 15u) + 1u;
				for (uint32_t i = first; i != first + count; ++i)
					if (tri_any_hit(s.tris[i], o, d, t_min, t_max)) return true;
			}
			else if (next < 0) next = child;
			else stack[sp++] = child;
		}
		if (next >= 0) node = next;
		else if (sp) node = stack[--sp];
		else return false;
	}
}

// Nearest front-facing triangle along a primary ray that survives the depth clip
// 0 <= z_clip <= w_clip; ties in t go to the lower primitive id (same rule as the oracle).
// Returns the id word (index | emitter bit << 31) or 0xFFFFFFFF.
__device__ uint32_t bvh_closest_front(const SceneView& s, float3 o, float3 d, const float (*w2p)[4]) {
	float zo = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w2p[2][0], o.x), __fmul_rn(w2p[2][1], o.y)), __fmul_rn(w2p[2][2], o.z)), w2p[2][3]);
	float zd = __fadd_rn(__fadd_rn(__fmul_rn(w2p[2][0], d.x), __fmul_rn(w2p[2][1], d.y)), __fmul_rn(w2p[2][2], d.z));
	float wo = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w2p[3][0], o.x), __fmul_rn(w2p[3][1], o.y)), __fmul_rn(w2p[3][2], o.z)), w2p[3][3]);
	float wd = __fadd_rn(__fadd_rn(__fmul_rn(w2p[3][0], d.x), __fmul_rn(w2p[3][1], d.y)), __fmul_rn(w2p[3][2], d.z));
	float3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	float best_t = INFINITY;
	uint32_t best = 0xFFFFFFFFu, best_index = 0xFFFFFFFFu;
	int stack[RL_STACK];
	int sp = 0;
	int node = 0;
	while (true) {
		const BvhNode n = s.nodes[node];
		float tl, tr_;
		bool hl = slab(n.a.x, n.a.y, n.a.z, n.a.w, n.b.x, n.b.y, o, inv, 0.0f, best_t, tl);
		bool hr = slab(n.b.z, n.b.w, n.c.x, n.c.y, n.c.z, n.c.w, o, inv, 0.0f, best_t, tr_);
		int inner[2]; float inner_t[2]; int ni = 0;
		#pragma unroll
		for (int side = 0; side != 2; ++side) {
			bool h = side ? hr : hl;
			int child = side ? n.d.y : n.d.x;
			if (!h) continue;
			if (child < 0) {
				uint32_t ref = ~(uint32_t) child;
				uint32_t first = ref >> 4, count = (ref & 15u) + 1u;
				for (uint32_t i = first; i != first + count; ++i) {
					const BvhTri tri = s.tris[i];
					TriTerms k = tri_terms(tri, o, d);
					// back faces (det <= 0) are culled; a certain miss of the edge tests or a t certainly behind the eye or
					// certainly beyond the best hit cannot change the result
					if (!(k.det > 0.0f)) continue;
					float r = approx_rcp(k.det);
					float ua = k.su * r, va = k.sv * r, ta = k.st * r;
					if (ua < -RL_TRI_TINY || va < -RL_TRI_TINY || ua + va > 1.0f + RL_TRI_EPS || ta < -RL_TRI_TINY || ta > best_t + RL_TRI_EPS * fabsf(ta) + RL_TRI_TINY) continue;
					float t;
					if (!tri_exact(k, t)) continue;
					if (!(t > 0.0f)) continue;
					float zc = __fadd_rn(zo, __fmul_rn(t, zd)), wc = __fadd_rn(wo, __fmul_rn(t, wd));
					if (!(zc >= 0.0f) || !(zc <= wc)) continue;
					uint32_t id = __float_as_uint(tri.v0.w), index = id & 0x7FFFFFFFu;
					if (t < best_t || (t == best_t && index < best_index)) { best_t = t; best = id; best_index = index; }
				}
			}
			else { inner[ni] = child; inner_t[ni] = side ? tr_ : tl; ++ni; }
		}
		if (ni == 2) {
			bool swap = inner_t[1] < inner_t[0];
			stack[sp++] = swap ? inner[0] : inner[1];
			node = swap ? inner[1] : inner[0];
		}
		else if (ni == 1) node = inner[0];
		else if (sp) node = stack[--sp];
		else return best;
	}
}

}  // namespace RL_NS
