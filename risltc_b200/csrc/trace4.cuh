// trace4.cuh -- kernel (3), shadow rays (get_polygon_visibility, shading_pass.frag.glsl:112-129), second generation.
//
// Why it replaces trace_kernel (kernels.cuh) as the default: ncu showed the binary-tree kernel bound by the L1 data
// pipe (l1tex__data_pipe_lsu_wavefronts 90 % of peak, profiles/r1_ncu_trace_kernel.txt): every lane fetched 64 bytes
// per TWO child boxes with four 16-byte loads, each of which costs a tag lookup per distinct node in the warp, and kept
// its stacks in local memory, where lanes with different stack heights hit different lines. Here
//   * a node holds FOUR children in the same 64 bytes (Qbvh4Node: boxes quantised to 8 bits relative to the node), so a
//     ray needs half the node visits, half the dependent round trips and half the tag lookups;
//   * a plane is decoded with one PRMT (the byte becomes the mantissa of a float in [1, 2)) and evaluated with one FFMA
//     whose scale and offset are computed once per node; near / far planes are picked per axis by the sign of the ray
//     direction, so a child costs 6 PRMT + 6 FFMA + 2 FMNMX3 + 2 FMNMX, and its hit is dispatched without a branch;
//   * both stacks live in shared memory, one column per lane (bank = lane: one wavefront whatever the heights);
//     a node stack deeper than RL_T4_NSTACK spills to local memory (never in the test scenes);
//   * the triangle test and the decision rules are those of the oracle (tri_any_hit, bvh.cuh): hit / no-hit is
//     independent of the shape of the tree, so the image is bit-identical to the binary-tree kernel's.
#pragma once
#include "kernels.cuh"

namespace RL_NS {

#define RL_T4_NSTACK 16     // node stack entries per lane in shared memory
#define RL_T4_LSTACK 12     // leaf stack entries per lane in shared memory (the node track pauses when < 4 are free)
#define RL_T4_OVERFLOW 88   // node stack entries per lane in local memory (spilled from / refilled into the shared part 8 at a time)

__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
// plane byte k of `word` -> float 1 + q * 2^-15. `one` holds the bits of 1.0f IN A REGISTER: PRMT takes one immediate, and
// with the constant written here the compiler spends it on the 1.0f and rebuilds the selector in a register before each of
// the 24 PRMTs of a node (ncu source view, round 2); the kernels receive the bits as a launch parameter instead.
template <int CHILD>
__device__ __forceinline__ float q4_plane(uint32_t word, uint32_t one) {
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(one), "n"(0x7604 | (CHILD << 4)));
	return __uint_as_float(r);
}
template <int C> struct ChildIndex { static constexpr int value = C; };

// COUNT = true additionally counts the rays, node visits, triangle tests and occluded rays of the launch into
// px.counters[4..7] (bench.py's roofline_trace block: bytes per ray); the frame path runs COUNT = false unless asked.
template <bool COUNT>
__global__ void __launch_bounds__(128) trace4_kernel(SceneView s, PixelBuffers px, uint32_t ray_count, uint32_t tri_vote, uint32_t refill, uint32_t one) {
	__shared__ float4 sm_stage[4][RL_TRACE_STAGE][2];
	__shared__ int sm_nstack[RL_T4_NSTACK][128];
	__shared__ int sm_lstack[RL_T4_LSTACK][128];
	int overflow[RL_T4_OVERFLOW];
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	float4 (*stage)[2] = sm_stage[warp];
	uint32_t stage_next = 0, stage_count = 0;   // warp-uniform
	bool exhausted = false, busy = false;
	uint32_t ray = 0, tri_i = 0, tri_end = 0;
	float3 o = mk3(0.0f, 0.0f, 0.0f), d = mk3(0.0f, 0.0f, 1.0f), inv = mk3(0.0f, 0.0f, 0.0f), oi = mk3(0.0f, 0.0f, 0.0f);
	float t_max = 0.0f;
	int node = -1, nsp = 0, lsp = 0, spilled = 0;
	const float t_min = 1.0e-3f;
	uint32_t ahead = 0;
	uint32_t n_rays = 0, n_nodes = 0, n_tris = 0, n_occluded = 0;
	if (lane == 0) ahead = atomicAdd(px.ticket, RL_TRACE_STAGE);
	ahead = __shfl_sync(0xFFFFFFFFu, ahead, 0);
	while (true) {
		const unsigned idle = __ballot_sync(0xFFFFFFFFu, !busy);
		if ((uint32_t) __popc(idle) >= refill || (idle && exhausted)) {
			while (stage_next >= stage_count && !exhausted) {
				const uint32_t base = ahead;
				if (base >= ray_count) { exhausted = true; break; }
				if (lane == 0) ahead = atomicAdd(px.ticket, RL_TRACE_STAGE);
				ahead = __shfl_sync(0xFFFFFFFFu, ahead, 0);
				if (ahead + lane * 2u < ray_count) {
					prefetch_l2(&px.ray_a[ahead + lane * 2u]); prefetch_l2(&px.ray_b[ahead + lane * 2u]);
					prefetch_l2(&px.origin[(ahead + lane * 2u) % px.pixel_count]);
				}
				stage_next = stage_count = 0u;
				#pragma unroll
				for (uint32_t half = 0; half != 2u; ++half) {
					const uint32_t r = base + half * 32u + lane;
					bool valid = r < ray_count && __ldg(&((const float*) px.ray_b)[4 * (size_t) r + 3]) == 1.0f;
					float4 ra = make_float4(0.0f, 0.0f, 1.0f, 0.0f), og = ra;
					if (valid) { ra = px.ray_a[r]; og = px.origin[r % px.pixel_count]; valid = t_min < ra.w; }
					const unsigned have = __ballot_sync(0xFFFFFFFFu, valid);
					if (valid) {
						const uint32_t slot = stage_count + __popc(have & ((1u << lane) - 1u));
						stage[slot][0] = make_float4(og.x, og.y, og.z, ra.w);
						stage[slot][1] = make_float4(ra.x, ra.y, ra.z, __uint_as_float(r));
					}
					stage_count += (uint32_t) __popc(have);
				}
				__syncwarp();
			}
			if (exhausted && stage_next >= stage_count && idle == 0xFFFFFFFFu) break;
			if (!busy) {
				const uint32_t mine = stage_next + __popc(idle & ((1u << lane) - 1u));
				if (mine < stage_count) {
					const float4 a0 = stage[mine][0], a1 = stage[mine][1];
					ray = __float_as_uint(a1.w); busy = true;
					if (COUNT) ++n_rays;
					o = mk3(a0.x, a0.y, a0.z); d = mk3(a1.x, a1.y, a1.z); t_max = a0.w;
					// box tests only: the error of the approximate reciprocal is covered by the outward rounding of the boxes
					inv = mk3(box_reciprocal(d.x), box_reciprocal(d.y), box_reciprocal(d.z));
					oi = mk3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
					node = 0; nsp = 0; lsp = 0; spilled = 0; tri_i = tri_end = 0u;
				}
			}
			stage_next = min(stage_next + (uint32_t) __popc(idle), stage_count);
			__syncwarp();
		}
		const bool node_ready = busy && node >= 0 && lsp <= RL_T4_LSTACK - 4;
		const bool tri_pending = busy && (tri_i != tri_end || lsp != 0);
		const unsigned node_votes = __ballot_sync(0xFFFFFFFFu, node_ready), tri_votes = __ballot_sync(0xFFFFFFFFu, tri_pending);
		const bool run_tri = (uint32_t) __popc(tri_votes) >= tri_vote || node_votes == 0u;
		// ---- track A: one node = four child boxes. No divergent branches: every child stores its reference to the top of
		// both stacks and CLAIMS the slot only if it belongs there; the last inner hit stays in a register as the next node.
		if (node_ready) {
			if (COUNT) ++n_nodes;
			const Qbvh4Node* np = s.nodes4 + node;
			const uint4 na = __ldg(&np->a), nb = __ldg(&np->b), nc = __ldg(&np->c);
			const int4 refs = __ldg(&np->refs);
			// plane = origin + q * step; with v = 1 + q * 2^-15 (q4_plane): t = v * S + B, S = (2^15 step) / d, B = (origin - o) / d - S
			const float sx = __uint_as_float(na.w) * inv.x, sy = __uint_as_float(nc.z) * inv.y, sz = __uint_as_float(nc.w) * inv.z;
			const float bx = fmaf(__uint_as_float(na.x), inv.x, -oi.x) - sx;
			const float by = fmaf(__uint_as_float(na.y), inv.y, -oi.y) - sy;
			const float bz = fmaf(__uint_as_float(na.z), inv.z, -oi.z) - sz;
			// the plane the ray enters through is the low one where the direction is positive
			const bool neg_x = inv.x < 0.0f, neg_y = inv.y < 0.0f, neg_z = inv.z < 0.0f;
			const uint32_t near_x = neg_x ? nb.w : nb.x, far_x = neg_x ? nb.x : nb.w;
			const uint32_t near_y = neg_y ? nc.x : nb.y, far_y = neg_y ? nb.y : nc.x;
			const uint32_t near_z = neg_z ? nc.y : nb.z, far_z = neg_z ? nb.z : nc.y;
			if (nsp > RL_T4_NSTACK - 4) {
				// (almost) never: the shared part of the node stack could overflow -> move its 8 oldest entries to local memory
				#pragma unroll 1
				for (int i = 0; i != 8; ++i) overflow[spilled + i] = sm_nstack[i][tid];
				#pragma unroll 1
				for (int i = 8; i < nsp; ++i) sm_nstack[i - 8][tid] = sm_nstack[i][tid];
				spilled += 8; nsp -= 8;
			}
			int next = -1;
			float next_t = 0.0f;
			auto child = [&](auto index) {
				constexpr int c = decltype(index)::value;
				const float t0 = fmaxf(max3(fmaf(q4_plane<c>(near_x, one), sx, bx), fmaf(q4_plane<c>(near_y, one), sy, by), fmaf(q4_plane<c>(near_z, one), sz, bz)), t_min);
				const float t1 = fminf(min3(fmaf(q4_plane<c>(far_x, one), sx, bx), fmaf(q4_plane<c>(far_y, one), sy, by), fmaf(q4_plane<c>(far_z, one), sz, bz)), t_max);
				const int ref = (c == 0) ? refs.x : (c == 1) ? refs.y : (c == 2) ? refs.z : refs.w;
				const bool hit = t0 <= t1 && ref != RL_Q4_EMPTY;
				const bool leaf = hit && ref < 0, inner = hit && ref >= 0;
				sm_lstack[lsp][tid] = ref;
				lsp += leaf ? 1 : 0;
				// of the inner children that are hit, the one the ray enters first is visited next (occluded rays end sooner);
				// whichever of {next, ref} loses goes on the stack
				const bool push = inner && next >= 0;
				const bool closer = inner && (next < 0 || t0 < next_t);
				sm_nstack[nsp][tid] = closer ? next : ref;
				nsp += push ? 1 : 0;
				next = closer ? ref : next;
				next_t = closer ? t0 : next_t;
			};
			child(ChildIndex<0>()); child(ChildIndex<1>()); child(ChildIndex<2>()); child(ChildIndex<3>());
			if (next < 0 && nsp == 0 && spilled != 0) {
				spilled -= 8; nsp = 8;
				#pragma unroll 1
				for (int i = 0; i != 8; ++i) sm_nstack[i][tid] = overflow[spilled + i];
			}
			const bool pop = next < 0 && nsp > 0;
			nsp -= pop ? 1 : 0;
			const int top = sm_nstack[nsp][tid];
			node = pop ? top : next;
		}
		// ---- track B: one triangle
		if (run_tri && tri_pending) {
			if (tri_i == tri_end) {
				const uint32_t ref = ~(uint32_t) sm_lstack[--lsp][tid];
				tri_i = ref >> 4; tri_end = tri_i + (ref & 15u) + 1u;
			}
			if (COUNT) ++n_tris;
			if (tri_any_hit(s.tris[tri_i], o, d, t_min, t_max)) {
				((float*) px.ray_b)[4 * (size_t) ray + 3] = 2.0f;
				busy = false;
				if (COUNT) ++n_occluded;
			}
			++tri_i;
		}
		if (busy && node < 0 && tri_i == tri_end && lsp == 0) busy = false;   // nothing left on either track: the ray reaches the light
	}
	if (COUNT) {
		n_rays = __reduce_add_sync(0xFFFFFFFFu, n_rays); n_nodes = __reduce_add_sync(0xFFFFFFFFu, n_nodes);
		n_tris = __reduce_add_sync(0xFFFFFFFFu, n_tris); n_occluded = __reduce_add_sync(0xFFFFFFFFu, n_occluded);
		if (lane == 0) {
			atomicAdd(&px.counters[4], (unsigned long long) n_rays); atomicAdd(&px.counters[5], (unsigned long long) n_nodes);
			atomicAdd(&px.counters[6], (unsigned long long) n_tris); atomicAdd(&px.counters[7], (unsigned long long) n_occluded);
		}
	}
}


// ---------------------------------------------------------------------------------------------------------------------
// trace4p_kernel: TWO rays per lane. Ray slots 2m and 2m + 1 of a pixel are the two MIS techniques' samples of the SAME
// light seen from the SAME shading point (shading_pass.frag.glsl:373-394 casts them one after the other), so they share
// their origin and most of the nodes they visit. A lane traverses the tree once for the pair: the node is fetched and its
// 24 plane bytes decoded once (PRMT, the ALU pipe that bounds trace4_kernel), both rays evaluate the planes with their own
// scale / offset (FFMA, the pipe with slack), a child is entered when EITHER ray hits it, and a triangle is fetched once and
// tested against both (s = o - v0, q = s x e1 and e2 . q do not depend on the direction and are computed once, with the
// very operations tri_terms uses). The any-hit decision of a ray does not depend on the order or the set of boxes visited
// beyond those it hits itself, so every ray's result is bit-identical to trace4_kernel's. A ray that found its occluder
// leaves the pair (t_max = -1 fails every box and triangle test). The near / far plane of an axis is picked by the sign of
// the direction, so only rays of the same octant form a pair; the others (and single rays) run alone in slot A.
#define RL_T4P_STAGE 32u   // pairs staged per warp and refill (each may become two entries)

struct TriShared { float3 s, q; float st; };
__device__ __forceinline__ TriShared tri_shared(const BvhTri& tr, float3 o) {
	TriShared r;
	r.s = xsub3(o, mk3(tr.v0.x, tr.v0.y, tr.v0.z));
	r.q = xcross3(r.s, mk3(tr.e1.x, tr.e1.y, tr.e1.z));
	r.st = xdot3(mk3(tr.e2.x, tr.e2.y, tr.e2.z), r.q);
	return r;
}
// tri_any_hit (bvh.cuh) on the shared terms: the same operations in the same order
__device__ __forceinline__ bool tri_any_hit_shared(const BvhTri& tr, const TriShared& h, float3 d, float t_min, float t_max) {
	const float3 e1 = mk3(tr.e1.x, tr.e1.y, tr.e1.z), e2 = mk3(tr.e2.x, tr.e2.y, tr.e2.z);
	const float3 p = xcross3(d, e2);
	TriTerms k;
	k.det = xdot3(e1, p);
	k.su = xdot3(h.s, p);
	k.sv = xdot3(d, h.q);
	k.st = h.st;
	const float r = approx_rcp(k.det);
	const float ua = k.su * r, va = k.sv * r, ta = k.st * r, sa = ua + va;
	const float mt = RL_TRI_EPS * fmaxf(fabsf(ta), t_max) + RL_TRI_TINY;
	if (ua < -RL_TRI_TINY || va < -RL_TRI_TINY || sa > 1.0f + RL_TRI_EPS || ta < t_min - mt || ta > t_max + mt) return false;
	if (ua > RL_TRI_TINY && va > RL_TRI_TINY && sa < 1.0f - RL_TRI_EPS && ta > t_min + mt && ta < t_max - mt && fabsf(k.det) > RL_TRI_TINY) return true;
	float t;
	return tri_exact(k, t) && t > t_min && t < t_max;
}

// pair_count = pixel_count * (ray slots / 2); pair p = m * pixel_count + pixel holds the rays 2 m * pixel_count + pixel (slot A)
// and that + pixel_count (slot B). COUNT: px.counters[4..7] += rays, node visits, triangle fetches, occluded rays.
// ORDERED: of the inner children that are hit, the one ray A enters first is visited next (occluded rays end sooner: C3 -10 %,
// C4 -12 % kernel time); ORDERED = false pushes them all and pops the last -- seven instructions less per child, which wins
// where few rays are occluded (C2: 27 % occluded, -5.5 %). The frame path times both on the first frames after a scene upload.
template <bool COUNT, bool ORDERED = true>
__global__ void __launch_bounds__(128) trace4p_kernel(SceneView s, PixelBuffers px, uint32_t pair_count, uint32_t tri_vote, uint32_t refill, uint32_t one) {
	__shared__ float4 sm_stage[4][2 * RL_T4P_STAGE][3];
	__shared__ int sm_nstack[RL_T4_NSTACK][128];
	__shared__ int sm_lstack[RL_T4_LSTACK][128];
	int overflow[RL_T4_OVERFLOW];
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	float4 (*stage)[3] = sm_stage[warp];
	const uint32_t pc = px.pixel_count;
	uint32_t stage_next = 0, stage_count = 0;   // warp-uniform
	bool exhausted = false, busy = false;
	uint32_t ray = 0, tri_i = 0, tri_end = 0;
	float3 o = mk3(0.0f, 0.0f, 0.0f), da = mk3(0.0f, 0.0f, 1.0f), db = da, inva = mk3(0.0f, 0.0f, 0.0f), invb = inva;
	float tmax_a = -1.0f, tmax_b = -1.0f;
	int node = -1, nsp = 0, lsp = 0, spilled = 0;
	const float t_min = 1.0e-3f;
	uint32_t ahead = 0;
	uint32_t n_rays = 0, n_nodes = 0, n_tris = 0, n_occluded = 0;
	if (lane == 0) ahead = atomicAdd(px.ticket, RL_T4P_STAGE);
	ahead = __shfl_sync(0xFFFFFFFFu, ahead, 0);
	while (true) {
		const unsigned idle = __ballot_sync(0xFFFFFFFFu, !busy);
		if ((uint32_t) __popc(idle) >= refill || (idle && exhausted)) {
			while (stage_next >= stage_count && !exhausted) {
				const uint32_t base = ahead;
				if (base >= pair_count) { exhausted = true; break; }
				if (lane == 0) ahead = atomicAdd(px.ticket, RL_T4P_STAGE);
				ahead = __shfl_sync(0xFFFFFFFFu, ahead, 0);
				if (ahead + lane < pair_count && (lane & 3u) == 0u) {
					const uint32_t m = (ahead + lane) / pc, r = ahead + lane + m * pc;
					prefetch_l2(&px.ray_a[r]); prefetch_l2(&px.ray_b[r]); prefetch_l2(&px.ray_a[r + pc]); prefetch_l2(&px.ray_b[r + pc]);
					prefetch_l2(&px.origin[r - 2u * m * pc]);
				}
				stage_next = stage_count = 0u;
				const uint32_t p = base + lane;
				const bool in = p < pair_count;
				const uint32_t m = in ? p / pc : 0u, r0 = p + m * pc, r1 = r0 + pc;
				bool va = in && __ldg(&((const float*) px.ray_b)[4 * (size_t) r0 + 3]) == 1.0f;
				bool vb = in && __ldg(&((const float*) px.ray_b)[4 * (size_t) r1 + 3]) == 1.0f;
				float4 ra = make_float4(0.0f, 0.0f, 1.0f, 0.0f), rb = ra, og = ra;
				if (va) { ra = px.ray_a[r0]; va = t_min < ra.w; }
				if (vb) { rb = px.ray_a[r1]; vb = t_min < rb.w; }
				if (va || vb) og = px.origin[r0 - 2u * m * pc];
				const bool same_octant = (((__float_as_uint(ra.x) ^ __float_as_uint(rb.x)) | (__float_as_uint(ra.y) ^ __float_as_uint(rb.y)) | (__float_as_uint(ra.z) ^ __float_as_uint(rb.z))) >> 31) == 0u;
				const bool both = va && vb && same_octant, split = va && vb && !same_octant;
				// entry 1: the pair, or ray A alone, or (A invalid) ray B alone in slot A; entry 2 (split pairs): ray B alone
				const unsigned have1 = __ballot_sync(0xFFFFFFFFu, va || vb), have2 = __ballot_sync(0xFFFFFFFFu, split);
				if (va || vb) {
					const uint32_t slot = __popc(have1 & lt_mask);
					const float4 first = va ? ra : rb;
					stage[slot][0] = make_float4(og.x, og.y, og.z, first.w);
					stage[slot][1] = make_float4(first.x, first.y, first.z, both ? rb.w : -1.0f);
					stage[slot][2] = both ? make_float4(rb.x, rb.y, rb.z, __uint_as_float(r0)) : make_float4(first.x, first.y, first.z, __uint_as_float(va ? r0 : r1));
				}
				if (split) {
					const uint32_t slot = __popc(have1) + __popc(have2 & lt_mask);
					stage[slot][0] = make_float4(og.x, og.y, og.z, rb.w);
					stage[slot][1] = make_float4(rb.x, rb.y, rb.z, -1.0f);
					stage[slot][2] = make_float4(rb.x, rb.y, rb.z, __uint_as_float(r1));
				}
				stage_count = (uint32_t) (__popc(have1) + __popc(have2));
				__syncwarp();
			}
			if (exhausted && stage_next >= stage_count && idle == 0xFFFFFFFFu) break;
			if (!busy) {
				const uint32_t mine = stage_next + __popc(idle & lt_mask);
				if (mine < stage_count) {
					const float4 a0 = stage[mine][0], a1 = stage[mine][1], a2 = stage[mine][2];
					ray = __float_as_uint(a2.w); busy = true;
					o = mk3(a0.x, a0.y, a0.z); da = mk3(a1.x, a1.y, a1.z); db = mk3(a2.x, a2.y, a2.z); tmax_a = a0.w; tmax_b = a1.w;
					if (COUNT) n_rays += (tmax_b > 0.0f) ? 2u : 1u;
					// box tests only: the error of the approximate reciprocal is covered by the outward rounding of the boxes
					inva = mk3(box_reciprocal(da.x), box_reciprocal(da.y), box_reciprocal(da.z));
					invb = mk3(box_reciprocal(db.x), box_reciprocal(db.y), box_reciprocal(db.z));
					node = 0; nsp = 0; lsp = 0; spilled = 0; tri_i = tri_end = 0u;
				}
			}
			stage_next = min(stage_next + (uint32_t) __popc(idle), stage_count);
			__syncwarp();
		}
		const bool node_ready = busy && node >= 0 && lsp <= RL_T4_LSTACK - 4;
		const bool tri_pending = busy && (tri_i != tri_end || lsp != 0);
		const unsigned node_votes = __ballot_sync(0xFFFFFFFFu, node_ready), tri_votes = __ballot_sync(0xFFFFFFFFu, tri_pending);
		const bool run_tri = (uint32_t) __popc(tri_votes) >= tri_vote || node_votes == 0u;
		// ---- track A: one node = four child boxes, both rays (see trace4_kernel for the branch-free bookkeeping)
		if (node_ready) {
			if (COUNT) ++n_nodes;
			const Qbvh4Node* np = s.nodes4 + node;
			const uint4 na = __ldg(&np->a), nb = __ldg(&np->b), nc = __ldg(&np->c);
			const int4 refs = __ldg(&np->refs);
			const float ox = __uint_as_float(na.x), oy = __uint_as_float(na.y), oz = __uint_as_float(na.z);
			const float sax = __uint_as_float(na.w) * inva.x, say = __uint_as_float(nc.z) * inva.y, saz = __uint_as_float(nc.w) * inva.z;
			const float sbx = __uint_as_float(na.w) * invb.x, sby = __uint_as_float(nc.z) * invb.y, sbz = __uint_as_float(nc.w) * invb.z;
			const float rx = ox - o.x, ry = oy - o.y, rz = oz - o.z;
			const float bax = fmaf(rx, inva.x, -sax), bay = fmaf(ry, inva.y, -say), baz = fmaf(rz, inva.z, -saz);
			const float bbx = fmaf(rx, invb.x, -sbx), bby = fmaf(ry, invb.y, -sby), bbz = fmaf(rz, invb.z, -sbz);
			const bool neg_x = inva.x < 0.0f, neg_y = inva.y < 0.0f, neg_z = inva.z < 0.0f;
			const uint32_t near_x = neg_x ? nb.w : nb.x, far_x = neg_x ? nb.x : nb.w;
			const uint32_t near_y = neg_y ? nc.x : nb.y, far_y = neg_y ? nb.y : nc.x;
			const uint32_t near_z = neg_z ? nc.y : nb.z, far_z = neg_z ? nb.z : nc.y;
			if (nsp > RL_T4_NSTACK - 4) {
				#pragma unroll 1
				for (int i = 0; i != 8; ++i) overflow[spilled + i] = sm_nstack[i][tid];
				#pragma unroll 1
				for (int i = 8; i < nsp; ++i) sm_nstack[i - 8][tid] = sm_nstack[i][tid];
				spilled += 8; nsp -= 8;
			}
			int next = -1;
			float next_t = 0.0f;
			auto child = [&](auto index) {
				constexpr int c = decltype(index)::value;
				const float nx = q4_plane<c>(near_x, one), ny = q4_plane<c>(near_y, one), nz = q4_plane<c>(near_z, one);
				const float fx = q4_plane<c>(far_x, one), fy = q4_plane<c>(far_y, one), fz = q4_plane<c>(far_z, one);
				const float t0a = fmaxf(max3(fmaf(nx, sax, bax), fmaf(ny, say, bay), fmaf(nz, saz, baz)), t_min);
				const float t1a = fminf(min3(fmaf(fx, sax, bax), fmaf(fy, say, bay), fmaf(fz, saz, baz)), tmax_a);
				const float t0b = fmaxf(max3(fmaf(nx, sbx, bbx), fmaf(ny, sby, bby), fmaf(nz, sbz, bbz)), t_min);
				const float t1b = fminf(min3(fmaf(fx, sbx, bbx), fmaf(fy, sby, bby), fmaf(fz, sbz, bbz)), tmax_b);
				const int ref = (c == 0) ? refs.x : (c == 1) ? refs.y : (c == 2) ? refs.z : refs.w;
				const bool hit_a = t0a <= t1a;
				const bool hit = (hit_a || t0b <= t1b) && ref != RL_Q4_EMPTY;
				const float t0 = hit_a ? t0a : t0b;
				const bool leaf = hit && ref < 0, inner = hit && ref >= 0;
				sm_lstack[lsp][tid] = ref;
				lsp += leaf ? 1 : 0;
				if (ORDERED) {
					const bool push = inner && next >= 0;
					const bool closer = inner && (next < 0 || t0 < next_t);
					sm_nstack[nsp][tid] = closer ? next : ref;
					nsp += push ? 1 : 0;
					next = closer ? ref : next;
					next_t = closer ? t0 : next_t;
				}
				else {
					sm_nstack[nsp][tid] = ref;
					nsp += inner ? 1 : 0;
				}
			};
			child(ChildIndex<0>()); child(ChildIndex<1>()); child(ChildIndex<2>()); child(ChildIndex<3>());
			if (next < 0 && nsp == 0 && spilled != 0) {
				spilled -= 8; nsp = 8;
				#pragma unroll 1
				for (int i = 0; i != 8; ++i) sm_nstack[i][tid] = overflow[spilled + i];
			}
			const bool pop = next < 0 && nsp > 0;
			nsp -= pop ? 1 : 0;
			const int top = sm_nstack[nsp][tid];
			node = pop ? top : next;
		}
		// ---- track B: one triangle, both rays
		if (run_tri && tri_pending) {
			if (tri_i == tri_end) {
				const uint32_t ref = ~(uint32_t) sm_lstack[--lsp][tid];
				tri_i = ref >> 4; tri_end = tri_i + (ref & 15u) + 1u;
			}
			if (COUNT) ++n_tris;
			const BvhTri tr = s.tris[tri_i];
			const TriShared h = tri_shared(tr, o);
			if (tri_any_hit_shared(tr, h, da, t_min, tmax_a)) {
				((float*) px.ray_b)[4 * (size_t) ray + 3] = 2.0f;
				tmax_a = -1.0f;
				if (COUNT) ++n_occluded;
			}
			if (tri_any_hit_shared(tr, h, db, t_min, tmax_b)) {
				((float*) px.ray_b)[4 * (size_t) (ray + pc) + 3] = 2.0f;
				tmax_b = -1.0f;
				if (COUNT) ++n_occluded;
			}
			++tri_i;
		}
		// both rays occluded, or nothing left on either track (the remaining rays reach the light)
		if (busy && ((tmax_a < 0.0f && tmax_b < 0.0f) || (node < 0 && tri_i == tri_end && lsp == 0))) busy = false;
	}
	if (COUNT) {
		n_rays = __reduce_add_sync(0xFFFFFFFFu, n_rays); n_nodes = __reduce_add_sync(0xFFFFFFFFu, n_nodes);
		n_tris = __reduce_add_sync(0xFFFFFFFFu, n_tris); n_occluded = __reduce_add_sync(0xFFFFFFFFu, n_occluded);
		if (lane == 0) {
			atomicAdd(&px.counters[4], (unsigned long long) n_rays); atomicAdd(&px.counters[5], (unsigned long long) n_nodes);
			atomicAdd(&px.counters[6], (unsigned long long) n_tris); atomicAdd(&px.counters[7], (unsigned long long) n_occluded);
		}
	}
}

}  // namespace RL_NS
