// kernels.cuh -- the four passes of one frame as sm_100a kernels:
//   (1) gbuffer_kernel      primary visibility      (visibility_pass.*, main.c:2048)
//   (2) shade_kernel        RIS + LTC + PSA shading (shading_pass.frag.glsl:674-770, main.c:2055)
//   (3)+(4) resolve_kernel  shadow rays (shading_pass.frag.glsl:112-129), MIS sum, NaN guard, exposure
//                           and the running mean of accum_pass.frag.glsl:45-53
// Pixels are mapped to threads in 8x4 tiles per warp (16x8 per 128-thread block) so that a warp's
// shading points, BVH paths and framebuffer lines stay close together.
#pragma once
#include "shading.cuh"
#include "bvh.cuh"

namespace RL_NS {


__device__ __forceinline__ bool tile_pixel(const FrameUniforms& f, const Stripes& st, uint32_t& x, uint32_t& local_row, uint32_t& y) {
	uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	x = blockIdx.x * 16u + (warp & 1u) * 8u + (lane & 7u);
	local_row = blockIdx.y * 8u + (warp >> 1) * 4u + (lane >> 3);
	if (x >= f.width || local_row >= st.owned_rows) return false;
	y = st.global_row(local_row);
	return y < f.height;
}

__device__ __forceinline__ float3 primary_ray(const FrameUniforms& f, uint32_t x, uint32_t y) {
	float px = (float) (int) x, py = (float) (int) y;
	return mk3(__fadd_rn(__fadd_rn(__fmul_rn(f.pixel_to_ray[0][0], px), __fmul_rn(f.pixel_to_ray[0][1], py)), f.pixel_to_ray[0][2]),
	           __fadd_rn(__fadd_rn(__fmul_rn(f.pixel_to_ray[1][0], px), __fmul_rn(f.pixel_to_ray[1][1], py)), f.pixel_to_ray[1][2]),
	           __fadd_rn(__fadd_rn(__fmul_rn(f.pixel_to_ray[2][0], px), __fmul_rn(f.pixel_to_ray[2][1], py)), f.pixel_to_ray[2][2]));
}

// ---------------------------------------------------------------- (1)
__global__ void __launch_bounds__(128) gbuffer_kernel(SceneView s, FrameUniforms f, Stripes st, PixelBuffers out) {
	uint32_t x, row, y;
	if (!tile_pixel(f, st, x, row, y)) return;
	float3 d = primary_ray(f, x, y);
	out.visibility[row * f.width + x] = bvh_closest_front(s, mk3(f.camera[0], f.camera[1], f.camera[2]), d, f.world_to_projection);
}

// ---------------------------------------------------------------- (2)
// Ray handling policy of the shading kernel. DEFER = true: shadow rays are written out for the
// traversal kernel (legal when every ray only gates its own term). DEFER = false: traced in place
// (Turk baseline, whose visibility gates the whole light sample, and optimal MIS, whose occluded
// samples still contribute).
template <int V, bool DEFER>
struct ShadeContext {
	const SceneView& s; const FrameUniforms& f; const Variant& var; const PixelBuffers& out;
	uint32_t pixel;      // local pixel index
	uint32_t seed;       // noise_accessor_t.seed
	uint32_t rays;       // rays issued by this pixel
	__device__ float next() { return noise_next(seed); }
};

// evaluate_polygonal_light_shading_peters / the PSA branch of evaluate_polygonal_light_shading for the
// final light sample `group`: draws the technique samples and either records or traces their rays.
// candidate = true: PSA target function of a RIS candidate (visibility not requested, no rays).
template <int V, bool DEFER>
__device__ float3 sample_light(ShadeContext<V, DEFER>& c, const ShadingPoint& sp, const LtcFrame& ltc, const Light<V>& light,
	bool peters, bool candidate, uint32_t group)
{
	Techniques<V> t;
	prepare_techniques<V>(t, sp, ltc, light, c.var);
	if (!t.valid) return mk3(0.0f, 0.0f, 0.0f);
	const bool fast = c.var.fast_atan != 0, biased = c.var.polygon_technique == TECH_PSA_BIASED;
	const uint32_t S = c.var.sample_count;
	float3 result = mk3(0.0f, 0.0f, 0.0f);
	for (uint32_t smp = 0; smp != S; ++smp) {
		float u0 = c.next(), u1 = c.next();
		float3 dir[2];
		dir[0] = psa_sample_rt<V + 1>(t.diffuse, u0, u1, fast, biased);
		dir[1] = mk3(0.0f, 0.0f, 0.0f);
		int techniques = 1;
		if (t.specular.total > 0.0f) {
			u0 = c.next(); u1 = c.next();
			dir[1] = cosine_to_shading_dir(ltc, psa_sample_rt<V + 1>(t.specular, u0, u1, fast, biased));
			techniques = 2;
		}
		for (int j = 0; j != techniques; ++j) {
			RayRequest ray; bool side_visible; float3 if_occluded;
			ray.dir = mk3(0.0f, 0.0f, 1.0f); ray.t_max = -1.0f; ray.if_visible = mk3(0.0f, 0.0f, 0.0f);
			uint32_t slot = (group * S + smp) * 2u + (uint32_t) j;
			bool have = technique_sample<V>(t, sp, ltc, light, c.var, j, dir[j], c.f.mis_visibility_estimate, peters, ray, side_visible, if_occluded);
			if (!have) continue;
			if (candidate) { result = add3(result, side_visible ? ray.if_visible : if_occluded); continue; }
			if (!side_visible) { result = add3(result, if_occluded); continue; }
			if (DEFER) {
				c.out.ray_a[(size_t) slot * c.out.pixel_count + c.pixel] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, ray.t_max);
				c.out.ray_b[(size_t) slot * c.out.pixel_count + c.pixel] = make_float4(ray.if_visible.x, ray.if_visible.y, ray.if_visible.z, 1.0f);
				c.rays++;
			}
			else {
				c.rays++;
				bool visible = !bvh_any_hit(c.s, sp.position, ray.dir, 1.0e-3f, ray.t_max);
				result = add3(result, visible ? ray.if_visible : if_occluded);
			}
		}
	}
	// DEFER: `result` is the carry (terms that needed no ray); the caller stores it un-normalised and the
	// resolve kernel applies 1 / SAMPLE_COUNT after adding the visible terms, in the reference's order.
	if (DEFER && !candidate) return result;
	return scale3(result, 1.0f / (float) S);
}

// Turk baseline: get_polygonal_light_mis_estimate over SAMPLE_COUNT area samples
// (shading_pass.frag.glsl:411-418, :281-287, :200-207). `visibility` has the reference's inout semantics.
template <int V, bool DEFER>
__device__ float3 turk_light(ShadeContext<V, DEFER>& c, const ShadingPoint& sp, const Light<V>& light, float3& light_sample, bool eval_only, bool& visibility) {
	float3 result = mk3(0.0f, 0.0f, 0.0f);
	const uint32_t S = c.var.sample_count;
	for (uint32_t smp = 0; smp != S; ++smp) {
		if (!eval_only) { float u0 = c.next(), u1 = c.next(); light_sample = turk_sample<V>(light, u0, u1); }
		float3 d = sub3(light_sample, sp.position);
		float dist2 = dot3(d, d);
		d = scale3(d, inversesqrt(dist2));
		float projected = fabsf(dot3(mk3(light.plane.x, light.plane.y, light.plane.z), d)) * light_area_012<V>(light);
		float density = dist2 / projected;
		float lambert = dot3(sp.normal, d);
		float3 rb = mk3(0.0f, 0.0f, 0.0f);
		if (lambert > 0.0f) {
			if (visibility) {
				float t_max = -plane_side(sp.position, light.plane) / (d.x * light.plane.x + d.y * light.plane.y + d.z * light.plane.z) - 1e-3f;
				c.rays++;
				visibility = !bvh_any_hit(c.s, sp.position, d, 1.0e-3f, t_max);
			}
			rb = mul3(light.radiance, evaluate_brdf(sp, d));
		}
		if (density > 0.0f) result = add3(result, scale3(rb, lambert / density));
	}
	return scale3(result, 1.0f / (float) S);
}

template <int V, bool DEFER>
__global__ void __launch_bounds__(128) shade_kernel(SceneView s, FrameUniforms f, Variant var, Stripes st, PixelBuffers out) {
	uint32_t x, row, y;
	if (!tile_pixel(f, st, x, row, y)) return;
	const uint32_t pixel = row * f.width + x;
	const uint32_t prim = out.visibility[pixel];
	const uint32_t L = var.light_samples;
	if (prim == 0xFFFFFFFFu || (prim >> 31) != 0u) {
		// background -> (0,0,0), emitter -> (1,1,1) (shading_pass.frag.glsl:686-697); the resolve pass applies exposure
		float v = (prim == 0xFFFFFFFFu) ? 0.0f : 1.0f;
		out.base[pixel] = make_float4(v, v, v, (prim == 0xFFFFFFFFu) ? 1.0f : 0.0f);
		out.origin[pixel] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
		return;
	}
	ShadingPoint sp = reconstruct_shading_point(s, f, prim, primary_ray(f, x, y));
	float fresnel_luminance = dot3(sp.fresnel_0, mk3(0.2126f, 0.7152f, 0.0722f));
	LtcFrame ltc = make_ltc_frame(s, fresnel_luminance, sp.roughness, sp.position, sp.normal, sp.outgoing, f.ltc_constants);
	ShadeContext<V, DEFER> c = { s, f, var, out, pixel, noise_seed(x, y, f.width, f.frame_word), 0u };
	const int N = (int) s.light_count;
	const uint32_t tech = var.polygon_technique;
	const bool psa_like = (tech == TECH_PSA || tech == TECH_PSA_BIASED);
	float3 final_color = mk3(0.0f, 0.0f, 0.0f);
	float3 light_sample = mk3(0.0f, 0.0f, 0.0f);
	uint32_t candidates = 0;
	for (uint32_t j = 0; j != L; ++j) {
		float scale = 0.0f;
		float3 carry = mk3(0.0f, 0.0f, 0.0f);
		float3 color = mk3(0.0f, 0.0f, 0.0f);
		bool visibility = true, have_light = false;
		Light<V> light;
		if (var.light_sampling == 0u) {
			// uniform light pick, shading_pass.frag.glsl:707-721 (index clamped: float(seed) * 2^-32 can round to 1)
			int idx = min((int) (c.next() * (float) N), N - 1);
			light = load_light<V>(s, (uint32_t) idx);
			have_light = true;
			scale = (float) N;
		}
		else {
			// RIS over m = 32 uniformly drawn candidates, shading_pass.frag.glsl:723-736, reservoir.glsl:34-40
			float w_sum = 0.0f, chosen_p_hat = 0.0f;
			int chosen = -1;
			float3 chosen_sample = mk3(0.0f, 0.0f, 0.0f);
			for (int i = 0; i != 32; ++i) {
				int idx = min((int) (c.next() * (float) N), N - 1);
				Light<V> cand = load_light<V>(s, (uint32_t) idx);
				float3 target;
				if (tech == TECH_LTC_CP) target = ltc_target<V>(sp, ltc, cand, var.min_light_vertices);
				else if (psa_like) target = sample_light<V, DEFER>(c, sp, ltc, cand, false, true, 0u);
				else if (tech == TECH_TURK) { bool dummy = false; target = turk_light<V, DEFER>(c, sp, cand, light_sample, false, dummy); }
				else target = mk3(0.0f, 0.0f, 0.0f);
				float p_hat = sqrtf(dot3(target, target));
				float w = p_hat / (1.0f / (float) N);
				float r = c.next();
				w_sum += w;
				if (w > 0.0f && r < (w / w_sum)) { chosen = idx; chosen_p_hat = p_hat; chosen_sample = light_sample; }
			}
			candidates += 32u;
			if (chosen >= 0) {
				light = load_light<V>(s, (uint32_t) chosen);
				have_light = true;
				light_sample = chosen_sample;
				scale = w_sum / (32.0f * chosen_p_hat);
				if ((tech == TECH_LTC_CP || psa_like) && chosen_p_hat == 0.0f) scale = 0.0f;
			}
		}
		if (have_light) {
			if (tech == TECH_LTC_CP) color = sample_light<V, DEFER>(c, sp, ltc, light, true, false, j);
			else if (psa_like) color = sample_light<V, DEFER>(c, sp, ltc, light, false, false, j);
			else if (tech == TECH_TURK) {
				color = turk_light<V, DEFER>(c, sp, light, light_sample, var.light_sampling != 0u, visibility);
				if (var.light_sampling != 0u) scale = scale * (float) (int) visibility;
			}
		}
		if (DEFER) {
			carry = color;
			out.group[(size_t) j * out.pixel_count + pixel] = make_float4(carry.x, carry.y, carry.z, scale);
		}
		else if (have_light) {
			// (color * W) / LIGHT_SAMPLES, resp. (result * N) / LIGHT_SAMPLES * int(visibility)
			float3 r = scale3(color, scale);
			r = mk3(r.x / (float) L, r.y / (float) L, r.z / (float) L);
			if (var.light_sampling == 0u) r = scale3(r, (float) (int) visibility);
			final_color = add3(final_color, r);
		}
	}
	out.base[pixel] = make_float4(final_color.x, final_color.y, final_color.z, 0.0f);
	out.origin[pixel] = make_float4(sp.position.x, sp.position.y, sp.position.z, __uint_as_float(DEFER ? L : 0u));
	// counters: one atomic per warp
	unsigned active = __activemask();
	unsigned rays = __reduce_add_sync(active, c.rays), cands = __reduce_add_sync(active, candidates);
	if ((threadIdx.x & 31u) == (unsigned) (__ffs(active) - 1)) {
		atomicAdd(&out.counters[0], (unsigned long long) __popc(active));
		if (!DEFER) atomicAdd(&out.counters[1], (unsigned long long) rays);
		atomicAdd(&out.counters[3], (unsigned long long) cands);
	}
}

// ---------------------------------------------------------------- (3)
// Shadow rays of the deferred variants: a persistent kernel in which every lane runs its own any-hit traversal as a
// two-track state machine -- per loop iteration at most one inner node (both children's boxes) AND one leaf triangle,
// taken from separate stacks, which is legal because an any-hit query may visit the tree in any order -- and takes the
// next ray from the warp's pool once enough lanes are idle, so that lanes never wait for the longest ray of a fixed
// group of 32; ray records are loaded 64 at a time, coalesced, into a per-warp stage in shared memory, the following
// 64 being prefetched into L2. Rays are numbered slot * pixel_count + pixel; a ray record is {dir, t_max} + {term, state}: state 0 = no
// ray, 1 = requested (visible unless proven otherwise), 2 = occluded (written here); consumed by the accumulation kernel.
#define RL_TRACE_STAGE 64u     // rays staged per warp in shared memory (two per lane, loaded coalesced)
#define RL_TRACE_REFILL 6      // hand out staged rays when this many lanes are idle
#define RL_LEAF_STACK 24

// Reciprocal of a direction component for the fma slab tests t = plane * inv - o * inv. A zero (or flushed) component would
// give inv = inf and then inf - inf = NaN on BOTH planes of the axis, which fminf / fmaxf drop: the axis would count as
// overlapping even when the origin lies outside the slab -- and for o * inv = NaN with a finite other plane a box could be
// missed. Clamping |d| to 1e-30 keeps every product finite (planes and origins are < 1e6) and the test exact enough for
// padded boxes: t = +-1e30 * (plane - o) has the sign the true limit has.
__device__ __forceinline__ float box_reciprocal(float d) { return approx_rcp(fabsf(d) < 1.0e-30f ? copysignf(1.0e-30f, d) : d); }

__device__ __forceinline__ bool slab_fma(float4 lo_hi_a, float2 hi_b, float3 inv, float3 oi, float t_min, float t_max) {
	// box {lo.xyz, hi.x} + {hi.y, hi.z}; t = plane * inv - o * inv (boxes are padded, see bvh.cuh)
	float ax = fmaf(lo_hi_a.x, inv.x, -oi.x), bx = fmaf(lo_hi_a.w, inv.x, -oi.x);
	float ay = fmaf(lo_hi_a.y, inv.y, -oi.y), by = fmaf(hi_b.x, inv.y, -oi.y);
	float az = fmaf(lo_hi_a.z, inv.z, -oi.z), bz = fmaf(hi_b.y, inv.z, -oi.z);
	float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), t_min));
	float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), t_max));
	return t0 <= t1;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

__global__ void __launch_bounds__(128) trace_kernel(SceneView s, PixelBuffers px, uint32_t ray_count, uint32_t tri_vote) {
	// per warp: RL_TRACE_STAGE compacted rays as {origin, t_max}, {direction, ray number}
	__shared__ float4 sm_stage[4][RL_TRACE_STAGE][2];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	float4 (*stage)[2] = sm_stage[warp];
	int node_stack[RL_STACK];
	int leaf_stack[RL_LEAF_STACK];
	uint32_t stage_next = 0, stage_count = 0;   // warp-uniform
	bool exhausted = false, busy = false;
	uint32_t ray = 0, tri_i = 0, tri_end = 0;
	float3 o = mk3(0.0f, 0.0f, 0.0f), d = mk3(0.0f, 0.0f, 1.0f), inv = mk3(0.0f, 0.0f, 0.0f), oi = mk3(0.0f, 0.0f, 0.0f);
	float t_max = 0.0f;
	int node = -1, nsp = 0, lsp = 0;
	const float t_min = 1.0e-3f;
	// the chunk of ray numbers this warp will stage next: claimed one stage ahead so that its records can be pulled into L2
	uint32_t ahead = 0;
	if (lane == 0) ahead = atomicAdd(px.ticket, RL_TRACE_STAGE);
	ahead = __shfl_sync(0xFFFFFFFFu, ahead, 0);
	while (true) {
		const unsigned idle = __ballot_sync(0xFFFFFFFFu, !busy);
		if (__popc(idle) >= RL_TRACE_REFILL || (idle && exhausted)) {
			while (stage_next >= stage_count && !exhausted) {
				// stage the chunk claimed earlier (coalesced: lane l takes rays base + l and base + l + 32) and claim the next one
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
					o = mk3(a0.x, a0.y, a0.z); d = mk3(a1.x, a1.y, a1.z); t_max = a0.w;
					// box tests only: an approximate reciprocal is covered by the padding of the boxes (box_reciprocal keeps it finite)
					inv = mk3(box_reciprocal(d.x), box_reciprocal(d.y), box_reciprocal(d.z));
					oi = mk3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
					node = 0; nsp = 0; lsp = 0; tri_i = tri_end = 0u;
				}
			}
			stage_next = min(stage_next + (uint32_t) __popc(idle), stage_count);
			__syncwarp();   // all reads of the stage are done before a later iteration restages
		}
		// ---- votes: the inner-node track runs for every lane that has a node; the triangle track only once enough lanes
		// have a triangle waiting (or nobody can make progress on nodes), so that its ~85 instructions run on a well
		// filled warp instead of on the ~10 lanes that happen to have reached a leaf in this iteration
		const bool node_ready = busy && node >= 0 && lsp <= RL_LEAF_STACK - 2;
		const bool tri_pending = busy && (tri_i != tri_end || lsp != 0);
		const unsigned node_votes = __ballot_sync(0xFFFFFFFFu, node_ready), tri_votes = __ballot_sync(0xFFFFFFFFu, tri_pending);
		const bool run_tri = (uint32_t) __popc(tri_votes) >= tri_vote || node_votes == 0u;
		// ---- track A: one inner node (paused while the leaf stack could overflow: track B drains it)
		if (node_ready) {
			const BvhNode n = s.nodes[node];
			const bool hl = slab_fma(n.a, make_float2(n.b.x, n.b.y), inv, oi, t_min, t_max);
			const bool hr = slab_fma(make_float4(n.b.z, n.b.w, n.c.x, n.c.y), make_float2(n.c.z, n.c.w), inv, oi, t_min, t_max);
			const int cl = n.d.x, cr = n.d.y;
			const bool il = hl && cl >= 0, ir = hr && cr >= 0;
			if (hl && cl < 0) leaf_stack[lsp++] = cl;
			if (hr && cr < 0) leaf_stack[lsp++] = cr;
			if (il && ir) node_stack[nsp++] = cr;
			if (il) node = cl;
			else if (ir) node = cr;
			else if (nsp) node = node_stack[--nsp];
			else node = -1;
		}
		// ---- track B: one triangle
		if (run_tri && tri_pending) {
			if (tri_i == tri_end) {
				const uint32_t ref = ~(uint32_t) leaf_stack[--lsp];
				tri_i = ref >> 4; tri_end = tri_i + (ref & 15u) + 1u;
			}
			if (tri_any_hit(s.tris[tri_i], o, d, t_min, t_max)) {
				((float*) px.ray_b)[4 * (size_t) ray + 3] = 2.0f;
				busy = false;
			}
			++tri_i;
		}
		if (busy && node < 0 && tri_i == tri_end && lsp == 0) busy = false;   // nothing left on either track: the ray reaches the light
	}
}

// ------------------------------------------------------------ (3) + (4)
// In-place variant: traces the deferred rays of its own pixel (kept for devices / variants that do not use trace_kernel);
// with TRACED = true the rays were decided by trace_kernel and this is kernel (4) alone: MIS sum in the reference's
// order, NaN guard, exposure (shading_pass.frag.glsl:764-769) and the running mean of accum_pass.frag.glsl:45-53.
template <bool TRACED>
__global__ void __launch_bounds__(128) resolve_kernel(SceneView s, FrameUniforms f, Variant var, Stripes st, PixelBuffers out) {
	uint32_t x, row, y;
	if (TRACED && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 3) out.ticket[threadIdx.x] = 0u;   // next frame's ray pool and tile tickets
	if (!tile_pixel(f, st, x, row, y)) return;
	const uint32_t pixel = row * f.width + x;
	float4 o4 = out.origin[pixel], b4 = out.base[pixel];
	float3 origin = mk3(o4.x, o4.y, o4.z);
	float3 color = mk3(b4.x, b4.y, b4.z);
	const uint32_t groups = __float_as_uint(o4.w), S = var.sample_count;
	uint32_t rays = 0;
	for (uint32_t j = 0; j != groups; ++j) {
		float4 g = out.group[(size_t) j * out.pixel_count + pixel];
		float3 sum = mk3(g.x, g.y, g.z);
		for (uint32_t k = 0; k != 2u * S; ++k) {
			size_t slot = (size_t) (j * S * 2u + k) * out.pixel_count + pixel;
			float4 rb = out.ray_b[slot];
			if (rb.w == 0.0f) continue;
			out.ray_b[slot].w = 0.0f;   // slots are consumed: the next frame starts clean
			++rays;
			bool visible;
			if (TRACED) visible = (rb.w == 1.0f);
			else {
				float4 ra = out.ray_a[slot];
				visible = !bvh_any_hit(s, origin, mk3(ra.x, ra.y, ra.z), 1.0e-3f, ra.w);
			}
			if (visible) sum = add3(sum, mk3(rb.x, rb.y, rb.z));
		}
		sum = scale3(sum, 1.0f / (float) S);
		sum = scale3(sum, g.w);
		color = add3(color, mk3(sum.x / (float) groups, sum.y / (float) groups, sum.z / (float) groups));
	}
	// NaN / Inf guard and exposure, shading_pass.frag.glsl:764-769
	if (isnan(color.x) || isnan(color.y) || isnan(color.z) || isinf(color.x) || isinf(color.y) || isinf(color.z))
		color = mk3(1.0f / f.exposure, 0.0f / f.exposure, 0.8f / f.exposure);
	float4 cur = make_float4(color.x * f.exposure, color.y * f.exposure, color.z * f.exposure, 1.0f);
	if (b4.w != 0.0f) cur = make_float4(0.0f, 0.0f, 0.0f, 1.0f);   // background early-out writes (0,0,0,1) without exposure
	// accum_pass.frag.glsl:45-53
	float n = (float) f.accum_num, rcp = 1.0f / (float) (f.accum_num + 1u);
	float4 prev = out.accum[pixel];
	out.accum[pixel] = make_float4(fmaf(fmaf(prev.x, n, cur.x), rcp, 0.0f), fmaf(fmaf(prev.y, n, cur.y), rcp, 0.0f),
	                               fmaf(fmaf(prev.z, n, cur.z), rcp, 0.0f), fmaf(fmaf(prev.w, n, cur.w), rcp, 0.0f));
	unsigned active = __activemask();
	unsigned warp_rays = __reduce_add_sync(active, rays);
	if ((threadIdx.x & 31u) == (unsigned) (__ffs(active) - 1) && warp_rays) atomicAdd(&out.counters[1], (unsigned long long) warp_rays);
}

}  // namespace RL_NS
