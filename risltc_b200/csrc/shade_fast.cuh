// shade_fast.cuh -- the fused RIS + shading kernel specialised for the default estimator
// ("ours", main.c:220-236): light_reservoir with m = 32 candidates whose target function is the
// analytic LTC integral (shading_pass.frag.glsl:723-761, :430-456) over TRIANGLE lights, winner
// shaded by projected-solid-angle + LTC-warped samples combined with optimal-clamped MIS
// (:292-397). Selected by RISLTC_PRECISION_FAST (include/risltc_cuda.h).
//
// What differs from the generic kernel (kernels.cuh) is organisation, not the algorithm:
//  * ONE persistent CTA per SM (up to 24 warps); the light table is staged once per CTA into shared memory as
//    48-byte records {v0 | Le.r, v1 | Le.g, v2 | Le.b} and the 32 random candidates of every pixel are gathered from
//    there with three LDS.128; a warp owns 8x4 pixel tiles;
//  * the candidate needs no plane equation: flipping the shading frame's y row mirrors the polygon,
//    which negates the signed edge sum exactly and calculate_ltc takes abs() (polygon_sampling.glsl:529);
//  * cosine-space vertices come from the shading-space ones through the 5 non-zeros of
//    shading_to_cosine_space instead of a second 4x3 transform of the world-space vertices;
//  * a lane only transforms and classifies its pixel's candidates; the polygons are queued per warp in shared
//    memory and evaluated 32 at a time with all lanes busy (see "work redistribution" below);
//  * r < w / w_sum is tested as r * w_sum < w; the reservoir stays strictly sequential per pixel, so
//    the random stream and the prefix sums keep the reference's order;
//  * the pass is split at the reservoir into ris_ltc3_kernel (G-buffer decode, LTC lookup, 32 candidates,
//    reservoir; FP32-bound) and winner_kernel (the chosen light's PSA + LTC MIS estimator,
//    exactly rounded; ~90 KB of SASS executed once per pixel): fused, the winner's code evicted the candidate
//    loop from the instruction cache (ncu: 7.4 warps stalled on no_instruction per issue, profiles/).
#pragma once
#include "kernels.cuh"

namespace RL_NS {

__device__ __forceinline__ float ff_edge(float3 a, float3 b) {   // integrateEdgeVec on unit vectors
	float x = fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)), y = fabsf(x);
	float num = fmaf(fmaf(0.0145206f, y, 0.4965155f), y, 0.8543985f);
	float den = fmaf(4.1616724f + y, y, 3.4175940f);
	float v = num * approx_rcp(den);
	float alt = fmaf(0.5f, approx_rsqrt(fmaxf(fmaf(-x, x, 1.0f), 1e-7f)), -v);
	float t = (x > 0.0f) ? v : alt;
	return fmaf(a.x, b.y, -a.y * b.x) * t;
}
__device__ __forceinline__ float3 unit3(float3 a) { return scale3(a, approx_rsqrt(fmaf(a.x, a.x, fmaf(a.y, a.y, a.z * a.z)))); }

// Mixed horizon cases of a triangle: 1 vertex above -> triangle, 2 above -> quad (polygon_clipping.glsl:35-225
// restricted to vertex_count == 3). Registers only.
__device__ __noinline__ float ff_clipped_triangle(float3 p0, float3 p1, float3 p2, uint32_t mask) {
	// rotate so that (A, B, C) is cyclic with A the lone vertex above (one above) or C the lone vertex below (two above)
	bool one = (mask == 1u || mask == 2u || mask == 4u);
	uint32_t key = one ? mask : (7u ^ mask);           // the odd vertex out
	float3 A, B, C;                                     // odd vertex -> slot `odd`, successors follow
	if (key == 1u) { A = p0; B = p1; C = p2; }
	else if (key == 2u) { A = p1; B = p2; C = p0; }
	else { A = p2; B = p0; C = p1; }
	// A is the odd one out: crossings on AB and CA
	float3 iab = horizon_crossing(A, B), ica = horizon_crossing(C, A);
	float sum;
	if (one) {
		float3 a = unit3(A), b = unit3(iab), c = unit3(ica);
		sum = ff_edge(a, b) + ff_edge(b, c) + ff_edge(c, a);
	}
	else {
		// A below: polygon B, C, crossing CA, crossing AB
		float3 b = unit3(B), c = unit3(C), d = unit3(ica), e = unit3(iab);
		sum = ff_edge(b, c) + ff_edge(c, d) + ff_edge(d, e) + ff_edge(e, b);
	}
	return fabsf(sum);
}

// ---- work redistribution inside a warp
// A lane owns a pixel, but only ~half of its candidate polygons lie entirely above the horizon (the expensive, common
// case), ~10 % cross it (the very expensive case) and the rest are below it (free). Evaluated in place, the expensive
// code would run with 17 of 32 lanes busy (ncu, profiles/r1_ncu_shade_kernels.txt). Instead every lane only transforms
// and classifies its candidate and PUSHES the polygon into its warp's queue in shared memory -- polygons above the
// horizon from the bottom, horizon-crossing polygons from the top -- and whenever 32 of a kind are waiting the warp
// evaluates them, one per lane, with all lanes busy. Slots are assigned with ballots (no atomics; the counts live in
// warp-uniform registers). Results go to the warp's form-factor table, from which the pixel's lane replays its
// reservoir in the reference's order.
#define RL_CHUNK 8            // candidates per pass (four passes cover m = 32)
// Measured alternatives (round 2, C2, 0.727 ms per frame as is): 16 candidates per pass (half the partially filled drains at
// the end of a pass, 22 warps) 0.722 ms; the candidate loop unrolled by two (loads of the next candidate under the arithmetic of
// this one; 8 bytes of spills) 0.736 ms.
constexpr int kPass1Unroll = 1;
#define RL_QUEUE 128          // polygons per warp queue: <= 31 + 31 left waiting + <= 64 pushed per candidate
#define RL_ITEM_WORDS 12      // p0, p1, p2, mask, destination slot, pad: three 16-byte accesses, conflict-free at this stride
#define RL_WARP_WORDS (2 * RL_CHUNK * 32 + RL_QUEUE * RL_ITEM_WORDS)   // 8 KB per warp
#define RL_FAST_MAX_WARPS 24
#define RL_SMEM_LIMIT (227u * 1024u)

// Shared memory of the CTA (one per SM): [3 N float4 light table][warps x {2 * RL_CHUNK * 32 form factors, RL_QUEUE work items}]
__host__ __device__ inline size_t shade_fast_smem_bytes(uint32_t staged_lights, uint32_t warps) {
	return (size_t) staged_lights * 48 + (size_t) warps * RL_WARP_WORDS * 4;
}
// Largest number of warps whose shared memory fits beside the light table
__host__ inline uint32_t shade_fast_warps(uint32_t staged_lights) {
	uint32_t warps = (RL_SMEM_LIMIT - staged_lights * 48u) / (RL_WARP_WORDS * 4u);
	return warps > RL_FAST_MAX_WARPS ? RL_FAST_MAX_WARPS : warps;
}

// integrateEdgeVec split in two: the rational fit v(|x|) that every edge needs, and the branch for x <= 0 (more than 90
// degrees between two vertices as seen from the shading point) that few polygons need
__device__ __forceinline__ float ff_fit(float y) {
	float num = fmaf(fmaf(0.0145206f, y, 0.4965155f), y, 0.8543985f);
	float den = fmaf(4.1616724f + y, y, 3.4175940f);
	return num * approx_rcp(den);
}
__device__ __forceinline__ float ff_obtuse(float x, float v) { return fmaf(0.5f, approx_rsqrt(fmaxf(fmaf(-x, x, 1.0f), 1e-7f)), -v); }

// |calculate_ltc| of a triangle that lies entirely above the horizon
__device__ __forceinline__ float ff_above(float3 p0, float3 p1, float3 p2) {
	float3 a = unit3(p0), b = unit3(p1), c = unit3(p2);
	float xab = fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)), xbc = fmaf(b.x, c.x, fmaf(b.y, c.y, b.z * c.z)), xca = fmaf(c.x, a.x, fmaf(c.y, a.y, c.z * a.z));
	float tab = ff_fit(fabsf(xab)), tbc = ff_fit(fabsf(xbc)), tca = ff_fit(fabsf(xca));
	if (fminf(xab, fminf(xbc, xca)) <= 0.0f) {
		tab = (xab > 0.0f) ? tab : ff_obtuse(xab, tab);
		tbc = (xbc > 0.0f) ? tbc : ff_obtuse(xbc, tbc);
		tca = (xca > 0.0f) ? tca : ff_obtuse(xca, tca);
	}
	float sum = fmaf(a.x, b.y, -a.y * b.x) * tab;
	sum = fmaf(fmaf(b.x, c.y, -b.y * c.x), tbc, sum);
	sum = fmaf(fmaf(c.x, a.y, -c.y * a.x), tca, sum);
	return fabsf(sum);
}

__device__ __forceinline__ uint32_t horizon_mask(float3 p0, float3 p1, float3 p2) {
	return (p0.z > 0.0f ? 1u : 0u) | (p1.z > 0.0f ? 2u : 0u) | (p2.z > 0.0f ? 4u : 0u);
}

// Push one polygon of this lane: above the horizon -> bottom of the queue, crossing -> top. Polygons below the horizon
// (and those of inactive lanes, whose z are 0) are not queued: their form factor stays at the 0 the table was cleared to.
__device__ __forceinline__ void push_polygon(float* queue, uint32_t& count_above, uint32_t& count_crossing, uint32_t lt_mask,
	float3 p0, float3 p1, float3 p2, uint32_t dest)
{
	const bool above = fminf(fminf(p0.z, p1.z), p2.z) > 0.0f, crossing = !above && fmaxf(fmaxf(p0.z, p1.z), p2.z) > 0.0f;
	const unsigned ballot_above = __ballot_sync(0xFFFFFFFFu, above), ballot_crossing = __ballot_sync(0xFFFFFFFFu, crossing);
	if (above || crossing) {
		const uint32_t rank = __popc((above ? ballot_above : ballot_crossing) & lt_mask);
		const uint32_t slot = above ? count_above + rank : (RL_QUEUE - 1u) - count_crossing - rank;
		float4* item = (float4*) (queue + slot * RL_ITEM_WORDS);
		item[0] = make_float4(p0.x, p0.y, p0.z, p1.x);
		item[1] = make_float4(p1.y, p1.z, p2.x, p2.y);
		item[2] = make_float4(p2.z, __uint_as_float(dest), 0.0f, 0.0f);
	}
	count_above += __popc(ballot_above);
	count_crossing += __popc(ballot_crossing);
}
// Evaluate `n` (<= 32) queued polygons starting at queue slot `first`, one per lane
template <bool ABOVE>
__device__ __forceinline__ void drain(const float* queue, float* ff, uint32_t first, uint32_t n, uint32_t lane) {
	__syncwarp();
	if (lane < n) {
		const float4* item = (const float4*) (queue + (first + lane) * RL_ITEM_WORDS);
		const float4 a = item[0], b = item[1];
		const float2 c = *(const float2*) (item + 2);
		const float3 p0 = mk3(a.x, a.y, a.z), p1 = mk3(a.w, b.x, b.y), p2 = mk3(b.z, b.w, c.x);
		ff[__float_as_uint(c.y)] = ABOVE ? ff_above(p0, p1, p2) : ff_clipped_triangle(p0, p1, p2, horizon_mask(p0, p1, p2));
	}
	__syncwarp();
}

template <bool SMEM, bool TEXTURED>
__global__ void __launch_bounds__(32 * RL_FAST_MAX_WARPS, 1) ris_ltc3_kernel(SceneView s, FrameUniforms f, Stripes st, PixelBuffers out, uint32_t tiles_x, uint32_t tile_count) {
	extern __shared__ float4 sm_base[];
	const int N = (int) s.light_count;
	const uint32_t staged = SMEM ? (uint32_t) N : 0u;
	if (SMEM) for (uint32_t i = threadIdx.x; i < 3u * staged; i += blockDim.x) sm_base[i] = __ldg(&s.lights_tri[i]);
	__syncthreads();
	const float4* table = SMEM ? sm_base : s.lights_tri;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	float* ff = (float*) (sm_base + 3u * staged) + warp * RL_WARP_WORDS;   // [2][RL_CHUNK][32]
	float* queue = ff + 2 * RL_CHUNK * 32;                                   // [RL_QUEUE][RL_ITEM_WORDS]
	const float Nf = (float) N, index_scale = Nf * 2.3283064365386962890625e-10f;
	uint32_t shaded = 0;
	// a warp owns 8x4 pixel tiles, handed out in order through a ticket: tiles differ a lot in cost (background, lights below
	// the horizon, clipped polygons), and when the image is split over several devices a static share is only 2-3 tiles per warp
	while (true) {
		uint32_t tile = 0;
		if (lane == 0) tile = atomicAdd(&out.ticket[1], 1u);
		tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
		if (tile >= tile_count) break;
		const uint32_t x = (tile % tiles_x) * 8u + (lane & 7u);
		const uint32_t row = (tile / tiles_x) * 4u + (lane >> 3);
		const bool inside = x < f.width && row < st.owned_rows;
		const uint32_t y = st.global_row(inside ? row : 0u);
		const uint32_t pixel = row * f.width + x;
		const uint32_t prim = (inside && y < f.height) ? out.visibility[pixel] : 0xFFFFFFFEu;
		const bool active = prim != 0xFFFFFFFFu && (prim >> 31) == 0u;
		if (inside && y < f.height && !active) {
			float v = (prim == 0xFFFFFFFFu) ? 0.0f : 1.0f;
			out.base[pixel] = make_float4(v, v, v, (prim == 0xFFFFFFFFu) ? 1.0f : 0.0f);
			out.origin[pixel] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
		}
		if (!__any_sync(0xFFFFFFFFu, active)) continue;
		ShadingPoint sp;
		LtcFrame ltc;
		ltc.rx = ltc.ry = ltc.rz = ltc.t = mk3(0.0f, 0.0f, 0.0f);
		ltc.s00 = ltc.s02 = ltc.s11 = ltc.s20 = ltc.s22 = 0.0f;
		uint32_t seed = 0;
		if (active) {
			++shaded;
			sp = reconstruct_shading_point<TEXTURED>(s, f, prim, primary_ray(f, x, y));
			float fresnel_luminance = dot3(sp.fresnel_0, mk3(0.2126f, 0.7152f, 0.0722f));
			ltc = make_ltc_frame(s, fresnel_luminance, sp.roughness, sp.position, sp.normal, sp.outgoing, f.ltc_constants);
			seed = noise_seed(x, y, f.width, f.frame_word);
			// the shading point and the fetched LTC table values for winner_kernel (d1 = -s20 is the table value itself)
			const size_t n = out.pixel_count;
			out.shade[pixel] = make_float4(sp.position.x, sp.position.y, sp.position.z, sp.roughness);
			out.shade[n + pixel] = make_float4(sp.normal.x, sp.normal.y, sp.normal.z, ltc.s00);
			out.shade[2 * n + pixel] = make_float4(sp.diffuse_albedo.x, sp.diffuse_albedo.y, sp.diffuse_albedo.z, -ltc.s20);
			out.shade[3 * n + pixel] = make_float4(sp.fresnel_0.x, sp.fresnel_0.y, sp.fresnel_0.z, ltc.s11);
			out.shade[4 * n + pixel] = make_float4(sp.outgoing.x, sp.outgoing.y, sp.outgoing.z, ltc.s02);
			out.shade[5 * n + pixel] = make_float4(ltc.s22, ltc.albedo, 0.0f, 0.0f);
		}
		// ---- RIS over 32 candidates, RL_CHUNK at a time
		float w_sum = 0.0f, chosen_p_hat = 0.0f;
		int chosen = -1;
		for (int chunk = 0; chunk != 32 / RL_CHUNK; ++chunk) {
			const uint32_t chunk_seed = seed;
			uint32_t count_above = 0, count_crossing = 0;   // warp-uniform
			uint32_t dest = lane;
			// clear the form-factor table: polygons below the horizon are never written
			#pragma unroll
			for (int i = 0; i != 2 * RL_CHUNK / 4; ++i) ((float4*) ff)[i * 32 + lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			__syncwarp();
			// pass 1: transform and classify the candidates (target function: LTC integrals, shading_pass.frag.glsl:430-456)
			#pragma unroll kPass1Unroll
			for (int j = 0; j != RL_CHUNK; ++j) {
				// inactive lanes run the same arithmetic on a frame of zeros: their polygons have z = 0 and are never queued
				seed = 1664525u * seed + 1013904223u;
				int idx = min((int) (__uint2float_rn(seed) * index_scale), N - 1);
				seed = 1664525u * seed + 1013904223u;   // the reservoir's draw, consumed in pass 2
				const float4 A = table[3 * idx], B = table[3 * idx + 1], C = table[3 * idx + 2];
				float3 p0, p1, p2;
				p0.x = fmaf(ltc.rx.x, A.x, fmaf(ltc.rx.y, A.y, fmaf(ltc.rx.z, A.z, ltc.t.x)));
				p0.y = fmaf(ltc.ry.x, A.x, fmaf(ltc.ry.y, A.y, fmaf(ltc.ry.z, A.z, ltc.t.y)));
				p0.z = fmaf(ltc.rz.x, A.x, fmaf(ltc.rz.y, A.y, fmaf(ltc.rz.z, A.z, ltc.t.z)));
				p1.x = fmaf(ltc.rx.x, B.x, fmaf(ltc.rx.y, B.y, fmaf(ltc.rx.z, B.z, ltc.t.x)));
				p1.y = fmaf(ltc.ry.x, B.x, fmaf(ltc.ry.y, B.y, fmaf(ltc.ry.z, B.z, ltc.t.y)));
				p1.z = fmaf(ltc.rz.x, B.x, fmaf(ltc.rz.y, B.y, fmaf(ltc.rz.z, B.z, ltc.t.z)));
				p2.x = fmaf(ltc.rx.x, C.x, fmaf(ltc.rx.y, C.y, fmaf(ltc.rx.z, C.z, ltc.t.x)));
				p2.y = fmaf(ltc.ry.x, C.x, fmaf(ltc.ry.y, C.y, fmaf(ltc.ry.z, C.z, ltc.t.y)));
				p2.z = fmaf(ltc.rz.x, C.x, fmaf(ltc.rz.y, C.y, fmaf(ltc.rz.z, C.z, ltc.t.z)));
				const float3 q0 = mk3(fmaf(ltc.s00, p0.x, ltc.s02 * p0.z), ltc.s11 * p0.y, fmaf(ltc.s20, p0.x, ltc.s22 * p0.z));
				const float3 q1 = mk3(fmaf(ltc.s00, p1.x, ltc.s02 * p1.z), ltc.s11 * p1.y, fmaf(ltc.s20, p1.x, ltc.s22 * p1.z));
				const float3 q2 = mk3(fmaf(ltc.s00, p2.x, ltc.s02 * p2.z), ltc.s11 * p2.y, fmaf(ltc.s20, p2.x, ltc.s22 * p2.z));
				push_polygon(queue, count_above, count_crossing, lt_mask, p0, p1, p2, dest);
				push_polygon(queue, count_above, count_crossing, lt_mask, q0, q1, q2, dest + RL_CHUNK * 32u);
				dest += 32u;
				while (count_above >= 32u) { count_above -= 32u; drain<true>(queue, ff, count_above, 32u, lane); }
				while (count_crossing >= 32u) { count_crossing -= 32u; drain<false>(queue, ff, RL_QUEUE - 32u - count_crossing, 32u, lane); }
			}
			if (count_above) drain<true>(queue, ff, 0u, count_above, lane);
			if (count_crossing) drain<false>(queue, ff, RL_QUEUE - count_crossing, count_crossing, lane);
			__syncwarp();
			// pass 2: the reservoir in the reference's order (reservoir.glsl:34-40), replaying the chunk's draws
			if (active) {
				uint32_t replay = chunk_seed;
				#pragma unroll 4
				for (int j = 0; j != RL_CHUNK; ++j) {
					replay = 1664525u * replay + 1013904223u;
					int idx = min((int) (__uint2float_rn(replay) * index_scale), N - 1);
					replay = 1664525u * replay + 1013904223u;
					float r = __uint2float_rn(replay) * 2.3283064365386962890625e-10f;
					const float* rec = (const float*) (table + 3 * idx);
					float fd = ff[j * 32 + lane], fs = ff[(RL_CHUNK + j) * 32 + lane] * ltc.albedo;
					float cr = fmaf(sp.diffuse_albedo.x, fd, fs) * rec[3], cg = fmaf(sp.diffuse_albedo.y, fd, fs) * rec[7], cb = fmaf(sp.diffuse_albedo.z, fd, fs) * rec[11];
					float p_hat = approx_sqrt(fmaf(cr, cr, fmaf(cg, cg, cb * cb)));
					float w = p_hat * Nf;
					w_sum += w;
					if (w > 0.0f && r * w_sum < w) { chosen = idx; chosen_p_hat = p_hat; }
				}
			}
			__syncwarp();   // pass 1 of the next chunk overwrites the form factors
		}
		// ---- hand the winner to winner_kernel: light index, W = w_sum / (m p_hat) (shading_pass.frag.glsl:747-752), RNG state
		if (active) {
			float scale = (chosen < 0 || chosen_p_hat == 0.0f) ? 0.0f : w_sum / (32.0f * chosen_p_hat);
			out.pick[pixel] = make_uint4((uint32_t) chosen, __float_as_uint(scale), seed, 0u);
		}
	}
	// counters: one atomic per warp at the end of the CTA's life
	unsigned total = __reduce_add_sync(0xFFFFFFFFu, shaded);
	if (lane == 0 && total) {
		atomicAdd(&out.counters[0], (unsigned long long) total);
		atomicAdd(&out.counters[3], 32ull * total);
	}
}

// ---------------------------------------------------------------------------------------------------------------------
// The same kernel for QUAD lights (MAX_POLYGONAL_LIGHT_VERTEX_COUNT = 4, the shape polygonal_light.c creates by default and the
// V = 4 variants of the comparison matrix use). Triangles among them are fine: write_lights repeats a light's first vertex up
// to the maximal count (main.c:483-487), and a quad {v0, v1, v2, v0} has the triangle's form factor (the extra edges have
// zero length: their cross products are exactly 0) and the triangle's clipped polygon. What changes is the size of
// things: table records {v0 | Le.r, v1 | Le.g, v2 | Le.b, v3 | 0}, a four-edge form factor, and the horizon clip of a quad
// (up to five vertices, rare: a small walk through local arrays). Strides are chosen ODD in 16-byte units, so that
// consecutive slots / random records spread over all eight 16-byte bank groups like the triangle kernel's 48-byte records
// do: table records are 80 bytes apart in shared memory (the fifth vector is padding), queue items hold the twelve
// coordinates in 48 bytes and their destination slots live in an array of their own. (A first version with 64-byte
// records and items ran into four-way bank conflicts on every gather, push and drain: 2.1 instead of 0.9 ms per C2 frame.)
#define RL_ITEM_WORDS4 12     // p0..p3: three 16-byte accesses
#define RL_TABLE_STRIDE4 5u   // float4 per staged quad light
#define RL_WARP_WORDS4 (2 * RL_CHUNK * 32 + RL_QUEUE * RL_ITEM_WORDS4 + RL_QUEUE)   // 8.5 KB per warp
__host__ __device__ inline size_t shade_fast4_smem_bytes(uint32_t staged_lights, uint32_t warps) {
	return (size_t) staged_lights * 16 * RL_TABLE_STRIDE4 + (size_t) warps * RL_WARP_WORDS4 * 4;
}
__host__ inline uint32_t shade_fast4_warps(uint32_t staged_lights) {
	uint32_t warps = (RL_SMEM_LIMIT - staged_lights * 16u * RL_TABLE_STRIDE4) / (RL_WARP_WORDS4 * 4u);
	return warps > RL_FAST_MAX_WARPS ? RL_FAST_MAX_WARPS : warps;
}

// |calculate_ltc| of a quad that lies entirely above the horizon
__device__ __forceinline__ float ff_above4(float3 p0, float3 p1, float3 p2, float3 p3) {
	float3 a = unit3(p0), b = unit3(p1), c = unit3(p2), d = unit3(p3);
	float xab = fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)), xbc = fmaf(b.x, c.x, fmaf(b.y, c.y, b.z * c.z));
	float xcd = fmaf(c.x, d.x, fmaf(c.y, d.y, c.z * d.z)), xda = fmaf(d.x, a.x, fmaf(d.y, a.y, d.z * a.z));
	float tab = ff_fit(fabsf(xab)), tbc = ff_fit(fabsf(xbc)), tcd = ff_fit(fabsf(xcd)), tda = ff_fit(fabsf(xda));
	if (fminf(fminf(xab, xbc), fminf(xcd, xda)) <= 0.0f) {
		tab = (xab > 0.0f) ? tab : ff_obtuse(xab, tab);
		tbc = (xbc > 0.0f) ? tbc : ff_obtuse(xbc, tbc);
		tcd = (xcd > 0.0f) ? tcd : ff_obtuse(xcd, tcd);
		tda = (xda > 0.0f) ? tda : ff_obtuse(xda, tda);
	}
	float sum = fmaf(a.x, b.y, -a.y * b.x) * tab;
	sum = fmaf(fmaf(b.x, c.y, -b.y * c.x), tbc, sum);
	sum = fmaf(fmaf(c.x, d.y, -c.y * d.x), tcd, sum);
	sum = fmaf(fmaf(d.x, a.y, -d.y * a.x), tda, sum);
	return fabsf(sum);
}
// A quad that crosses the horizon: the clipped polygon (polygon_clipping.glsl:35-225 for vertex_count == 4; the start of
// the walk does not matter for the absolute value of the closed edge sum), three to five vertices
__device__ __noinline__ float ff_clipped_quad(float3 p0, float3 p1, float3 p2, float3 p3) {
	const float3 v[4] = { p0, p1, p2, p3 };
	float3 out[6];
	int n = 0;
	#pragma unroll
	for (int i = 0; i != 4; ++i) {
		const int j = (i + 1) & 3;
		const bool a = v[i].z > 0.0f, b = v[j].z > 0.0f;
		if (a) out[n++] = unit3(v[i]);
		if (a != b) out[n++] = unit3(horizon_crossing(v[i], v[j]));
	}
	if (n < 3) return 0.0f;
	float sum = 0.0f;
	for (int k = 0; k != n; ++k) sum += ff_edge(out[k], out[(k + 1 == n) ? 0 : k + 1]);
	return fabsf(sum);
}
__device__ __forceinline__ void push_quad(float* queue, uint32_t* queue_dest, uint32_t& count_above, uint32_t& count_crossing, uint32_t lt_mask,
	float3 p0, float3 p1, float3 p2, float3 p3, uint32_t dest)
{
	const float zmin = fminf(fminf(p0.z, p1.z), fminf(p2.z, p3.z)), zmax = fmaxf(fmaxf(p0.z, p1.z), fmaxf(p2.z, p3.z));
	const bool above = zmin > 0.0f, crossing = !above && zmax > 0.0f;
	const unsigned ballot_above = __ballot_sync(0xFFFFFFFFu, above), ballot_crossing = __ballot_sync(0xFFFFFFFFu, crossing);
	if (above || crossing) {
		const uint32_t rank = __popc((above ? ballot_above : ballot_crossing) & lt_mask);
		const uint32_t slot = above ? count_above + rank : (RL_QUEUE - 1u) - count_crossing - rank;
		float4* item = (float4*) (queue + slot * RL_ITEM_WORDS4);
		item[0] = make_float4(p0.x, p0.y, p0.z, p1.x);
		item[1] = make_float4(p1.y, p1.z, p2.x, p2.y);
		item[2] = make_float4(p2.z, p3.x, p3.y, p3.z);
		queue_dest[slot] = dest;
	}
	count_above += __popc(ballot_above);
	count_crossing += __popc(ballot_crossing);
}
template <bool ABOVE>
__device__ __forceinline__ void drain4(const float* queue, const uint32_t* queue_dest, float* ff, uint32_t first, uint32_t n, uint32_t lane) {
	__syncwarp();
	if (lane < n) {
		const float4* item = (const float4*) (queue + (first + lane) * RL_ITEM_WORDS4);
		const float4 a = item[0], b = item[1], c = item[2];
		const uint32_t dest = queue_dest[first + lane];
		const float3 p0 = mk3(a.x, a.y, a.z), p1 = mk3(a.w, b.x, b.y), p2 = mk3(b.z, b.w, c.x), p3 = mk3(c.y, c.z, c.w);
		ff[dest] = ABOVE ? ff_above4(p0, p1, p2, p3) : ff_clipped_quad(p0, p1, p2, p3);
	}
	__syncwarp();
}

template <bool SMEM, bool TEXTURED>
__global__ void __launch_bounds__(32 * RL_FAST_MAX_WARPS, 1) ris_ltc4_kernel(SceneView s, FrameUniforms f, Stripes st, PixelBuffers out, uint32_t tiles_x, uint32_t tile_count) {
	extern __shared__ float4 sm_base[];
	const int N = (int) s.light_count;
	const uint32_t staged = SMEM ? (uint32_t) N : 0u;
	if (SMEM) for (uint32_t i = threadIdx.x; i < 4u * staged; i += blockDim.x) sm_base[(i >> 2) * RL_TABLE_STRIDE4 + (i & 3u)] = __ldg(&s.lights_tri[i]);
	__syncthreads();
	const float4* table = SMEM ? sm_base : s.lights_tri;
	const uint32_t table_stride = SMEM ? RL_TABLE_STRIDE4 : 4u;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	float* ff = (float*) (sm_base + RL_TABLE_STRIDE4 * staged) + warp * RL_WARP_WORDS4;   // [2][RL_CHUNK][32]
	float* queue = ff + 2 * RL_CHUNK * 32;                                                 // [RL_QUEUE][RL_ITEM_WORDS4]
	uint32_t* queue_dest = (uint32_t*) (queue + RL_QUEUE * RL_ITEM_WORDS4);                // [RL_QUEUE]
	const float Nf = (float) N, index_scale = Nf * 2.3283064365386962890625e-10f;
	uint32_t shaded = 0;
	while (true) {
		uint32_t tile = 0;
		if (lane == 0) tile = atomicAdd(&out.ticket[1], 1u);
		tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
		if (tile >= tile_count) break;
		const uint32_t x = (tile % tiles_x) * 8u + (lane & 7u);
		const uint32_t row = (tile / tiles_x) * 4u + (lane >> 3);
		const bool inside = x < f.width && row < st.owned_rows;
		const uint32_t y = st.global_row(inside ? row : 0u);
		const uint32_t pixel = row * f.width + x;
		const uint32_t prim = (inside && y < f.height) ? out.visibility[pixel] : 0xFFFFFFFEu;
		const bool active = prim != 0xFFFFFFFFu && (prim >> 31) == 0u;
		if (inside && y < f.height && !active) {
			float v = (prim == 0xFFFFFFFFu) ? 0.0f : 1.0f;
			out.base[pixel] = make_float4(v, v, v, (prim == 0xFFFFFFFFu) ? 1.0f : 0.0f);
			out.origin[pixel] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
		}
		if (!__any_sync(0xFFFFFFFFu, active)) continue;
		ShadingPoint sp;
		LtcFrame ltc;
		ltc.rx = ltc.ry = ltc.rz = ltc.t = mk3(0.0f, 0.0f, 0.0f);
		ltc.s00 = ltc.s02 = ltc.s11 = ltc.s20 = ltc.s22 = 0.0f;
		uint32_t seed = 0;
		if (active) {
			++shaded;
			sp = reconstruct_shading_point<TEXTURED>(s, f, prim, primary_ray(f, x, y));
			float fresnel_luminance = dot3(sp.fresnel_0, mk3(0.2126f, 0.7152f, 0.0722f));
			ltc = make_ltc_frame(s, fresnel_luminance, sp.roughness, sp.position, sp.normal, sp.outgoing, f.ltc_constants);
			seed = noise_seed(x, y, f.width, f.frame_word);
			const size_t n = out.pixel_count;
			out.shade[pixel] = make_float4(sp.position.x, sp.position.y, sp.position.z, sp.roughness);
			out.shade[n + pixel] = make_float4(sp.normal.x, sp.normal.y, sp.normal.z, ltc.s00);
			out.shade[2 * n + pixel] = make_float4(sp.diffuse_albedo.x, sp.diffuse_albedo.y, sp.diffuse_albedo.z, -ltc.s20);
			out.shade[3 * n + pixel] = make_float4(sp.fresnel_0.x, sp.fresnel_0.y, sp.fresnel_0.z, ltc.s11);
			out.shade[4 * n + pixel] = make_float4(sp.outgoing.x, sp.outgoing.y, sp.outgoing.z, ltc.s02);
			out.shade[5 * n + pixel] = make_float4(ltc.s22, ltc.albedo, 0.0f, 0.0f);
		}
		float w_sum = 0.0f, chosen_p_hat = 0.0f;
		int chosen = -1;
		for (int chunk = 0; chunk != 32 / RL_CHUNK; ++chunk) {
			const uint32_t chunk_seed = seed;
			uint32_t count_above = 0, count_crossing = 0;   // warp-uniform
			uint32_t dest = lane;
			#pragma unroll
			for (int i = 0; i != 2 * RL_CHUNK / 4; ++i) ((float4*) ff)[i * 32 + lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			__syncwarp();
			#pragma unroll 1
			for (int j = 0; j != RL_CHUNK; ++j) {
				seed = 1664525u * seed + 1013904223u;
				int idx = min((int) (__uint2float_rn(seed) * index_scale), N - 1);
				seed = 1664525u * seed + 1013904223u;   // the reservoir's draw, consumed in pass 2
				float3 p[4], q[4];
				#pragma unroll
				for (int k = 0; k != 4; ++k) {
					const float4 A = table[table_stride * idx + k];
					p[k].x = fmaf(ltc.rx.x, A.x, fmaf(ltc.rx.y, A.y, fmaf(ltc.rx.z, A.z, ltc.t.x)));
					p[k].y = fmaf(ltc.ry.x, A.x, fmaf(ltc.ry.y, A.y, fmaf(ltc.ry.z, A.z, ltc.t.y)));
					p[k].z = fmaf(ltc.rz.x, A.x, fmaf(ltc.rz.y, A.y, fmaf(ltc.rz.z, A.z, ltc.t.z)));
					q[k] = mk3(fmaf(ltc.s00, p[k].x, ltc.s02 * p[k].z), ltc.s11 * p[k].y, fmaf(ltc.s20, p[k].x, ltc.s22 * p[k].z));
				}
				push_quad(queue, queue_dest, count_above, count_crossing, lt_mask, p[0], p[1], p[2], p[3], dest);
				push_quad(queue, queue_dest, count_above, count_crossing, lt_mask, q[0], q[1], q[2], q[3], dest + RL_CHUNK * 32u);
				dest += 32u;
				while (count_above >= 32u) { count_above -= 32u; drain4<true>(queue, queue_dest, ff, count_above, 32u, lane); }
				while (count_crossing >= 32u) { count_crossing -= 32u; drain4<false>(queue, queue_dest, ff, RL_QUEUE - 32u - count_crossing, 32u, lane); }
			}
			if (count_above) drain4<true>(queue, queue_dest, ff, 0u, count_above, lane);
			if (count_crossing) drain4<false>(queue, queue_dest, ff, RL_QUEUE - count_crossing, count_crossing, lane);
			__syncwarp();
			if (active) {
				uint32_t replay = chunk_seed;
				#pragma unroll 4
				for (int j = 0; j != RL_CHUNK; ++j) {
					replay = 1664525u * replay + 1013904223u;
					int idx = min((int) (__uint2float_rn(replay) * index_scale), N - 1);
					replay = 1664525u * replay + 1013904223u;
					float r = __uint2float_rn(replay) * 2.3283064365386962890625e-10f;
					const float* rec = (const float*) (table + table_stride * idx);
					float fd = ff[j * 32 + lane], fs = ff[(RL_CHUNK + j) * 32 + lane] * ltc.albedo;
					float cr = fmaf(sp.diffuse_albedo.x, fd, fs) * rec[3], cg = fmaf(sp.diffuse_albedo.y, fd, fs) * rec[7], cb = fmaf(sp.diffuse_albedo.z, fd, fs) * rec[11];
					float p_hat = approx_sqrt(fmaf(cr, cr, fmaf(cg, cg, cb * cb)));
					float w = p_hat * Nf;
					w_sum += w;
					if (w > 0.0f && r * w_sum < w) { chosen = idx; chosen_p_hat = p_hat; }
				}
			}
			__syncwarp();
		}
		if (active) {
			float scale = (chosen < 0 || chosen_p_hat == 0.0f) ? 0.0f : w_sum / (32.0f * chosen_p_hat);
			out.pick[pixel] = make_uint4((uint32_t) chosen, __float_as_uint(scale), seed, 0u);
		}
	}
	unsigned total = __reduce_add_sync(0xFFFFFFFFu, shaded);
	if (lane == 0 && total) {
		atomicAdd(&out.counters[0], (unsigned long long) total);
		atomicAdd(&out.counters[3], 32ull * total);
	}
}

// ---------------------------------------------------------------------------------------------------------------------
// light_uniform + sample_polygon_ltc_cp (shading_pass.frag.glsl:707-721 with :292-397): no reservoir -- the light is one
// uniform draw and the estimate is N times the same estimator that the RIS winner gets. This kernel is kernel 2a without
// its candidates: G-buffer decode, LTC lookup, the draw, and the hand-over records (pick = {light, W = N, RNG state},
// shading point) that winner_kernel, the shadow-ray kernel and resolve_kernel continue from. One thread per pixel.
template <bool TEXTURED>
__global__ void __launch_bounds__(128) pick_uniform_kernel(SceneView s, FrameUniforms f, Stripes st, PixelBuffers out) {
	uint32_t x, row, y;
	if (!tile_pixel(f, st, x, row, y)) return;
	const uint32_t pixel = row * f.width + x;
	const uint32_t prim = out.visibility[pixel];
	const bool active = prim != 0xFFFFFFFFu && (prim >> 31) == 0u;
	if (!active) {
		const float v = (prim == 0xFFFFFFFFu) ? 0.0f : 1.0f;
		out.base[pixel] = make_float4(v, v, v, (prim == 0xFFFFFFFFu) ? 1.0f : 0.0f);
		out.origin[pixel] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
		return;
	}
	const ShadingPoint sp = reconstruct_shading_point<TEXTURED>(s, f, prim, primary_ray(f, x, y));
	const float fresnel_luminance = dot3(sp.fresnel_0, mk3(0.2126f, 0.7152f, 0.0722f));
	const LtcFrame ltc = make_ltc_frame(s, fresnel_luminance, sp.roughness, sp.position, sp.normal, sp.outgoing, f.ltc_constants);
	uint32_t seed = noise_seed(x, y, f.width, f.frame_word);
	const size_t n = out.pixel_count;
	out.shade[pixel] = make_float4(sp.position.x, sp.position.y, sp.position.z, sp.roughness);
	out.shade[n + pixel] = make_float4(sp.normal.x, sp.normal.y, sp.normal.z, ltc.s00);
	out.shade[2 * n + pixel] = make_float4(sp.diffuse_albedo.x, sp.diffuse_albedo.y, sp.diffuse_albedo.z, -ltc.s20);
	out.shade[3 * n + pixel] = make_float4(sp.fresnel_0.x, sp.fresnel_0.y, sp.fresnel_0.z, ltc.s11);
	out.shade[4 * n + pixel] = make_float4(sp.outgoing.x, sp.outgoing.y, sp.outgoing.z, ltc.s02);
	out.shade[5 * n + pixel] = make_float4(ltc.s22, ltc.albedo, 0.0f, 0.0f);
	// light_idx = int(get_noise_1() * N), clamped (the draw can round to 1.0)
	const int N = (int) s.light_count;
	seed = 1664525u * seed + 1013904223u;
	const int idx = min((int) (__uint2float_rn(seed) * ((float) N * 2.3283064365386962890625e-10f)), N - 1);
	out.pick[pixel] = make_uint4((uint32_t) idx, __float_as_uint((float) N), seed, 0u);
	const unsigned active_mask = __activemask();
	const unsigned total = (unsigned) __popc(active_mask);
	if ((threadIdx.x & 31u) == (unsigned) (__ffs(active_mask) - 1)) atomicAdd(&out.counters[0], (unsigned long long) total);
}

// clip_to_horizon<4> for a triangle (vertex count == MIN_POLYGON_VERTEX_COUNT_BEFORE_CLIPPING == 3) without the dynamically
// indexed walk: the same vertices in the same order (polygon_clipping.glsl:56-75: row 0 of the rotation table), every slot
// addressed statically so that the polygon stays in registers. horizon_crossing(v_i, v_i+1) per crossed edge, as there.
__device__ __forceinline__ uint32_t clip_triangle_to_horizon(float3 (&v)[4]) {
	const uint32_t mask = (v[0].z > 0.0f ? 1u : 0u) | (v[1].z > 0.0f ? 2u : 0u) | (v[2].z > 0.0f ? 4u : 0u);
	if (mask == 0u) return 0u;
	if (mask == 7u) return 3u;
	const float3 a = v[0], b = v[1], c = v[2];
	switch (mask) {
	case 1u: v[1] = horizon_crossing(a, b); v[2] = horizon_crossing(c, a); return 3u;            // v0, x01, x20
	case 2u: v[0] = horizon_crossing(a, b); v[2] = horizon_crossing(b, c); return 3u;            // x01, v1, x12
	case 3u: v[2] = horizon_crossing(b, c); v[3] = horizon_crossing(c, a); return 4u;            // v0, v1, x12, x20
	case 4u: v[0] = horizon_crossing(c, a); v[1] = horizon_crossing(b, c); return 3u;            // x20, x12, v2 (walk rotated by 2)
	case 5u: v[1] = horizon_crossing(a, b); v[2] = horizon_crossing(b, c); v[3] = c; return 4u;  // v0, x01, x12, v2
	default: v[0] = horizon_crossing(a, b); v[3] = horizon_crossing(c, a); return 4u;            // 6: x01, v1, v2, x20
	}
}

// The winner's estimator (evaluate_polygonal_light_shading_peters, shading_pass.frag.glsl:292-397) for the light that
// ris_ltc3_kernel chose; shadow rays are deferred to the trace kernel. Every operation here is rounded as in the oracle.
//
// Organisation. The estimator is ~80 KB of straight-line SASS, far beyond the 32 KB instruction cache of an SM: a CTA walks
// through it in PHASES separated by __syncthreads(), so that all its warps execute the same < 32 KB of code at the same time
// and every instruction line is fetched once per CTA and phase instead of once per warp (ncu, first version without
// phases: no_instruction 3.8 warps per issue, issue-active 47 %).
// The two techniques (diffuse: the polygon in shading space, specular: in cosine space) run through the SAME code, a loop
// of two iterations with {transform + clip + PSA preparation | barrier | sample | barrier} as its body. Only one prepared
// polygon exists at a time, and it crosses its barrier in SHARED memory (24 words per thread, one column per thread:
// bank = lane), all slots addressed statically; the per-sample body only needs the two solid angles. The round-1 kernel
// kept both polygons, both clipped vertex lists and the light in per-thread local memory behind out-of-line calls
// (496-byte frames x 768 threads = 380 KB per SM, more than L1 holds): ncu showed 48 M + 47 M local sectors per launch and
// 477 MB of DRAM writes against 232 MB algorithmic (profiles/r1_ncu_final_kernels.txt).
// The draws keep the reference's order (diffuse pair, then specular pair only if its solid angle is positive).
// V = MAX_POLYGONAL_LIGHT_VERTEX_COUNT (3: triangles, the bench configurations; 4: quads, the reference's default light
// shape, or a mix of triangles and quads -- the clipped polygon then has up to P = V + 1 vertices and the clip is the generic
// rotation-table walk of shading.cuh, which takes the light's own vertex count and MIN_POLYGONAL_LIGHT_VERTEX_COUNT).
template <int RL_WIN_THREADS, int RL_WIN_RESIDENT_THREADS, int V = 3>
__global__ void __launch_bounds__(RL_WIN_THREADS, RL_WIN_RESIDENT_THREADS / RL_WIN_THREADS) winner_kernel(SceneView s, FrameUniforms f, Stripes st, PixelBuffers out, uint32_t tiles_x, uint32_t tile_count, uint32_t min_light_vertices) {
	constexpr int P = V + 1;
	constexpr int RL_POLY_WORDS = 5 * P + 4;   // vc, v[P], e[P], inner0, sector[P], total
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, warps = RL_WIN_THREADS / 32;
	const Variant var = { 1u, TECH_LTC_CP, MIS_OPTIMAL_CLAMPED, 1u, 1u, 0u, (V == 3) ? 3u : min_light_vertices, (uint32_t) V };
	// a warp owns 8x4 pixel tiles; the CTA takes `warps` neighbouring tiles per round through a ticket (barriers inside)
	__shared__ uint32_t sm_base;
	__shared__ float sm_poly[RL_POLY_WORDS][RL_WIN_THREADS];
	float* const poly = &sm_poly[0][threadIdx.x];
	#define RL_POLY(k) poly[(k) * RL_WIN_THREADS]
	while (true) {
		__syncthreads();
		if (threadIdx.x == 0) sm_base = atomicAdd(&out.ticket[2], warps);
		__syncthreads();
		const uint32_t base = sm_base;
		if (base >= tile_count) break;
		const uint32_t tile = base + warp;
		const uint32_t x = (tile % tiles_x) * 8u + (lane & 7u);
		const uint32_t row = (tile / tiles_x) * 4u + (lane >> 3);
		const bool inside = tile < tile_count && x < f.width && row < st.owned_rows;
		const uint32_t y = st.global_row(inside ? row : 0u);
		const uint32_t pixel = row * f.width + x;
		const uint32_t prim = (inside && y < f.height) ? out.visibility[pixel] : 0xFFFFFFFFu;
		const bool shaded = prim != 0xFFFFFFFFu && (prim >> 31) == 0u;   // base / origin of the other pixels were written by ris_ltc3_kernel
		uint4 pick = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
		if (shaded) pick = out.pick[pixel];
		bool live = shaded && (int) pick.x >= 0;
		// ---- (A) shading point, LTC frame, side of the light's plane
		ShadingPoint sp;
		LtcFrame ltc;
		TechniqueTerms t;
		t.flip = false; t.specular_total = 0.0f;
		const float4* light_record = s.lights + (size_t) (live ? pick.x : 0u) * s.light_stride4;
		if (shaded) {
			// the shading point and the LTC table values as ris_ltc3_kernel computed them (G-buffer decode, texture fetches, acos
			// and the table lookup are not repeated here); everything derived from them is recomputed with the same operations
			const size_t n = out.pixel_count;
			const float4 q0 = out.shade[pixel], q1 = out.shade[n + pixel], q2 = out.shade[2 * n + pixel], q3 = out.shade[3 * n + pixel];
			const float4 q4 = out.shade[4 * n + pixel], q5 = out.shade[5 * n + pixel];
			sp.position = mk3(q0.x, q0.y, q0.z); sp.roughness = q0.w;
			sp.normal = mk3(q1.x, q1.y, q1.z);
			sp.diffuse_albedo = mk3(q2.x, q2.y, q2.z);
			sp.fresnel_0 = mk3(q3.x, q3.y, q3.z);
			sp.outgoing = mk3(q4.x, q4.y, q4.z);
			sp.lambert_outgoing = dot3(sp.normal, sp.outgoing);
			if (live) {
				const float d[6] = { q1.w, q2.w, q3.w, q4.w, q5.x, q5.y };
				ltc = ltc_frame_from_fetch(d, sp.position, sp.normal, sp.outgoing);
				t.flip = plane_side(sp.position, __ldg(light_record + 1)) < 0.0f;
			}
		}
		uint32_t seed = pick.z;
		float3 dir0 = mk3(0.0f, 0.0f, 0.0f), dir1 = dir0;
		float total_d = 0.0f, total_s = 0.0f;
		bool polygon_d = false;   // the diffuse polygon survived the clip (prepare_techniques returns early otherwise, :307-310)
		// ---- (B), (C) per technique: prepare_techniques + the technique's sample, shading_pass.frag.glsl:296-372
		#pragma unroll 1
		for (int tech = 0; tech != 2; ++tech) {
			__syncthreads();
			bool prepared = false;
			if (live && (tech == 0 || polygon_d)) {
				float3 pv[P];
				#pragma unroll
				for (int i = 0; i != V; ++i) {
					const float4 w = __ldg(light_record + 3 + i);
					pv[i] = tech ? to_cosine_space(ltc, mk3(w.x, w.y, w.z), t.flip) : to_shading_space(ltc, mk3(w.x, w.y, w.z), t.flip);
				}
				pv[V] = mk3(0.0f, 0.0f, 0.0f);
				uint32_t vc;
				if (V == 3) vc = clip_triangle_to_horizon(reinterpret_cast<float3 (&)[4]>(pv));
				else vc = clip_to_horizon<P>(__float_as_uint(__ldg(light_record + 2).x), pv, var.min_light_vertices);   // the light's own vertex count (main.c:478)
				if (tech == 0) { polygon_d = vc != 0u; live = polygon_d; }
				if (vc != 0u) {
					PsaPolygon<P> p;
					#pragma unroll
					for (int i = 0; i != P; ++i) { p.v[i] = mk2(0.0f, 0.0f); p.e[i] = mk2(0.0f, 0.0f); p.sector[i] = 0.0f; }
					psa_prepare<P, false>(p, vc, pv);
					if (tech == 0) { total_d = p.total; live = total_d != 0.0f; }
					else total_s = p.total;
					prepared = live && (tech == 0 || total_s > 0.0f);
					if (prepared) {
						RL_POLY(0) = __uint_as_float(p.vc);
						#pragma unroll
						for (int i = 0; i != P; ++i) {
							RL_POLY(1 + 2 * i) = p.v[i].x; RL_POLY(2 + 2 * i) = p.v[i].y;
							RL_POLY(1 + 2 * P + 2 * i) = p.e[i].x; RL_POLY(2 + 2 * P + 2 * i) = p.e[i].y;
							RL_POLY(3 + 4 * P + i) = p.sector[i];
						}
						RL_POLY(1 + 4 * P) = p.inner0.x; RL_POLY(2 + 4 * P) = p.inner0.y; RL_POLY(3 + 5 * P) = p.total;
					}
				}
			}
			__syncthreads();
			if (prepared) {
				PsaPolygon<P> p;
				p.vc = __float_as_uint(RL_POLY(0));
				#pragma unroll
				for (int i = 0; i != P; ++i) {
					p.v[i] = mk2(RL_POLY(1 + 2 * i), RL_POLY(2 + 2 * i));
					p.e[i] = mk2(RL_POLY(1 + 2 * P + 2 * i), RL_POLY(2 + 2 * P + 2 * i));
					p.sector[i] = RL_POLY(3 + 4 * P + i);
				}
				p.inner0 = mk2(RL_POLY(1 + 4 * P), RL_POLY(2 + 4 * P)); p.total = RL_POLY(3 + 5 * P);
				const float u0 = noise_next(seed), u1 = noise_next(seed);
				const float3 d = psa_sample<P, false, false>(p, u0, u1);
				if (tech == 0) dir0 = d;
				else dir1 = cosine_to_shading_dir(ltc, d);
			}
		}
		__syncthreads();
		// ---- (D): densities, BRDF, MIS (shading_pass.frag.glsl:373-394); the rays are recorded for the trace kernel
		float3 carry = mk3(0.0f, 0.0f, 0.0f);
		if (live) {
			Light<V> light;
			const float4 la = __ldg(light_record), lb = __ldg(light_record + 1);
			light.radiance = mk3(la.x, la.y, la.z); light.plane = lb; light.count = (V == 3) ? 3u : __float_as_uint(__ldg(light_record + 2).x);
			technique_weights(t, sp, ltc.albedo, total_d, total_s, light.radiance, false);
			const int techniques = (total_s > 0.0f) ? 2 : 1;
			for (int j = 0; j != techniques; ++j) {
				RayRequest ray; bool side_visible; float3 if_occluded;
				ray.dir = mk3(0.0f, 0.0f, 1.0f); ray.t_max = -1.0f; ray.if_visible = mk3(0.0f, 0.0f, 0.0f);
				if (!technique_sample<V>(t, sp, ltc, light, var, j, j ? dir1 : dir0, f.mis_visibility_estimate, true, ray, side_visible, if_occluded)) continue;
				if (!side_visible) { carry = add3(carry, if_occluded); continue; }
				out.ray_a[(size_t) j * out.pixel_count + pixel] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, ray.t_max);
				out.ray_b[(size_t) j * out.pixel_count + pixel] = make_float4(ray.if_visible.x, ray.if_visible.y, ray.if_visible.z, 1.0f);
			}
		}
		if (shaded) {
			out.group[pixel] = make_float4(carry.x, carry.y, carry.z, __uint_as_float(pick.y));
			out.base[pixel] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			out.origin[pixel] = make_float4(sp.position.x, sp.position.y, sp.position.z, __uint_as_float(1u));
		}
	}
	#undef RL_POLY
}

}  // namespace RL_NS
