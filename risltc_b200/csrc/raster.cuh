// raster.cuh -- kernel (1), primary visibility (visibility_pass.vert/frag.glsl, raster state main.c:715-721,751-756),
// second generation: TRIANGLE-parallel instead of ray-parallel.
//
// The visibility buffer is defined (oracle: closest_front_hit) as the lexicographic minimum of (t, triangle index) over
// all front-facing triangles that the pixel-centre ray hits inside the depth range -- a definition that does not mention
// a tree. gbuffer_kernel (kernels.cuh) evaluates it by walking the BVH per pixel (~42 node visits per ray, issue-bound);
// here every triangle finds ITS pixels, as the reference's rasteriser does, and competes for them with a 64-bit
// atomicMin on the key {bits(t) : index}:
//   * the hit conditions u >= 0, v >= 0, u + v <= 1 (times det > 0) are LINEAR functions of the pixel coordinates because
//     the ray direction is (pixel_to_ray * (px, py, 1)); the three functions, evaluated with generous margins, reject
//     32x32 tiles, 8x4 blocks and single pixels that the triangle cannot cover (homogeneous rasterisation: no near-plane
//     clipping, no division);
//   * a pixel that survives runs EXACTLY the decision sequence of bvh_closest_front (tri_terms / tri_exact / depth clip with
//     the same roundings), so the buffer is bit-identical to the ray-cast one;
//   * raster_setup_kernel (one thread per triangle) culls back faces and triangles in front of the near plane, rasterises
//     triangles with a small screen bounding box itself and queues the others as units of 32x32 pixels;
//     raster_tiles_kernel (persistent warps, one unit at a time through a ticket) classifies a unit's 32 blocks of 8x4
//     pixels one per lane and tests the pixels of the surviving blocks one per lane;
//   * raster_resolve_kernel turns the keys into the u32 ids (index | emitter << 31, 0xFFFFFFFF = background).
#pragma once
#include "kernels.cuh"

namespace RL_NS {

#define RL_RASTER_SMALL 48u            // largest bounding box (pixels) that the setup thread rasterises itself
// A unit of raster_tiles_kernel is 1024 pixels = 32 blocks of 8x4, one per lane: 2^tile_shift_x pixels wide and
// 2^(10 - tile_shift_x) of the device's LOCAL rows high. 32 x 32 when the device owns every row; a device that owns every
// N-th stripe sees triangles squeezed to 1 / N of their height in its row space, where wide, flat units (128 x 8) cover them
// with fewer units (the per-unit chain -- ticket, item search, classification -- is what the kernel costs at small shares).
#define RL_RASTER_MAX_ITEMS (1u << 20) // queued triangles; scenes with more triangles than that use the BVH walk (api.cu)

struct __align__(16) RasterItem {
	uint32_t tri, first_unit;      // slot in SceneView::tris; number of this triangle's first unit
	uint16_t tx0, ty0, ntx, nty;   // its tiles: origin and counts, in tiles
	float fu[3], fv[3], fw[3];     // the three edge functions A px + B py + C (>= 0 inside, scaled by det > 0)
	float slack[3];                // bound on their evaluation error anywhere on the screen
};

struct RasterBuffers {
	unsigned long long* zbuf;        // [pixel_count] {bits(t) << 32 | index << 1 | emitter}, ~0 = background
	RasterItem* items;
	unsigned long long* counter;     // {items << 32 | units}, bumped by one atomic per queued triangle
	unsigned int* ticket;            // next unit of raster_tiles_kernel
	uint32_t tile_shift_x;           // log2 of the unit width in pixels, 5 .. 8 (the unit height is 2^(10 - tile_shift_x) local rows)
};

// slack[k]: 1e-4 of the magnitudes of the terms that function k is summed from, anywhere on the screen -- hundreds of ulps of
// the fp32 evaluation (of the coefficients and of the sum), a fraction of a pixel in distance
struct EdgeFunctions { float fu[3], fv[3], fw[3], slack[3]; };

// Is the largest value of A px + B py + C over the pixel rectangle clearly negative?
__device__ __forceinline__ bool edge_rejects(const float* e, float slack, float x0, float y0, float x1, float y1) {
	return fmaf(e[0], e[0] > 0.0f ? x1 : x0, fmaf(e[1], e[1] > 0.0f ? y1 : y0, e[2])) < -slack;
}
__device__ __forceinline__ bool rect_rejected(const EdgeFunctions& ef, float x0, float y0, float x1, float y1) {
	return edge_rejects(ef.fu, ef.slack[0], x0, y0, x1, y1) || edge_rejects(ef.fv, ef.slack[1], x0, y0, x1, y1) || edge_rejects(ef.fw, ef.slack[2], x0, y0, x1, y1);
}
__device__ __forceinline__ float edge_magnitude(float3 c0, float3 c1, float3 c2, float3 g, float w, float h) {
	const float3 a = mk3(fabsf(g.x), fabsf(g.y), fabsf(g.z));
	return dot3(mk3(fabsf(c0.x), fabsf(c0.y), fabsf(c0.z)), a) * w + dot3(mk3(fabsf(c1.x), fabsf(c1.y), fabsf(c1.z)), a) * h + dot3(mk3(fabsf(c2.x), fabsf(c2.y), fabsf(c2.z)), a);
}

// The per-frame quantities of the depth clip (rows 2 and 3 of world_to_projection applied to the ray) and the ray itself
struct RasterFrame {
	float3 o;
	float zo, wo;
};
__device__ __forceinline__ RasterFrame raster_frame(const FrameUniforms& f) {
	RasterFrame r;
	r.o = mk3(f.camera[0], f.camera[1], f.camera[2]);
	const float (*w2p)[4] = f.world_to_projection;
	r.zo = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w2p[2][0], r.o.x), __fmul_rn(w2p[2][1], r.o.y)), __fmul_rn(w2p[2][2], r.o.z)), w2p[2][3]);
	r.wo = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w2p[3][0], r.o.x), __fmul_rn(w2p[3][1], r.o.y)), __fmul_rn(w2p[3][2], r.o.z)), w2p[3][3]);
	return r;
}

// One pixel against one triangle. Cheap rejections first: (a) the three edge functions at the pixel, (b) EARLY Z -- along a
// ray t = st / det where st = e2 . (s x e1) does not depend on the ray, so K / (Fu + Fv + Fw) estimates t for a few
// instructions; a pixel whose estimate is clearly behind the key already in the buffer is skipped (the buffer is read without
// synchronisation: a stale value is only larger, i.e. conservative). Whoever is left runs the candidate test of
// bvh_closest_front (bvh.cuh) with its exact roundings and competes with atomicMin.
// (x, y) global pixel coordinates of a pixel this device owns; `pixel` its local index; K = 0 disables early z.
__device__ __forceinline__ void raster_pixel(const FrameUniforms& f, const RasterFrame& rf, const BvhTri& tri, const EdgeFunctions& ef, float K,
	uint32_t x, uint32_t y, uint32_t pixel, unsigned long long* zbuf)
{
	const float fx = (float) x, fy = (float) y;
	const float eu = fmaf(ef.fu[0], fx, fmaf(ef.fu[1], fy, ef.fu[2])), ev = fmaf(ef.fv[0], fx, fmaf(ef.fv[1], fy, ef.fv[2])), ew = fmaf(ef.fw[0], fx, fmaf(ef.fw[1], fy, ef.fw[2]));
	if (eu < -ef.slack[0] || ev < -ef.slack[1] || ew < -ef.slack[2]) return;
	const unsigned long long current = __ldcg(&zbuf[pixel]);   // from L2: a line cached in L1 would stay stale for the life of the CTA
	const float current_t = __uint_as_float((uint32_t) (current >> 32));   // NaN while the pixel is background: never "behind"
	const float det_estimate = eu + ev + ew;
	// slack[2] is ~100x the rounding error of det_estimate: above 10 slack[2] the estimate of t is good to ~1e-3
	if (det_estimate > 10.0f * ef.slack[2] && K * approx_rcp(det_estimate) * 0.998f > current_t) return;
	const float3 d = primary_ray(f, x, y);
	const TriTerms k = tri_terms(tri, rf.o, d);
	if (!(k.det > 0.0f)) return;
	const float r = approx_rcp(k.det);
	const float ua = k.su * r, va = k.sv * r, ta = k.st * r;
	if (ua < -RL_TRI_TINY || va < -RL_TRI_TINY || ua + va > 1.0f + RL_TRI_EPS || ta < -RL_TRI_TINY) return;
	float t;
	if (!tri_exact(k, t)) return;
	if (!(t > 0.0f)) return;
	const float (*w2p)[4] = f.world_to_projection;
	const float zd = __fadd_rn(__fadd_rn(__fmul_rn(w2p[2][0], d.x), __fmul_rn(w2p[2][1], d.y)), __fmul_rn(w2p[2][2], d.z));
	const float wd = __fadd_rn(__fadd_rn(__fmul_rn(w2p[3][0], d.x), __fmul_rn(w2p[3][1], d.y)), __fmul_rn(w2p[3][2], d.z));
	const float zc = __fadd_rn(rf.zo, __fmul_rn(t, zd)), wc = __fadd_rn(rf.wo, __fmul_rn(t, wd));
	if (!(zc >= 0.0f) || !(zc <= wc)) return;
	const uint32_t id = __float_as_uint(tri.v0.w);
	const unsigned long long key = ((unsigned long long) __float_as_uint(t) << 32) | (unsigned long long) ((id << 1) | (id >> 31));
	if (key < current) atomicMin(&zbuf[pixel], key);
}
// st = e2 . (s x e1) = -(s . (e2 x e1)) of the triangle, or 0 (no early z) where it is too small against its terms to be trusted
__device__ __forceinline__ float early_z_constant(float3 sv, float3 g_det) {
	const float facing = dot3(sv, g_det);
	const float magnitude = fabsf(sv.x * g_det.x) + fabsf(sv.y * g_det.y) + fabsf(sv.z * g_det.z);
	return (-facing > 1.0e-3f * magnitude) ? -facing : 0.0f;
}

// Local index of the first row this device owns at or below the global row y (== owned_rows if there is none): the
// rasteriser works in the device's own row space, so that a device that owns 1 / N of the rows does 1 / N of the work
__device__ __forceinline__ uint32_t first_owned_local(const Stripes& st, uint32_t y) {
	const uint32_t band = y / st.stripe_h;
	const uint32_t before = band > st.stripe_index ? (band - st.stripe_index + st.stripe_count - 1u) / st.stripe_count : 0u;   // owned bands above
	return before * st.stripe_h + ((band % st.stripe_count == st.stripe_index) ? y - band * st.stripe_h : 0u);
}

__global__ void __launch_bounds__(128) raster_setup_kernel(SceneView s, FrameUniforms f, Stripes st, RasterBuffers rb) {
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= s.triangle_count) return;
	const BvhTri tri = s.tris[slot];
	const RasterFrame rf = raster_frame(f);
	const float3 v0 = mk3(tri.v0.x, tri.v0.y, tri.v0.z), e1 = mk3(tri.e1.x, tri.e1.y, tri.e1.z), e2 = mk3(tri.e2.x, tri.e2.y, tri.e2.z);
	const float3 sv = mk3(rf.o.x - v0.x, rf.o.y - v0.y, rf.o.z - v0.z);
	// det = d . (e2 x e1), su = d . (e2 x s), sv = d . (s x e1): linear in the ray direction, hence in (px, py)
	const float3 g_det = cross3(e2, e1), g_u = cross3(e2, sv), g_v = cross3(sv, e1);
	// a ray through a point p of the triangle has d ~ p - o, so det ~ (p - o) . (e2 x e1) = -(sv . g_det) whatever p: the
	// triangle is a back face (det <= 0) for every pixel when sv . g_det is clearly positive
	const float facing = dot3(sv, g_det);
	const float len_s = sqrtf(dot3(sv, sv)), len_n = sqrtf(dot3(g_det, g_det));
	if (facing > 1.0e-4f * len_s * len_n) return;
	const float3 c0 = mk3(f.pixel_to_ray[0][0], f.pixel_to_ray[1][0], f.pixel_to_ray[2][0]);
	const float3 c1 = mk3(f.pixel_to_ray[0][1], f.pixel_to_ray[1][1], f.pixel_to_ray[2][1]);
	const float3 c2 = mk3(f.pixel_to_ray[0][2], f.pixel_to_ray[1][2], f.pixel_to_ray[2][2]);
	const float3 g_w = mk3(g_det.x - g_u.x - g_v.x, g_det.y - g_u.y - g_v.y, g_det.z - g_u.z - g_v.z);
	EdgeFunctions ef;
	ef.fu[0] = dot3(c0, g_u); ef.fu[1] = dot3(c1, g_u); ef.fu[2] = dot3(c2, g_u);
	ef.fv[0] = dot3(c0, g_v); ef.fv[1] = dot3(c1, g_v); ef.fv[2] = dot3(c2, g_v);
	ef.fw[0] = dot3(c0, g_w); ef.fw[1] = dot3(c1, g_w); ef.fw[2] = dot3(c2, g_w);
	ef.slack[0] = 1.0e-4f * edge_magnitude(c0, c1, c2, g_u, (float) f.width, (float) f.height);
	ef.slack[1] = 1.0e-4f * edge_magnitude(c0, c1, c2, g_v, (float) f.width, (float) f.height);
	ef.slack[2] = ef.slack[0] + ef.slack[1] + 1.0e-4f * edge_magnitude(c0, c1, c2, g_det, (float) f.width, (float) f.height);
	// ---- screen bounding box by projection; triangles that reach (half-way) towards the camera plane take the whole screen
	const float (*w2p)[4] = f.world_to_projection;
	float min_x = INFINITY, max_x = -INFINITY, min_y = INFINITY, max_y = -INFINITY, max_z = -INFINITY;
	bool projectable = rf.zo < 0.0f;
	#pragma unroll
	for (int i = 0; i != 3; ++i) {
		const float3 p = (i == 0) ? v0 : (i == 1) ? add3(v0, e1) : add3(v0, e2);
		const float cx = w2p[0][0] * p.x + w2p[0][1] * p.y + w2p[0][2] * p.z + w2p[0][3];
		const float cy = w2p[1][0] * p.x + w2p[1][1] * p.y + w2p[1][2] * p.z + w2p[1][3];
		const float cz = w2p[2][0] * p.x + w2p[2][1] * p.y + w2p[2][2] * p.z + w2p[2][3];
		const float cw = w2p[3][0] * p.x + w2p[3][1] * p.y + w2p[3][2] * p.z + w2p[3][3];
		max_z = fmaxf(max_z, cz);
		// z_clip rises from zo (< 0) at the camera plane to 0 at the near plane
		if (!(cz > 0.5f * rf.zo) || !(cw > 0.0f)) projectable = false;
		const float rw = 1.0f / cw;
		const float px = (cx * rw + 1.0f) * (0.5f * (float) f.width) - 0.5f, py = (cy * rw + 1.0f) * (0.5f * (float) f.height) - 0.5f;
		if (!(fabsf(px) < 1.0e7f) || !(fabsf(py) < 1.0e7f)) projectable = false;
		min_x = fminf(min_x, px); max_x = fmaxf(max_x, px); min_y = fminf(min_y, py); max_y = fmaxf(max_y, py);
	}
	// entirely in front of the near plane (z_clip < 0 everywhere, with a margin relative to the camera's own -zo): never visible
	if (max_z < 1.0e-3f * rf.zo) return;
	int x0 = 0, y0 = 0, x1 = (int) f.width - 1, y1 = (int) f.height - 1;
	if (projectable) {
		x0 = max(x0, (int) floorf(min_x) - 1); x1 = min(x1, (int) ceilf(max_x) + 1);
		y0 = max(y0, (int) floorf(min_y) - 1); y1 = min(y1, (int) ceilf(max_y) + 1);
		if (x0 > x1 || y0 > y1) return;
	}
	if (rect_rejected(ef, (float) x0, (float) y0, (float) x1, (float) y1)) return;
	// the rows of the box that this device owns, as a range of its local rows
	const uint32_t l0 = first_owned_local(st, (uint32_t) y0), l_end = min(first_owned_local(st, (uint32_t) y1 + 1u), st.owned_rows);
	if (l0 >= l_end) return;
	const uint32_t w = (uint32_t) (x1 - x0 + 1), h = l_end - l0;
	const float K = early_z_constant(sv, g_det);
	bool inline_raster = w * h <= RL_RASTER_SMALL;
	if (!inline_raster) {
		const uint32_t sx = rb.tile_shift_x, sy = 10u - sx;
		const uint32_t tx0 = (uint32_t) x0 >> sx, ty0 = l0 >> sy;
		const uint32_t ntx = ((uint32_t) x1 >> sx) - tx0 + 1u, nty = ((l_end - 1u) >> sy) - ty0 + 1u;
		const unsigned long long old = atomicAdd(rb.counter, (1ull << 32) + (unsigned long long) (ntx * nty));
		const uint32_t index = (uint32_t) (old >> 32);
		if (index < RL_RASTER_MAX_ITEMS) {
			RasterItem it;
			it.tri = slot; it.first_unit = (uint32_t) old;
			it.tx0 = (uint16_t) tx0; it.ty0 = (uint16_t) ty0; it.ntx = (uint16_t) ntx; it.nty = (uint16_t) nty;
			#pragma unroll
			for (int i = 0; i != 3; ++i) { it.fu[i] = ef.fu[i]; it.fv[i] = ef.fv[i]; it.fw[i] = ef.fw[i]; it.slack[i] = ef.slack[i]; }
			rb.items[index] = it;
		}
		else inline_raster = true;   // queue full: correct, only slow (its units are skipped by the tile kernel)
	}
	if (inline_raster) {
		for (uint32_t row = l0; row != l_end; ++row) {
			const uint32_t y = st.global_row(row);
			for (int x = x0; x <= x1; ++x) raster_pixel(f, rf, tri, ef, K, (uint32_t) x, y, row * f.width + (uint32_t) x, rb.zbuf);
		}
	}
}

__global__ void __launch_bounds__(128) raster_tiles_kernel(SceneView s, FrameUniforms f, Stripes st, RasterBuffers rb) {
	const uint32_t lane = threadIdx.x & 31u;
	const unsigned long long counter = *rb.counter;
	const uint32_t item_count = min((uint32_t) (counter >> 32), RL_RASTER_MAX_ITEMS), unit_count = (uint32_t) counter;
	if (item_count == 0u) return;
	const RasterFrame rf = raster_frame(f);
	// Every WARP takes units on its own (no CTA barriers), `chunk` consecutive units per ticket: with few units (a 1080p frame
	// of a small scene: ~4 per warp) one at a time keeps all warps busy to the end; with many (4K, 50 k triangles: > 100 per warp)
	// consecutive units mostly belong to the same triangle, whose item search and loads are then done once per chunk. The next
	// ticket is claimed before the current chunk is processed so that the atomic's round trip overlaps the pixel tests.
	const uint32_t chunk = min(8u, max(1u, unit_count / (gridDim.x * 4u * 16u)));
	const uint32_t sx = rb.tile_shift_x, sy = 10u - sx, across_shift = sx - 3u, across_mask = (1u << across_shift) - 1u;
	uint32_t next_base = 0;
	if (lane == 0) next_base = atomicAdd(rb.ticket, chunk);
	next_base = __shfl_sync(0xFFFFFFFFu, next_base, 0);
	RasterItem it;
	it.first_unit = 0xFFFFFFFFu; it.ntx = it.nty = 0;
	BvhTri tri;
	EdgeFunctions ef;
	float K = 0.0f;
	while (next_base < unit_count) {
		const uint32_t base = next_base, end = min(base + chunk, unit_count);
		if (lane == 0) next_base = atomicAdd(rb.ticket, chunk);
		next_base = __shfl_sync(0xFFFFFFFFu, next_base, 0);
		for (uint32_t unit = base; unit != end; ++unit) {
			if (unit - it.first_unit >= (uint32_t) it.ntx * it.nty) {
				// the item whose range of units contains `unit` (first_unit rises with the item index; dropped items leave a gap at the end)
				uint32_t lo = 0, hi = item_count;
				while (hi - lo > 1u) {
					const uint32_t mid = (lo + hi) >> 1;
					if (__ldg(&rb.items[mid].first_unit) <= unit) lo = mid; else hi = mid;
				}
				it = rb.items[lo];
				if (unit - it.first_unit >= (uint32_t) it.ntx * it.nty) continue;   // a unit of a triangle that the setup thread kept for itself
				#pragma unroll
				for (int i = 0; i != 3; ++i) { ef.fu[i] = it.fu[i]; ef.fv[i] = it.fv[i]; ef.fw[i] = it.fw[i]; ef.slack[i] = it.slack[i]; }
				tri = s.tris[it.tri];
				K = early_z_constant(mk3(rf.o.x - tri.v0.x, rf.o.y - tri.v0.y, rf.o.z - tri.v0.z),
					cross3(mk3(tri.e2.x, tri.e2.y, tri.e2.z), mk3(tri.e1.x, tri.e1.y, tri.e1.z)));
			}
			const uint32_t local = unit - it.first_unit;
			// tiles are 2^sx pixels wide and 2^sy of this device's LOCAL rows high
			// local / ntx without an integer division (5.6 % of the kernel's instructions): local < 2^20 and ntx <= 2^10, so the
			// product with the approximate reciprocal is off by < 2.5e-4 while (local + 0.5) / ntx stays 0.5 / ntx >= 4.9e-4 away from an integer
			// (images beyond 32 k pixels in either direction take the division)
			const uint32_t tile_j = (it.ntx <= 1024u && it.nty <= 1024u) ? (uint32_t) (((float) local + 0.5f) * approx_rcp((float) it.ntx)) : local / it.ntx;
			const uint32_t tile_i = local - tile_j * it.ntx;
			const uint32_t tile_x = (it.tx0 + tile_i) << sx, tile_row = (it.ty0 + tile_j) << sy;
			const float tx1 = (float) min(tile_x + (1u << sx), f.width) - 1.0f;
			const float ty0 = (float) st.global_row(tile_row), ty1 = (float) st.global_row(min(tile_row + (1u << sy), st.owned_rows) - 1u);
			if (rect_rejected(ef, (float) tile_x, ty0, tx1, ty1)) continue;
			// 32 blocks of 8x4 pixels (2^(sx - 3) across), classified one per lane
			const uint32_t bx = tile_x + (lane & across_mask) * 8u, brow = tile_row + (lane >> across_shift) * 4u;
			bool live = bx < f.width && brow < st.owned_rows;
			if (live) live = !rect_rejected(ef, (float) bx, (float) st.global_row(brow), (float) min(bx + 7u, f.width - 1u), (float) st.global_row(min(brow + 3u, st.owned_rows - 1u)));
			unsigned todo = __ballot_sync(0xFFFFFFFFu, live);
			while (todo) {
				const uint32_t b = (uint32_t) __ffs(todo) - 1u;
				todo &= todo - 1u;
				const uint32_t x = tile_x + (b & across_mask) * 8u + (lane & 7u), row = tile_row + (b >> across_shift) * 4u + (lane >> 3);
				if (x < f.width && row < st.owned_rows) raster_pixel(f, rf, tri, ef, K, x, st.global_row(row), row * f.width + x, rb.zbuf);
			}
		}
	}
}

// Keys -> ids; the kernel also leaves the key buffer, the item counter and the unit ticket cleared for the next frame that uses
// this set of buffers (they are cleared once when they are allocated), so that a frame needs no memset operations.
// Four pixels per thread with 16-byte accesses (one pixel per thread ran at a seventh of the HBM rate: 45 us for the 41 MB of a
// 1080p frame); the tail of a buffer whose size is not a multiple of four goes one pixel at a time.
__device__ __forceinline__ uint32_t raster_key_to_id(unsigned long long key) {
	const uint32_t low = (uint32_t) key;
	return (key == ~0ull) ? 0xFFFFFFFFu : ((low >> 1) | (low << 31));
}
__global__ void __launch_bounds__(256) raster_resolve_kernel(unsigned long long* zbuf, uint32_t* visibility, uint32_t pixel_count, unsigned long long* counter_and_ticket) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t == 0u) { counter_and_ticket[0] = 0ull; counter_and_ticket[1] = 0ull; }
	const uint32_t i = 4u * t;
	if (i + 3u < pixel_count) {
		ulonglong2* keys = (ulonglong2*) (zbuf + i);
		const ulonglong2 a = keys[0], b = keys[1];
		keys[0] = make_ulonglong2(~0ull, ~0ull); keys[1] = make_ulonglong2(~0ull, ~0ull);
		*(uint4*) (visibility + i) = make_uint4(raster_key_to_id(a.x), raster_key_to_id(a.y), raster_key_to_id(b.x), raster_key_to_id(b.y));
	}
	else for (uint32_t k = i; k < pixel_count; ++k) {
		const unsigned long long key = zbuf[k];
		zbuf[k] = ~0ull;
		visibility[k] = raster_key_to_id(key);
	}
}

}  // namespace RL_NS
