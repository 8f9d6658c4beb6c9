// api.cu -- C ABI of librisltc_cuda.so (include/risltc_cuda.h) and kernel launches.
#define RL_PSA_ATTR __forceinline__   // the production kernels call the PSA functions once per loop body (shading.cuh)
#include "internal.h"
#include "kernels.cuh"
#include "shade_fast.cuh"
#include "trace4.cuh"
#include "raster.cuh"
#include "bvh_build.h"
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <utility>
#include <chrono>

using namespace exact;

static thread_local std::string g_last_error;

int rl_fail(const char* what, const char* detail) {
	g_last_error = std::string(what) + (detail ? std::string(": ") + detail : std::string());
	printf("risltc_cuda: %s\n", g_last_error.c_str());
	return 1;
}
#define fail rl_fail

#define RL_FRAME_EVENTS 6
#define RL_TRACE_CANDIDATES 3u
#define RL_TRACE_TRIALS (2u * RL_TRACE_CANDIDATES)   // every candidate is timed on two frames (the faster counts: the first frames after an upload run on cold caches)
#define RL_TRACE_DECIDED (RL_TRACE_TRIALS + 1u)
// auto: scenes from 8 M triangles on are built on the device. Measured on B200: 5 M triangles
// (C4) build in 7-8 s on the host and in 24 ms on the device (PLOC; 14.6 ms of kernels). The clustered tree is as good as the
// SAH tree for shadow rays (C2 -3 %, C3 +3 %, C4 +9 % kernel time) but costs the per-pixel BVH walk of C4 +40 %: a scene is
// built once and rendered for thousands of frames, so the host's binned-SAH tree stays the default until its build takes ~10 s.
#define RL_GPU_BUILD_TRIANGLES (1ull << 23)

struct risltc_device_s {
	int ordinal = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };   // [0] batch start, [4] batch end
	// scene
	uint2* positions = nullptr; ushort4* normals_uv = nullptr; uint8_t* material_indices = nullptr;
	TextureDesc* textures = nullptr; unsigned char* texels = nullptr; float* srgb_table = nullptr;
	float4* materials = nullptr; float4* lights = nullptr; float4* lights_tri = nullptr; ushort4* ltc_rgba = nullptr; ushort2* ltc_rg = nullptr;
	BvhNode* nodes = nullptr; BvhTri* tris = nullptr; Qbvh4Node* nodes4 = nullptr;
	SceneView view = {};
	float dequant_factor[3] = { 0, 0, 0 }, dequant_summand[3] = { 0, 0, 0 };
	// variant + targets
	Variant variant = { 1, TECH_LTC_CP, MIS_OPTIMAL_CLAMPED, 1, 1, 0, 3, 3 };
	uint32_t width = 0, height = 0;
	Stripes stripes = { 8, 0, 1, 0, 3 };
	PixelBuffers px = {};
	float4* own_accum = nullptr;
	uint32_t ray_slots = 0, group_slots = 0;
	// Frame overlap: consecutive frames of a render_frames call alternate between two streams and two sets of per-frame
	// buffers, so that the tail of one frame's persistent kernels (few warps left, SMs idling) is filled with the next
	// frame's work; only the accumulation (resolve) of frame i waits for frame i - 1. On by default up to 4.5 M pixels per
	// device (overlap_pays below), see risltc_cuda_set_frame_overlap.
	PixelBuffers px2 = {};
	RasterBuffers raster2 = {};
	bool set2_ready = false, overlap = false, overlap_pinned = false;
	cudaStream_t stream2 = nullptr;
	cudaEvent_t ev_resolved[2] = { nullptr, nullptr }, ev_fork = nullptr, ev_join = nullptr;
	uint32_t last_set = 0;
	uint32_t precision = RISLTC_PRECISION_FAST;
	int sm_count = 148, trace_resident = 1, trace4_resident = 1, trace4p_resident = 1, trace4pu_resident = 1;   // CTAs per SM of the shadow-ray kernels (pu: pairs, unordered)
	int trace_ctas_per_sm = 0;   // 0: as many as fit
	uint32_t refill = 4;   // idle lanes that trigger a refill of the warp from its staged rays (sweep with the pair kernel, round 2: 1 / 2 / 4 / 6 / 10 -> C2 trace -- / 10.19 / 10.08 / 10.18 / 10.49 ms, C3 231.4 / 226.8 / 226.8 / 229.7 / 239.2 ms; RISLTC_REFILL)
	// (1) has two bit-identical implementations: 1 = triangle-parallel rasteriser (raster.cuh; wins when few triangles cover
	// the screen), 0 = per-pixel BVH walk (gbuffer_kernel; wins at high depth complexity). Unless RISLTC_GBUFFER pins one, the
	// first two frames after a scene upload / resize run one each, timed with events, and the faster one is kept.
	uint32_t gbuffer_kind = 1;
	uint32_t gbuffer_tune = 0;    // 0: time the rasteriser next, 1: time the BVH walk next, 2: both in flight, 3: decided
	cudaEvent_t tune_ev[4] = { nullptr, nullptr, nullptr, nullptr };
	bool gbuffer_pinned = false;
	RasterBuffers raster = {};
	bool winner_cr = false;          // winner_cr.cu: correctly rounded transcendental functions in the winner's estimator (RISLTC_WINNER=cr)
	uint32_t raster_tile_shift = 0;  // RISLTC_RASTER_TILE_SHIFT=5..8: log2 of the rasteriser's unit width (tuning knob; 0 = by share)
	uint32_t ris_warps = 0;          // RISLTC_RIS_WARPS=<n>: cap on the warps of the RIS kernel's one CTA per SM (tuning knob; 0 = all that fit, <= 24)
	uint32_t winner_threads = 384;   // CTA size of the phase-synchronous winner kernel, two CTAs per SM: 384 (80 registers), 320 (96) or 256 (128)
	bool count_traversal = false; // trace4_kernel<true>: node visits and triangle tests are counted (risltc_cuda_traversal_counters)
	uint32_t trace_kind = 8;      // 8: trace4p_kernel (4-wide quantised tree, the two rays of a pixel per lane), 4: trace4_kernel (one ray per lane), 2: trace_kernel (binary tree)
	// lanes that must have a triangle waiting before the shadow-ray kernel runs its triangle track. Few occluded rays (C2: 27 %):
	// a well filled triangle track (8) wins; mostly occluded rays in a deep tree (C4: 93 %): testing the first triangle at once
	// (1) ends them ~10 % of their node visits earlier. Unless RISLTC_TRI_VOTE pins it, the first two frames after a scene
	// upload run one candidate each, timed with events, and the faster is kept (the image does not depend on it).
	uint32_t tri_vote = 8;
	// ... and, for the pair kernel, whether hit children are visited nearest first (trace4.cuh). RL_TRACE_CANDIDATES settings are timed.
	bool trace_ordered = true;
	uint32_t trace_tune = 0;      // 0 .. RL_TRACE_TRIALS - 1: time candidate (trace_tune % RL_TRACE_CANDIDATES) next, RL_TRACE_TRIALS: all in flight, RL_TRACE_DECIDED: decided
	bool trace_pinned = false;
	cudaEvent_t trace_tune_ev[2 * RL_TRACE_TRIALS] = {};
	// acceleration-structure builder: 0 = host (binned SAH, bvh_build.cpp), 1 = device (Morton-order radix tree, bvh_gpu.cu), 2 = auto:
	// the device builder from RL_GPU_BUILD_TRIANGLES triangles on, where the host build takes ~10 s (RISLTC_BVH_BUILD=host|gpu)
	uint32_t bvh_builder = 2;
	uint32_t ploc_radius = 16;   // device builder: search radius of the clustering (RISLTC_BVH_PLOC_RADIUS; 0 = radix tree over the Morton order)
	double bvh_stats[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };   // builder used, wall ms of the build, device ms x3, binary node slots, 4-wide nodes, binary depth << 16 | 4-wide depth
	uint64_t node_count = 0, node4_count = 0;
	unsigned long long launches = 0;
	bool timed = false;
	// per-frame events of the last batch: RL_FRAME_EVENTS per frame (before (1), after (1), after (2a), after (2b) = after (2), after (3), after (4))
	std::vector<cudaEvent_t> frame_events;
	uint32_t timed_frames = 0;
};

int rl_use(risltc_device_t* d) {
	if (!d) return fail("null device", nullptr);
	CU(cudaSetDevice(d->ordinal));
	return 0;
}
#define use rl_use
const SceneView& rl_view(const risltc_device_t* d) { return d->view; }
int rl_sm_count(const risltc_device_t* d) { return d->sm_count; }

extern "C" const char* risltc_cuda_last_error(void) { return g_last_error.c_str(); }

extern "C" int risltc_cuda_create_device(risltc_device_t** device, int cuda_ordinal) {
	if (!device) return fail("create_device: null output pointer", nullptr);
	*device = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) return fail("create_device: no CUDA device (there is no CPU fallback)", cudaGetErrorString(e));
	if (cuda_ordinal < 0 || cuda_ordinal >= count) return fail("create_device: ordinal out of range", nullptr);
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, cuda_ordinal));
	if (prop.major < 10) return fail("create_device: kernels are built for sm_100a only", prop.name);
	risltc_device_t* d = new risltc_device_t();
	d->ordinal = cuda_ordinal;
	CU(cudaSetDevice(cuda_ordinal));
	CU(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithFlags(&d->stream2, cudaStreamNonBlocking));
	CU(cudaEventCreateWithFlags(&d->ev_resolved[0], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&d->ev_resolved[1], cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
	for (auto& ev : d->ev) CU(cudaEventCreate(&ev));
	CU(cudaFuncSetAttribute(ris_ltc3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_LIMIT));
	CU(cudaFuncSetAttribute(ris_ltc3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_LIMIT));
	CU(cudaFuncSetAttribute(ris_ltc3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_LIMIT));
	CU(cudaFuncSetAttribute(ris_ltc3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_LIMIT));
	CU(cudaFuncSetAttribute(ris_ltc4_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_LIMIT));
	CU(cudaFuncSetAttribute(ris_ltc4_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_LIMIT));
	CU(cudaFuncSetAttribute(ris_ltc4_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_LIMIT));
	CU(cudaFuncSetAttribute(ris_ltc4_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_LIMIT));
	d->sm_count = prop.multiProcessorCount;
	CU(cudaMalloc(&d->px.counters, 8 * sizeof(unsigned long long)));
	CU(cudaMemset(d->px.counters, 0, 8 * sizeof(unsigned long long)));
	CU(cudaMalloc(&d->px.ticket, 4 * sizeof(unsigned int)));
	CU(cudaMemset(d->px.ticket, 0, 4 * sizeof(unsigned int)));
	// rasteriser: item queue, {counter, ticket} in one 16-byte block
	CU(cudaMalloc(&d->raster.items, (size_t) RL_RASTER_MAX_ITEMS * sizeof(RasterItem)));
	CU(cudaMalloc(&d->raster.counter, 16));
	CU(cudaMemset(d->raster.counter, 0, 16));
	d->raster.ticket = (unsigned int*) (d->raster.counter + 1);
	CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->trace_resident, trace_kernel, 128, 0));
	CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->trace4_resident, trace4_kernel<false>, 128, 0));
	CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->trace4p_resident, trace4p_kernel<false>, 128, 0));
	CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->trace4pu_resident, trace4p_kernel<false, false>, 128, 0));
	if (const char* e = getenv("RISLTC_TRI_VOTE")) { d->tri_vote = (uint32_t) atoi(e); d->trace_pinned = true; d->trace_tune = RL_TRACE_DECIDED; }   // tuning knobs
	if (const char* e = getenv("RISLTC_TRACE_ORDERED")) { d->trace_ordered = atoi(e) != 0; d->trace_pinned = true; d->trace_tune = RL_TRACE_DECIDED; }
	for (auto& ev : d->trace_tune_ev) CU(cudaEventCreate(&ev));
	if (const char* e = getenv("RISLTC_WIN_THREADS")) { int t = atoi(e); d->winner_threads = (t == 256 || t == 320) ? (uint32_t) t : 384u; }
	if (const char* e = getenv("RISLTC_WINNER")) d->winner_cr = strcmp(e, "cr") == 0;
	if (const char* e = getenv("RISLTC_GBUFFER")) { d->gbuffer_kind = (strcmp(e, "bvh") == 0) ? 0u : 1u; d->gbuffer_tune = 3u; d->gbuffer_pinned = true; }
	for (auto& ev : d->tune_ev) CU(cudaEventCreate(&ev));
	if (const char* e = getenv("RISLTC_OVERLAP")) { d->overlap = atoi(e) != 0; d->overlap_pinned = true; }
	if (const char* e = getenv("RISLTC_TRACE_CTAS")) d->trace_ctas_per_sm = atoi(e);
	if (const char* e = getenv("RISLTC_REFILL")) d->refill = (uint32_t) atoi(e);
	if (const char* e = getenv("RISLTC_BVH_PLOC_RADIUS")) d->ploc_radius = (uint32_t) atoi(e);
	if (const char* e = getenv("RISLTC_RASTER_TILE_SHIFT")) { const int v = atoi(e); d->raster_tile_shift = (v >= 5 && v <= 8) ? (uint32_t) v : 0u; }
	if (const char* e = getenv("RISLTC_RIS_WARPS")) d->ris_warps = (uint32_t) atoi(e);
	if (const char* e = getenv("RISLTC_BVH_BUILD")) d->bvh_builder = (strcmp(e, "gpu") == 0 || strcmp(e, "device") == 0) ? 1u : (strcmp(e, "radix") == 0) ? 3u : (strcmp(e, "host") == 0) ? 0u : 2u;
	if (const char* e = getenv("RISLTC_TRACE")) d->trace_kind = (atoi(e) == 2) ? 2u : (atoi(e) == 4) ? 4u : 8u;
	*device = d;
	return 0;
}

static void free_second_set(risltc_device_t* d) {
	if (d->stream2) cudaStreamSynchronize(d->stream2);
	cudaFree(d->px2.visibility); cudaFree(d->px2.origin); cudaFree(d->px2.base); cudaFree(d->px2.group); cudaFree(d->px2.ray_a); cudaFree(d->px2.ray_b);
	cudaFree(d->px2.pick); cudaFree(d->px2.shade); cudaFree(d->px2.ticket); cudaFree(d->raster2.zbuf); cudaFree(d->raster2.items); cudaFree(d->raster2.counter);
	d->px2 = PixelBuffers(); d->raster2 = RasterBuffers();
	d->set2_ready = false; d->last_set = 0;
}

static void free_targets(risltc_device_t* d) {
	free_second_set(d);
	cudaFree(d->px.pick); d->px.pick = nullptr;
	cudaFree(d->px.shade); d->px.shade = nullptr;
	cudaFree(d->px.visibility); cudaFree(d->px.origin); cudaFree(d->px.base); cudaFree(d->px.group);
	cudaFree(d->px.ray_a); cudaFree(d->px.ray_b); cudaFree(d->own_accum); cudaFree(d->raster.zbuf); d->raster.zbuf = nullptr;
	d->px.visibility = nullptr; d->px.origin = d->px.base = d->px.group = d->px.ray_a = d->px.ray_b = nullptr;
	d->own_accum = nullptr; d->px.accum = nullptr;
	d->ray_slots = d->group_slots = 0;
}

static void free_scene(risltc_device_t* d) {
	cudaFree(d->positions); cudaFree(d->normals_uv); cudaFree(d->material_indices); cudaFree(d->nodes); cudaFree(d->tris); cudaFree(d->nodes4);
	d->positions = nullptr; d->normals_uv = nullptr; d->material_indices = nullptr; d->nodes = nullptr; d->tris = nullptr; d->nodes4 = nullptr;
}

extern "C" void risltc_cuda_destroy_device(risltc_device_t* d) {
	if (!d) return;
	cudaSetDevice(d->ordinal);
	if (d->stream) cudaStreamSynchronize(d->stream);
	free_targets(d); free_scene(d);
	cudaFree(d->textures); cudaFree(d->texels); cudaFree(d->srgb_table);
	cudaFree(d->materials); cudaFree(d->lights); cudaFree(d->lights_tri); cudaFree(d->ltc_rgba); cudaFree(d->ltc_rg); cudaFree(d->px.counters); cudaFree(d->px.ticket); cudaFree(d->raster.items); cudaFree(d->raster.counter);
	for (auto& ev : d->ev) if (ev) cudaEventDestroy(ev);
	for (auto& ev : d->frame_events) if (ev) cudaEventDestroy(ev);
	for (auto& ev : d->tune_ev) if (ev) cudaEventDestroy(ev);
	for (auto& ev : d->trace_tune_ev) if (ev) cudaEventDestroy(ev);
	if (d->stream) cudaStreamDestroy(d->stream);
	if (d->stream2) cudaStreamDestroy(d->stream2);
	for (auto& ev : d->ev_resolved) if (ev) cudaEventDestroy(ev);
	if (d->ev_fork) cudaEventDestroy(d->ev_fork);
	if (d->ev_join) cudaEventDestroy(d->ev_join);
	delete d;
}

extern "C" int risltc_cuda_upload_scene(risltc_device_t* d, const uint32_t* quantized_positions, const uint16_t* normals_and_tex_coords,
	const uint8_t* material_indices, uint64_t T, const float factor[3], const float summand[3])
{
	if (use(d)) return 1;
	if (!quantized_positions || !normals_and_tex_coords || !material_indices || T == 0) return fail("upload_scene: the mesh is empty", nullptr);
	if (T >= (1ull << 27)) return fail("upload_scene: more than 2^27 triangles", nullptr);
	CU(cudaStreamSynchronize(d->stream));
	free_scene(d);
	memcpy(d->dequant_factor, factor, 12); memcpy(d->dequant_summand, summand, 12);
	CU(cudaMalloc(&d->positions, T * 3 * sizeof(uint2)));
	CU(cudaMalloc(&d->normals_uv, T * 3 * sizeof(ushort4)));
	CU(cudaMalloc(&d->material_indices, T));
	CU(cudaMemcpy(d->positions, quantized_positions, T * 3 * sizeof(uint2), cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d->normals_uv, normals_and_tex_coords, T * 3 * sizeof(ushort4), cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d->material_indices, material_indices, T, cudaMemcpyHostToDevice));
	// acceleration structure
	uint32_t max_leaf = 2;   // measured best for the 4-wide any-hit kernel (a triangle test costs about as much as three box tests)
	if (const char* e = getenv("RISLTC_BVH_LEAF")) max_leaf = (uint32_t) atoi(e);
	const bool on_device = (d->bvh_builder == 1u || d->bvh_builder == 3u || (d->bvh_builder == 2u && T >= RL_GPU_BUILD_TRIANGLES)) && T > 16;
	const auto wall_start = std::chrono::steady_clock::now();
	for (double& v : d->bvh_stats) v = 0.0;
	if (on_device) {
		uint64_t counts[2]; uint32_t depths[2]; float ms[3];
		if (rl_build_bvh_gpu(d->positions, T, factor, summand, max_leaf, d->bvh_builder == 3u ? 0u : d->ploc_radius, &d->nodes, &d->tris, &d->nodes4, counts, depths, ms)) return 1;
		if (depths[0] > RL_STACK) return fail("upload_scene: the device-built binary acceleration structure is deeper than the traversal stack (RISLTC_BVH_BUILD=host builds a balanced one)", nullptr);
		if (3u * depths[1] + 1u > RL_T4_OVERFLOW) return fail("upload_scene: the device-built acceleration structure is deeper than the traversal stack", nullptr);
		d->node_count = counts[0]; d->node4_count = counts[1];
		d->bvh_stats[0] = (d->bvh_builder == 3u || d->ploc_radius == 0u) ? 3.0 : 1.0; d->bvh_stats[2] = ms[0]; d->bvh_stats[3] = ms[1]; d->bvh_stats[4] = ms[2];
		d->bvh_stats[7] = (double) ((depths[0] << 16) | depths[1]);
	}
	else {
		std::vector<float> verts;
		dequantize_mesh_for_bvh(quantized_positions, T, factor, summand, verts);
		std::vector<BvhNodeHost> nodes; std::vector<uint32_t> order;
		build_bvh(verts.data(), T, nodes, order, max_leaf);
		// gbuffer_kernel, the exact-precision paths and trace_kernel walk the binary tree with RL_STACK entries per thread
		const uint32_t depth2 = bvh_depth(nodes);
		if (depth2 > RL_STACK) return fail("upload_scene: the binary acceleration structure is deeper than the traversal stack", nullptr);
		std::vector<BvhNode> dn(nodes.size());
		for (size_t i = 0; i != nodes.size(); ++i) {
			const BvhNodeHost& n = nodes[i];
			dn[i].a = make_float4(n.left_lo[0], n.left_lo[1], n.left_lo[2], n.left_hi[0]);
			dn[i].b = make_float4(n.left_hi[1], n.left_hi[2], n.right_lo[0], n.right_lo[1]);
			dn[i].c = make_float4(n.right_lo[2], n.right_hi[0], n.right_hi[1], n.right_hi[2]);
			dn[i].d = make_int4(n.left, n.right, 0, 0);
		}
		std::vector<BvhTri> dt(T);
		for (uint64_t slot = 0; slot != T; ++slot) {
			uint32_t t = order[slot];
			const float* v = verts.data() + 9 * (size_t) t;
			uint32_t id = t | ((quantized_positions[6 * (size_t) t + 1] >> 31) << 31);
			float idf; memcpy(&idf, &id, 4);
			volatile float e1x = v[3] - v[0], e1y = v[4] - v[1], e1z = v[5] - v[2];
			volatile float e2x = v[6] - v[0], e2y = v[7] - v[1], e2z = v[8] - v[2];
			dt[slot].v0 = make_float4(v[0], v[1], v[2], idf);
			dt[slot].e1 = make_float4(e1x, e1y, e1z, 0.0f);
			dt[slot].e2 = make_float4(e2x, e2y, e2z, 0.0f);
		}
		std::vector<Qbvh4NodeHost> dn4;
		const uint32_t depth4 = build_qbvh4(nodes, dn4);
		if (3u * depth4 + 1u > RL_T4_OVERFLOW) return fail("upload_scene: the acceleration structure is deeper than the traversal stack", nullptr);
		static_assert(sizeof(Qbvh4NodeHost) == sizeof(Qbvh4Node), "node layouts");
		CU(cudaMalloc(&d->nodes4, dn4.size() * sizeof(Qbvh4Node)));
		CU(cudaMemcpy(d->nodes4, dn4.data(), dn4.size() * sizeof(Qbvh4Node), cudaMemcpyHostToDevice));
		CU(cudaMalloc(&d->nodes, dn.size() * sizeof(BvhNode)));
		CU(cudaMalloc(&d->tris, dt.size() * sizeof(BvhTri)));
		CU(cudaMemcpy(d->nodes, dn.data(), dn.size() * sizeof(BvhNode), cudaMemcpyHostToDevice));
		CU(cudaMemcpy(d->tris, dt.data(), dt.size() * sizeof(BvhTri), cudaMemcpyHostToDevice));
		d->node_count = dn.size(); d->node4_count = dn4.size();
		d->bvh_stats[7] = (double) ((depth2 << 16) | depth4);
	}
	d->bvh_stats[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall_start).count();
	d->bvh_stats[5] = (double) d->node_count; d->bvh_stats[6] = (double) d->node4_count;
	d->view.positions = d->positions; d->view.normals_uv = d->normals_uv; d->view.material_indices = d->material_indices;
	d->view.nodes = d->nodes; d->view.nodes4 = d->nodes4; d->view.tris = d->tris; d->view.triangle_count = (uint32_t) T;
	if (!d->gbuffer_pinned) d->gbuffer_tune = 0;
	if (!d->trace_pinned) d->trace_tune = 0;
	return 0;
}

extern "C" int risltc_cuda_upload_materials(risltc_device_t* d, const float* m, uint64_t count) {
	if (use(d)) return 1;
	if (!m || count == 0) return fail("upload_materials: no materials", nullptr);
	CU(cudaStreamSynchronize(d->stream));
	cudaFree(d->materials); d->materials = nullptr;
	std::vector<float4> packed(2 * count);
	for (uint64_t i = 0; i != count; ++i) {
		const float* r = m + 8 * i;
		packed[2 * i] = make_float4(r[0], r[1], r[2], r[3]);
		packed[2 * i + 1] = make_float4(r[4], r[5], r[6], r[7]);   // linear roughness, metalicity, normal.r, normal.g
	}
	CU(cudaMalloc(&d->materials, packed.size() * sizeof(float4)));
	CU(cudaMemcpy(d->materials, packed.data(), packed.size() * sizeof(float4), cudaMemcpyHostToDevice));
	d->view.materials = d->materials;
	cudaFree(d->textures); cudaFree(d->texels); d->textures = nullptr; d->texels = nullptr;   // flat materials replace textures
	d->view.textures = nullptr; d->view.texels = nullptr;
	return 0;
}

extern "C" int risltc_cuda_upload_textures(risltc_device_t* d, const risltc_texture_t* textures, uint64_t count) {
	if (use(d)) return 1;
	if (!textures || count == 0 || count % 3 != 0) return fail("upload_textures: three textures per material (base colour, specular, normal)", nullptr);
	std::vector<TextureDesc> descs(count);
	size_t total = 0;
	for (uint64_t i = 0; i != count; ++i) {
		const risltc_texture_t& t = textures[i];
		if (!t.texels || t.width == 0 || t.height == 0 || t.mip_count == 0 || t.mip_count > 16 || t.format > RISLTC_TEXEL_RGBA8_SRGB)
			return fail("upload_textures: a texture needs texels, a positive extent, 1..16 mip levels and a known texel format", nullptr);
		size_t texel_count = 0;
		uint32_t w = t.width, h = t.height;
		for (uint32_t l = 0; l != t.mip_count; ++l) { texel_count += (size_t) w * h; w = (w > 1) ? w >> 1 : 1; h = (h > 1) ? h >> 1 : 1; }
		total = (total + 15) & ~(size_t) 15;
		descs[i].offset = total; descs[i].format = t.format; descs[i].width = t.width; descs[i].height = t.height; descs[i].levels = t.mip_count;
		descs[i].pad[0] = descs[i].pad[1] = 0;
		total += texel_count * (t.format == RISLTC_TEXEL_RGBA32F ? 16 : 4);
	}
	CU(cudaStreamSynchronize(d->stream));
	cudaFree(d->textures); cudaFree(d->texels); d->textures = nullptr; d->texels = nullptr;
	d->view.textures = nullptr; d->view.texels = nullptr;
	std::vector<unsigned char> staging(total);
	for (uint64_t i = 0; i != count; ++i) {
		const size_t end = (i + 1 != count) ? (size_t) descs[i + 1].offset : total;
		size_t bytes = 0;
		uint32_t w = textures[i].width, h = textures[i].height;
		for (uint32_t l = 0; l != textures[i].mip_count; ++l) { bytes += (size_t) w * h; w = (w > 1) ? w >> 1 : 1; h = (h > 1) ? h >> 1 : 1; }
		bytes *= (textures[i].format == RISLTC_TEXEL_RGBA32F ? 16 : 4);
		(void) end;
		memcpy(staging.data() + descs[i].offset, textures[i].texels, bytes);
	}
	CU(cudaMalloc(&d->texels, total ? total : 16));
	CU(cudaMemcpy(d->texels, staging.data(), total, cudaMemcpyHostToDevice));
	CU(cudaMalloc(&d->textures, count * sizeof(TextureDesc)));
	CU(cudaMemcpy(d->textures, descs.data(), count * sizeof(TextureDesc), cudaMemcpyHostToDevice));
	if (!d->srgb_table) {
		// sRGB byte -> linear, the exact curve in double precision rounded once (the oracle computes the same table)
		float table[256];
		for (int i = 0; i != 256; ++i) {
			const double c = (double) i / 255.0;
			table[i] = (float) ((c <= 0.04045) ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4));
		}
		CU(cudaMalloc(&d->srgb_table, sizeof(table)));
		CU(cudaMemcpy(d->srgb_table, table, sizeof(table), cudaMemcpyHostToDevice));
	}
	d->view.textures = d->textures; d->view.texels = d->texels; d->view.srgb_table = d->srgb_table;
	// the flat constants are not read while textures are bound, but render_frames checks that materials exist
	if (!d->materials) { CU(cudaMalloc(&d->materials, 2 * sizeof(float4))); CU(cudaMemset(d->materials, 0, 2 * sizeof(float4))); d->view.materials = d->materials; }
	return 0;
}

extern "C" int risltc_cuda_upload_lights(risltc_device_t* d, const void* records, uint32_t light_count, uint32_t max_vertex_count) {
	if (use(d)) return 1;
	if (!records || light_count == 0) return fail("upload_lights: no lights", nullptr);
	if (max_vertex_count < 3 || max_vertex_count > 7) return fail("upload_lights: polygons need 3 to 7 vertices", nullptr);
	size_t bytes = (size_t) light_count * (48 + 16 * max_vertex_count);
	if (d->view.light_count != light_count || d->view.light_stride4 != 3 + max_vertex_count) {
		CU(cudaStreamSynchronize(d->stream));
		cudaFree(d->lights); d->lights = nullptr;
		cudaFree(d->lights_tri); d->lights_tri = nullptr;
		CU(cudaMalloc(&d->lights, bytes));
		if (max_vertex_count <= 4) CU(cudaMalloc(&d->lights_tri, (size_t) light_count * 16 * max_vertex_count));
	}
	CU(cudaMemcpyAsync(d->lights, records, bytes, cudaMemcpyHostToDevice, d->stream));
	if (max_vertex_count <= 4) {
		// the candidate loop's view of a triangle / quad light: 48 bytes {v0 | Le.r, v1 | Le.g, v2 | Le.b} (+ {v3 | 0})
		const uint32_t V = max_vertex_count;
		const float* r = (const float*) records;
		std::vector<float> packed((size_t) light_count * 4 * V);
		for (uint32_t i = 0; i != light_count; ++i, r += 12 + 4 * V)
			for (uint32_t v = 0; v != V; ++v) {
				memcpy(&packed[4 * V * (size_t) i + 4 * v], r + 12 + 4 * v, 12);
				packed[4 * V * (size_t) i + 4 * v + 3] = (v < 3u) ? r[v] : 0.0f;
			}
		CU(cudaMemcpyAsync(d->lights_tri, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice, d->stream));
		CU(cudaStreamSynchronize(d->stream));   // `packed` is pageable stack-owned memory
	}
	d->view.lights_tri = d->lights_tri;
	d->view.lights = d->lights; d->view.light_count = light_count; d->view.light_stride4 = 3 + max_vertex_count;
	return 0;
}

extern "C" int risltc_cuda_upload_ltc(risltc_device_t* d, const uint16_t* rgba16, const uint16_t* rg16, uint32_t roughness_count, uint32_t inclination_count, uint32_t fresnel_count) {
	if (use(d)) return 1;
	if (!rgba16 || !rg16 || roughness_count == 0 || roughness_count != inclination_count || fresnel_count == 0)
		return fail("upload_ltc: tables must be square and non-empty", nullptr);
	CU(cudaStreamSynchronize(d->stream));
	cudaFree(d->ltc_rgba); cudaFree(d->ltc_rg); d->ltc_rgba = nullptr; d->ltc_rg = nullptr;
	size_t texels = (size_t) roughness_count * inclination_count * fresnel_count;
	CU(cudaMalloc(&d->ltc_rgba, texels * sizeof(ushort4)));
	CU(cudaMalloc(&d->ltc_rg, texels * sizeof(ushort2)));
	CU(cudaMemcpy(d->ltc_rgba, rgba16, texels * sizeof(ushort4), cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d->ltc_rg, rg16, texels * sizeof(ushort2), cudaMemcpyHostToDevice));
	d->view.ltc_rgba = d->ltc_rgba; d->view.ltc_rg = d->ltc_rg; d->view.ltc_res = roughness_count; d->view.ltc_layers = fresnel_count;
	return 0;
}

static int allocate_ray_buffers(risltc_device_t* d) {
	uint32_t groups = d->variant.light_samples, slots = groups * d->variant.sample_count * 2u;
	size_t pixels = d->px.pixel_count;
	if (!pixels) return 0;
	if (groups > d->group_slots || slots > d->ray_slots) free_second_set(d);   // rebuilt on demand with the new sizes
	if (groups > d->group_slots) {
		cudaFree(d->px.group); d->px.group = nullptr;
		CU(cudaMalloc(&d->px.group, pixels * groups * sizeof(float4)));
		d->group_slots = groups;
	}
	if (slots > d->ray_slots) {
		cudaFree(d->px.ray_a); cudaFree(d->px.ray_b); d->px.ray_a = d->px.ray_b = nullptr;
		CU(cudaMalloc(&d->px.ray_a, pixels * slots * sizeof(float4)));
		CU(cudaMalloc(&d->px.ray_b, pixels * slots * sizeof(float4)));
		d->ray_slots = slots;
	}
	CU(cudaMemsetAsync(d->px.ray_b, 0, pixels * d->ray_slots * sizeof(float4), d->stream));
	return 0;
}

// Measured on B200 with the round-2 kernels (bench.py, RISLTC_OVERLAP=0|1): overlapping frames wins
// 5-7 % on a whole 1080p frame (C2 33.0 -> 34.6, C4 17.2 -> 18.4 Gsamples/s), 9 % on half of one, 9-19 % on an eighth (the
// tails of the persistent kernels are then a fifth of a frame's time), 2 % on half a 4K frame (4.1 M pixels) and nothing on
// a whole 4K frame (8.3 M pixels), where the second set of per-frame buffers (140 bytes per pixel) is not worth its memory
static bool overlap_pays(const risltc_device_t* d) { return d->px.pixel_count <= 4500000u; }

// The second set of per-frame buffers (same sizes as the first), created the first time two frames overlap
static int ensure_second_set(risltc_device_t* d) {
	if (d->set2_ready) return 0;
	const size_t pixels = d->px.pixel_count;
	PixelBuffers& p = d->px2;
	p = PixelBuffers();
	p.pixel_count = d->px.pixel_count; p.counters = d->px.counters; p.accum = d->px.accum;
	CU(cudaMalloc(&p.visibility, pixels * sizeof(uint32_t)));
	CU(cudaMalloc(&p.origin, pixels * sizeof(float4)));
	CU(cudaMalloc(&p.base, pixels * sizeof(float4)));
	CU(cudaMalloc(&p.pick, pixels * sizeof(uint4)));
	CU(cudaMalloc(&p.shade, 6 * pixels * sizeof(float4)));
	CU(cudaMalloc(&p.group, pixels * d->group_slots * sizeof(float4)));
	CU(cudaMalloc(&p.ray_a, pixels * d->ray_slots * sizeof(float4)));
	CU(cudaMalloc(&p.ray_b, pixels * d->ray_slots * sizeof(float4)));
	CU(cudaMalloc(&p.ticket, 4 * sizeof(unsigned int)));
	CU(cudaMemset(p.ray_b, 0, pixels * d->ray_slots * sizeof(float4)));
	CU(cudaMemset(p.ticket, 0, 4 * sizeof(unsigned int)));
	CU(cudaMalloc(&d->raster2.zbuf, pixels * sizeof(unsigned long long)));
	CU(cudaMemset(d->raster2.zbuf, 0xFF, pixels * sizeof(unsigned long long)));   // kept clear by raster_resolve_kernel from here on
	CU(cudaMalloc(&d->raster2.items, (size_t) RL_RASTER_MAX_ITEMS * sizeof(RasterItem)));
	CU(cudaMalloc(&d->raster2.counter, 16));
	CU(cudaMemset(d->raster2.counter, 0, 16));
	d->raster2.ticket = (unsigned int*) (d->raster2.counter + 1);
	d->raster2.tile_shift_x = d->raster.tile_shift_x;
	d->set2_ready = true;
	return 0;
}

extern "C" int risltc_cuda_set_frame_overlap(risltc_device_t* d, uint32_t mode) {
	if (use(d)) return 1;
	if (mode > RISLTC_OVERLAP_AUTO) return fail("set_frame_overlap: unknown mode", nullptr);
	CU(cudaStreamSynchronize(d->stream));
	d->overlap_pinned = mode != RISLTC_OVERLAP_AUTO;
	d->overlap = (mode == RISLTC_OVERLAP_AUTO) ? overlap_pays(d) : mode == RISLTC_OVERLAP_ON;
	return 0;
}

extern "C" uint32_t risltc_cuda_frame_overlap_active(const risltc_device_t* d) { return (d && d->overlap) ? 1u : 0u; }

extern "C" int risltc_cuda_set_variant(risltc_device_t* d, const risltc_variant_t* v) {
	if (use(d)) return 1;
	if (!v) return fail("set_variant: null variant", nullptr);
	if (v->polygon_technique > TECH_LTC_CP || v->mis_heuristic > MIS_OPTIMAL || v->light_sampling > 1) return fail("set_variant: enum out of range", nullptr);
	if (v->sample_count == 0 || v->light_samples == 0) return fail("set_variant: sample counts must be positive", nullptr);
	if (v->max_light_vertices < 3 || v->max_light_vertices > 7 || v->min_light_vertices < 3 || v->min_light_vertices > v->max_light_vertices)
		return fail("set_variant: lights have 3 to 7 vertices (main.c:191-204) and min_light_vertices <= max_light_vertices", nullptr);
	CU(cudaStreamSynchronize(d->stream));
	memcpy(&d->variant, v, sizeof(Variant));
	return allocate_ray_buffers(d);
}

extern "C" int risltc_cuda_set_precision(risltc_device_t* d, uint32_t mode) {
	if (use(d)) return 1;
	if (mode > RISLTC_PRECISION_EXACT) return fail("set_precision: unknown mode", nullptr);
	CU(cudaStreamSynchronize(d->stream));
	d->precision = mode;
	return 0;
}

extern "C" int risltc_cuda_set_kernels(risltc_device_t* d, uint32_t gbuffer, uint32_t shadow) {
	if (use(d)) return 1;
	if (gbuffer > RISLTC_GBUFFER_AUTO || (shadow != RISLTC_SHADOW_BINARY && shadow != RISLTC_SHADOW_WIDE && shadow != RISLTC_SHADOW_PAIRS)) return fail("set_kernels: unknown kernel", nullptr);
	CU(cudaStreamSynchronize(d->stream));
	d->gbuffer_pinned = gbuffer != RISLTC_GBUFFER_AUTO;
	d->gbuffer_kind = (gbuffer == RISLTC_GBUFFER_BVH) ? 0u : 1u;
	d->gbuffer_tune = d->gbuffer_pinned ? 3u : 0u;
	d->trace_kind = shadow;
	return 0;
}

extern "C" int risltc_cuda_resize(risltc_device_t* d, uint32_t width, uint32_t height, uint32_t stripe_height, uint32_t stripe_index, uint32_t stripe_count) {
	if (use(d)) return 1;
	if (width == 0 || height == 0 || stripe_count == 0 || stripe_index >= stripe_count || stripe_height == 0) return fail("resize: bad extent or stripe layout", nullptr);
	CU(cudaStreamSynchronize(d->stream));
	// the render targets are recreated: accumulation restarts in the device's own buffer (a caller-owned buffer attached
	// with set_accum_buffer belongs to the old extent and has to be attached again, include/risltc_cuda.h)
	free_targets(d);
	uint32_t owned = 0;
	for (uint32_t y0 = stripe_index * stripe_height; y0 < height; y0 += stripe_height * stripe_count)
		owned += (y0 + stripe_height <= height) ? stripe_height : height - y0;
	d->width = width; d->height = height;
	d->stripes.stripe_h = stripe_height; d->stripes.stripe_index = stripe_index; d->stripes.stripe_count = stripe_count; d->stripes.owned_rows = owned;
	// rasteriser units: 32 x 32 pixels for a whole frame, wider and flatter the smaller the share of the rows (raster.cuh)
	d->raster.tile_shift_x = d->raster2.tile_shift_x = d->raster_tile_shift ? d->raster_tile_shift : (stripe_count >= 4u ? 7u : stripe_count >= 2u ? 6u : 5u);
	d->stripes.h_shift = 0xFFFFFFFFu;
	for (uint32_t k = 0; k != 32u; ++k) if (stripe_height == (1u << k)) d->stripes.h_shift = k;
	size_t pixels = (size_t) owned * width;
	d->px.pixel_count = (uint32_t) pixels;
	if (pixels == 0) return 0;
	CU(cudaMalloc(&d->px.visibility, pixels * sizeof(uint32_t)));
	CU(cudaMalloc(&d->px.origin, pixels * sizeof(float4)));
	CU(cudaMalloc(&d->px.base, pixels * sizeof(float4)));
	CU(cudaMalloc(&d->px.pick, pixels * sizeof(uint4)));
	CU(cudaMalloc(&d->px.shade, 6 * pixels * sizeof(float4)));
	CU(cudaMalloc(&d->raster.zbuf, pixels * sizeof(unsigned long long)));
	CU(cudaMemset(d->raster.zbuf, 0xFF, pixels * sizeof(unsigned long long)));   // kept clear by raster_resolve_kernel from here on
	CU(cudaMalloc(&d->own_accum, pixels * sizeof(float4)));
	CU(cudaMemset(d->own_accum, 0, pixels * sizeof(float4)));
	d->px.accum = d->own_accum;
	if (!d->gbuffer_pinned) d->gbuffer_tune = 0;
	if (!d->overlap_pinned) d->overlap = overlap_pays(d);
	return allocate_ray_buffers(d);
}

extern "C" int risltc_cuda_set_accum_buffer(risltc_device_t* d, void* device_pointer) {
	if (use(d)) return 1;
	CU(cudaStreamSynchronize(d->stream));
	d->px.accum = device_pointer ? (float4*) device_pointer : d->own_accum;
	return 0;
}

extern "C" uint32_t risltc_cuda_owned_rows(const risltc_device_t* d) { return d ? d->stripes.owned_rows : 0; }

extern "C" int risltc_cuda_owned_row_indices(const risltc_device_t* d, uint32_t* rows) {
	if (!d || !rows) return fail("owned_row_indices: null argument", nullptr);
	for (uint32_t r = 0; r != d->stripes.owned_rows; ++r) rows[r] = d->stripes.global_row(r);
	return 0;
}

// per_frame_constants_t (main.h:537-553, 256 bytes) -> the fields the kernels read
static void unpack_constants(FrameUniforms& f, const void* block, uint32_t accum_num) {
	const unsigned char* b = (const unsigned char*) block;
	memcpy(f.dequant_factor, b + 0, 12);
	memcpy(f.dequant_summand, b + 16, 12);
	memcpy(f.world_to_projection, b + 32, 64);
	memcpy(f.pixel_to_ray, b + 96, 48);
	memcpy(f.camera, b + 144, 12);
	memcpy(&f.mis_visibility_estimate, b + 156, 4);
	memcpy(&f.width, b + 160, 4); memcpy(&f.height, b + 164, 4);
	memcpy(&f.exposure, b + 176, 4); memcpy(&f.roughness_factor, b + 180, 4);
	memcpy(&f.frame_word, b + 208, 4);
	memcpy(f.ltc_constants, b + 224, 24);
	f.accum_num = accum_num;
}

static bool deferred_rays(const Variant& v) { return v.polygon_technique != TECH_TURK && v.polygon_technique != TECH_BASELINE && v.mis_heuristic != MIS_OPTIMAL; }

// Largest light table staged in shared memory (48 bytes per triangle light), keeping >= 2 CTAs per SM
static const uint32_t kMaxSmemLights = 2048;

// (2): the specialised persistent kernel of shade_fast.cuh when the variant is the default estimator on triangle
// lights and the device is in RISLTC_PRECISION_FAST, the generic kernel otherwise.
static int launch_shade(risltc_device_t* d, dim3 grid, const FrameUniforms& f, const PixelBuffers& px, cudaStream_t stream, cudaEvent_t between) {
	const Variant& v = d->variant;
	const bool defer = deferred_rays(v);
	const bool specialised = d->precision == RISLTC_PRECISION_FAST && v.light_sampling <= 1u && v.polygon_technique == TECH_LTC_CP
		&& v.mis_heuristic == MIS_OPTIMAL_CLAMPED && v.sample_count == 1u && v.light_samples == 1u && v.fast_atan == 0u
		&& ((v.max_light_vertices == 3u && v.min_light_vertices == 3u) || (v.max_light_vertices == 4u && v.min_light_vertices >= 3u))
		&& d->view.light_stride4 == 3u + v.max_light_vertices && d->view.lights_tri != nullptr;
	if (specialised) {
		const bool quads = v.max_light_vertices == 4u;
		const bool smem = d->view.light_count <= (quads ? kMaxSmemLights * 3u / 5u : kMaxSmemLights);
		const uint32_t staged = smem ? d->view.light_count : 0u;
		uint32_t warps = quads ? shade_fast4_warps(staged) : shade_fast_warps(staged);
		// a table beyond shared memory is gathered through L1: 16 warps (131 KB of shared memory) leave it ~100 KB of cache instead of
		// ~30 KB (C5, round 2: 4096 lights 0.55 -> 0.49 ms per frame, 16384 lights 0.48 -> 0.46; 12 warps lose again)
		if (!smem && warps > 16u) warps = 16u;
		if (d->ris_warps && d->ris_warps < warps) warps = d->ris_warps;
		const size_t bytes = quads ? shade_fast4_smem_bytes(staged, warps) : shade_fast_smem_bytes(staged, warps);
		// a warp owns 8x4 pixel tiles; one persistent CTA per SM
		const uint32_t tiles_x = (d->width + 7) / 8, tile_count = tiles_x * ((d->stripes.owned_rows + 3) / 4);
		uint32_t ctas = (uint32_t) d->sm_count;
		if (ctas * warps > tile_count) ctas = (tile_count + warps - 1) / warps;
		const bool textured = d->view.textures != nullptr;
		if (v.light_sampling == 0u) {
			// light_uniform: no candidates, one uniform draw per pixel (shade_fast.cuh)
			if (textured) pick_uniform_kernel<true><<<grid, 128, 0, stream>>>(d->view, f, d->stripes, px);
			else pick_uniform_kernel<false><<<grid, 128, 0, stream>>>(d->view, f, d->stripes, px);
		}
		else if (quads) {
			if (smem && !textured) ris_ltc4_kernel<true, false><<<ctas, 32 * warps, bytes, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count);
			else if (!textured) ris_ltc4_kernel<false, false><<<ctas, 32 * warps, bytes, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count);
			else if (smem) ris_ltc4_kernel<true, true><<<ctas, 32 * warps, bytes, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count);
			else ris_ltc4_kernel<false, true><<<ctas, 32 * warps, bytes, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count);
		}
		else if (smem && !textured) ris_ltc3_kernel<true, false><<<ctas, 32 * warps, bytes, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count);
		else if (!textured) ris_ltc3_kernel<false, false><<<ctas, 32 * warps, bytes, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count);
		else if (smem) ris_ltc3_kernel<true, true><<<ctas, 32 * warps, bytes, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count);
		else ris_ltc3_kernel<false, true><<<ctas, 32 * warps, bytes, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count);
		CU(cudaEventRecord(between, stream));
		{
			// phase-synchronous CTAs (shade_fast.cuh), two resident per SM, each walking over 8x4-pixel tiles
			const uint32_t threads = d->winner_threads, per_cta = threads / 32;
			uint32_t wctas = 2u * (uint32_t) d->sm_count;
			if (wctas * per_cta > tile_count) wctas = (tile_count + per_cta - 1) / per_cta;
			if (quads) winner_kernel<384, 768, 4><<<std::min(2u * (uint32_t) d->sm_count, (tile_count + 11u) / 12u), 384, 0, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count, v.min_light_vertices);
			else if (d->winner_cr) { if (rl_launch_winner_cr(d->view, f, d->stripes, px, tiles_x, tile_count, std::min(2u * (uint32_t) d->sm_count, (tile_count + 11u) / 12u), stream)) return 1; }
			else if (threads == 256) winner_kernel<256, 512><<<wctas, 256, 0, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count, 3u);
			else if (threads == 320) winner_kernel<320, 640><<<wctas, 320, 0, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count, 3u);
			else winner_kernel<384, 768><<<wctas, 384, 0, stream>>>(d->view, f, d->stripes, px, tiles_x, tile_count, 3u);
		}
		d->launches += 1;
	}
	else {
		if (rl_launch_generic_shade(d->view, f, v, d->stripes, px, grid, defer, stream)) return 1;
		CU(cudaEventRecord(between, stream));
	}
	return 0;
}

extern "C" int risltc_cuda_render_frames(risltc_device_t* d, const void* blocks, uint32_t frame_count, uint32_t first_accum_num) {
	if (use(d)) return 1;
	if (!blocks || frame_count == 0) return fail("render_frames: no constants", nullptr);
	if (!d->nodes || !d->materials || !d->lights || !d->ltc_rgba) return fail("render_frames: scene, materials, lights and LTC tables must be uploaded first", nullptr);
	if (!d->px.pixel_count) return fail("render_frames: call resize first", nullptr);
	if (d->view.light_stride4 != 3 + d->variant.max_light_vertices) return fail("render_frames: light buffer stride does not match the variant's max_light_vertices", nullptr);
	dim3 grid((d->width + 15) / 16, (d->stripes.owned_rows + 7) / 8);
	CU(cudaMemsetAsync(d->px.counters, 0, 8 * sizeof(unsigned long long), d->stream));
	CU(cudaMemsetAsync(d->px.ticket, 0, 4 * sizeof(unsigned int), d->stream));
	if (d->set2_ready) CU(cudaMemsetAsync(d->px2.ticket, 0, 4 * sizeof(unsigned int), d->stream));   // a failed call must not leave tickets behind
	while (d->frame_events.size() < RL_FRAME_EVENTS * (size_t) frame_count) { cudaEvent_t e; CU(cudaEventCreate(&e)); d->frame_events.push_back(e); }
	d->timed_frames = frame_count;
	CU(cudaEventRecord(d->ev[0], d->stream));
	const bool overlap = d->overlap && frame_count >= 2;
	if (overlap) {
		if (ensure_second_set(d)) return 1;
		d->px2.accum = d->px.accum;
		CU(cudaEventRecord(d->ev_fork, d->stream));
		CU(cudaStreamWaitEvent(d->stream2, d->ev_fork, 0));
	}
	uint32_t alternate = 0;
	bool used_second = false;
	for (uint32_t i = 0; i != frame_count; ++i) {
		FrameUniforms f;
		unpack_constants(f, (const unsigned char*) blocks + 256 * (size_t) i, first_accum_num + i);
		if (f.width != d->width || f.height != d->height) return fail("render_frames: viewport in the constants differs from resize()", nullptr);
		// while the G-buffer implementations are being timed the frames stay on one stream; afterwards they alternate
		const uint32_t set = (overlap && d->gbuffer_tune == 3 && d->trace_tune == RL_TRACE_DECIDED) ? (alternate++ & 1u) : 0u;
		const PixelBuffers& px = set ? d->px2 : d->px;
		const RasterBuffers& raster = set ? d->raster2 : d->raster;
		cudaStream_t stream = set ? d->stream2 : d->stream;
		used_second |= set != 0;
		d->last_set = set;
		cudaEvent_t* fe = &d->frame_events[RL_FRAME_EVENTS * (size_t) i];
		CU(cudaEventRecord(fe[0], stream));
		uint32_t kind = d->gbuffer_kind;
		if (d->gbuffer_tune < 3) {
			if (d->view.triangle_count > RL_RASTER_MAX_ITEMS) { d->gbuffer_kind = kind = 0; d->gbuffer_tune = 3; }   // beyond the rasteriser's queue
			else if (d->gbuffer_tune == 2) {
				float raster_ms = 0.0f, bvh_ms = 0.0f;
				CU(cudaEventSynchronize(d->tune_ev[3]));
				CU(cudaEventElapsedTime(&raster_ms, d->tune_ev[0], d->tune_ev[1]));
				CU(cudaEventElapsedTime(&bvh_ms, d->tune_ev[2], d->tune_ev[3]));
				d->gbuffer_kind = kind = (raster_ms <= bvh_ms) ? 1u : 0u;
				d->gbuffer_tune = 3;
			}
			else kind = (d->gbuffer_tune == 0) ? 1u : 0u;
		}
		if (d->gbuffer_tune < 2) CU(cudaEventRecord(d->tune_ev[2 * d->gbuffer_tune], stream));
		if (kind == 1) {
			// (1) every triangle finds its pixels and competes for them with atomicMin on {t, index} (raster.cuh)
			raster_setup_kernel<<<(d->view.triangle_count + 127) / 128, 128, 0, stream>>>(d->view, f, d->stripes, raster);
			raster_tiles_kernel<<<d->sm_count * 8, 128, 0, stream>>>(d->view, f, d->stripes, raster);
			raster_resolve_kernel<<<((px.pixel_count + 3) / 4 + 255) / 256, 256, 0, stream>>>(raster.zbuf, px.visibility, px.pixel_count, raster.counter);
			d->launches += 2;
		}
		else gbuffer_kernel<<<grid, 128, 0, stream>>>(d->view, f, d->stripes, px);
		if (d->gbuffer_tune < 2) { CU(cudaEventRecord(d->tune_ev[2 * d->gbuffer_tune + 1], stream)); d->gbuffer_tune++; }
		CU(cudaEventRecord(fe[1], stream));
		if (launch_shade(d, grid, f, px, stream, fe[2])) return 1;
		CU(cudaEventRecord(fe[3], stream));
		const bool traced = d->precision == RISLTC_PRECISION_FAST && deferred_rays(d->variant);
		if (traced) {
			// (3) persistent any-hit traversal over all ray slots
			const uint32_t ray_count = px.pixel_count * d->variant.light_samples * d->variant.sample_count * 2u;
			int resident = (d->trace_kind == 8) ? d->trace4p_resident : d->trace4_resident;
			// candidates: {triangle-track threshold, nearest child first}; the first is the default and wins ties (3 %)
			static const uint32_t vote_candidates[RL_TRACE_CANDIDATES] = { 8u, 1u, 8u };
			static const bool order_candidates[RL_TRACE_CANDIDATES] = { true, true, false };
			if (d->trace_tune == RL_TRACE_TRIALS) {
				CU(cudaEventSynchronize(d->trace_tune_ev[2 * RL_TRACE_TRIALS - 1]));
				float best_ms = 0.0f; uint32_t best = 0;
				for (uint32_t c = 0; c != RL_TRACE_CANDIDATES; ++c) {
					float first = 0.0f, second = 0.0f;
					CU(cudaEventElapsedTime(&first, d->trace_tune_ev[2 * c], d->trace_tune_ev[2 * c + 1]));
					CU(cudaEventElapsedTime(&second, d->trace_tune_ev[2 * (c + RL_TRACE_CANDIDATES)], d->trace_tune_ev[2 * (c + RL_TRACE_CANDIDATES) + 1]));
					const float ms = std::min(first, second);
					if (c == 0 || ms < 0.97f * best_ms) { best_ms = ms; best = c; }
				}
				d->tri_vote = vote_candidates[best]; d->trace_ordered = order_candidates[best];
				d->trace_tune = RL_TRACE_DECIDED;
			}
			const uint32_t tuning = d->trace_tune;
			if (tuning < RL_TRACE_TRIALS) { d->tri_vote = vote_candidates[tuning % RL_TRACE_CANDIDATES]; d->trace_ordered = order_candidates[tuning % RL_TRACE_CANDIDATES]; CU(cudaEventRecord(d->trace_tune_ev[2 * tuning], stream)); }
			const bool ordered = d->trace_ordered || d->trace_kind != 8;
			if (!ordered) resident = d->trace4pu_resident;   // the unordered variant needs fewer registers: one more CTA per SM
			const int per_sm = (d->trace_ctas_per_sm > 0 && d->trace_ctas_per_sm < resident) ? d->trace_ctas_per_sm : resident;
			if (d->trace_kind == 8 && d->count_traversal && ordered) trace4p_kernel<true, true><<<d->sm_count * per_sm, 128, 0, stream>>>(d->view, px, ray_count / 2u, d->tri_vote, d->refill, 0x3F800000u);
			else if (d->trace_kind == 8 && d->count_traversal) trace4p_kernel<true, false><<<d->sm_count * per_sm, 128, 0, stream>>>(d->view, px, ray_count / 2u, d->tri_vote, d->refill, 0x3F800000u);
			else if (d->trace_kind == 8 && ordered) trace4p_kernel<false, true><<<d->sm_count * per_sm, 128, 0, stream>>>(d->view, px, ray_count / 2u, d->tri_vote, d->refill, 0x3F800000u);
			else if (d->trace_kind == 8) trace4p_kernel<false, false><<<d->sm_count * per_sm, 128, 0, stream>>>(d->view, px, ray_count / 2u, d->tri_vote, d->refill, 0x3F800000u);
			else if (d->trace_kind == 4 && d->count_traversal) trace4_kernel<true><<<d->sm_count * per_sm, 128, 0, stream>>>(d->view, px, ray_count, d->tri_vote, d->refill, 0x3F800000u);
			else if (d->trace_kind == 4) trace4_kernel<false><<<d->sm_count * per_sm, 128, 0, stream>>>(d->view, px, ray_count, d->tri_vote, d->refill, 0x3F800000u);
			else trace_kernel<<<d->sm_count * d->trace_resident, 128, 0, stream>>>(d->view, px, ray_count, d->tri_vote);
			if (tuning < RL_TRACE_TRIALS) { CU(cudaEventRecord(d->trace_tune_ev[2 * tuning + 1], stream)); d->trace_tune = tuning + 1; }
			d->launches += 1;
		}
		CU(cudaEventRecord(fe[4], stream));
		// (4) MIS sum + accumulation: the running mean takes the frames in order, whichever stream they were rendered on
		if (overlap && i != 0) CU(cudaStreamWaitEvent(stream, d->ev_resolved[(i - 1) & 1u], 0));
		if (traced) resolve_kernel<true><<<grid, 128, 0, stream>>>(d->view, f, d->variant, d->stripes, px);
		else resolve_kernel<false><<<grid, 128, 0, stream>>>(d->view, f, d->variant, d->stripes, px);
		if (overlap) CU(cudaEventRecord(d->ev_resolved[i & 1u], stream));
		CU(cudaEventRecord(fe[5], stream));
		d->launches += 3;
	}
	if (used_second) {
		CU(cudaEventRecord(d->ev_join, d->stream2));
		CU(cudaStreamWaitEvent(d->stream, d->ev_join, 0));
	}
	CU(cudaEventRecord(d->ev[4], d->stream));
	CU(cudaGetLastError());
	d->timed = true;
	return 0;
}

extern "C" int risltc_cuda_render_frame(risltc_device_t* d, const void* block, uint32_t accum_num) {
	return risltc_cuda_render_frames(d, block, 1, accum_num);
}

extern "C" int risltc_cuda_synchronize(risltc_device_t* d) {
	if (use(d)) return 1;
	CU(cudaStreamSynchronize(d->stream));
	return 0;
}

extern "C" int risltc_cuda_read_accum(risltc_device_t* d, float* rgba) {
	if (use(d)) return 1;
	if (!rgba || !d->px.accum) return fail("read_accum: nothing to read", nullptr);
	CU(cudaMemcpyAsync(rgba, d->px.accum, (size_t) d->px.pixel_count * sizeof(float4), cudaMemcpyDeviceToHost, d->stream));
	CU(cudaStreamSynchronize(d->stream));
	return 0;
}

// copy_pass.frag.glsl:28-58 followed by the UNORM8 framebuffer write. frame_bits 0: linear -> sRGB (srgb_utility.glsl:20-34),
// 1 / 2: the low / high byte of every channel's half-float bits (packHalf2x16 rounds to nearest even)
__global__ void copy_pass_kernel(const float4* accum, uint8_t* rgb8, uint32_t pixel_count, uint32_t frame_bits) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= pixel_count) return;
	const float4 c = accum[i];
	const float channel[3] = { c.x, c.y, c.z };
	#pragma unroll
	for (int k = 0; k != 3; ++k) {
		uint32_t byte;
		if (frame_bits != 0u) {
			const uint32_t half_bits = __half_as_ushort(__float2half_rn(channel[k]));
			byte = (frame_bits == 1u) ? (half_bits & 0xFFu) : (half_bits >> 8);
		}
		else {
			const float linear = fminf(fmaxf(channel[k], 0.0f), 1.0f);
			// pow() correctly rounded, like the oracle defines GLSL's transcendental functions (common.cuh)
			const float srgb = (linear <= 0.0031308f) ? (12.92f * linear) : (1.055f * (float) pow((double) linear, (double) (1.0f / 2.4f)) - 0.055f);
			byte = (uint32_t) (fminf(fmaxf(srgb, 0.0f), 1.0f) * 255.0f + 0.5f);
		}
		rgb8[3 * (size_t) i + k] = (uint8_t) byte;
	}
}

extern "C" int risltc_cuda_copy_pass(risltc_device_t* d, uint32_t frame_bits, uint8_t* rgb8) {
	if (use(d)) return 1;
	if (!rgb8 || !d->px.accum) return fail("copy_pass: nothing to read", nullptr);
	if (frame_bits > 2u) return fail("copy_pass: frame_bits is 0 (display), 1 (low half bits) or 2 (high half bits)", nullptr);
	uint8_t* staging = nullptr;
	const size_t bytes = 3 * (size_t) d->px.pixel_count;
	CU(cudaMalloc(&staging, bytes));
	copy_pass_kernel<<<(d->px.pixel_count + 255) / 256, 256, 0, d->stream>>>(d->px.accum, staging, d->px.pixel_count, frame_bits);
	cudaError_t e = cudaMemcpyAsync(rgb8, staging, bytes, cudaMemcpyDeviceToHost, d->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
	cudaFree(staging);
	if (e != cudaSuccess) return fail("copy_pass", cudaGetErrorString(e));
	return 0;
}

extern "C" int risltc_cuda_read_visibility(risltc_device_t* d, uint32_t* ids) {
	if (use(d)) return 1;
	if (!ids || !d->px.visibility) return fail("read_visibility: nothing to read", nullptr);
	CU(cudaMemcpyAsync(ids, (d->last_set ? d->px2 : d->px).visibility, (size_t) d->px.pixel_count * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->stream));
	CU(cudaStreamSynchronize(d->stream));
	return 0;
}

extern "C" float risltc_cuda_last_frame_ms(risltc_device_t* d) {
	float ms[4];
	if (risltc_cuda_last_kernel_ms(d, ms)) return -1.0f;
	return ms[3];
}

extern "C" int risltc_cuda_last_pass_ms(risltc_device_t* d, float ms[6]) {
	if (use(d)) return 1;
	if (!d->timed) return fail("last_pass_ms: no frame rendered yet", nullptr);
	CU(cudaEventSynchronize(d->ev[4]));
	for (int k = 0; k != 5; ++k) ms[k] = 0.0f;
	for (uint32_t i = 0; i != d->timed_frames; ++i) {
		const cudaEvent_t* fe = &d->frame_events[RL_FRAME_EVENTS * (size_t) i];
		for (int k = 0; k != 5; ++k) { float t = 0.0f; CU(cudaEventElapsedTime(&t, fe[k], fe[k + 1])); ms[k] += t; }
	}
	CU(cudaEventElapsedTime(&ms[5], d->ev[0], d->ev[4]));
	return 0;
}

extern "C" int risltc_cuda_last_kernel_ms(risltc_device_t* d, float ms[4]) {
	float pass[6];
	if (risltc_cuda_last_pass_ms(d, pass)) return 1;
	ms[0] = pass[0]; ms[1] = pass[1] + pass[2]; ms[2] = pass[3] + pass[4]; ms[3] = pass[5];
	return 0;
}

extern "C" int risltc_cuda_counters(risltc_device_t* d, uint64_t counters[4]) {
	if (use(d)) return 1;
	unsigned long long h[4];
	CU(cudaMemcpyAsync(h, d->px.counters, sizeof(h), cudaMemcpyDeviceToHost, d->stream));
	CU(cudaStreamSynchronize(d->stream));
	counters[0] = h[0]; counters[1] = h[1]; counters[2] = d->launches; counters[3] = h[3];
	return 0;
}

extern "C" int risltc_cuda_traversal_counters(risltc_device_t* d, uint32_t enable, uint64_t counters[4]) {
	if (use(d)) return 1;
	CU(cudaStreamSynchronize(d->stream));
	if (counters) {
		unsigned long long h[4];
		CU(cudaMemcpy(h, d->px.counters + 4, sizeof(h), cudaMemcpyDeviceToHost));
		for (int i = 0; i != 4; ++i) counters[i] = h[i];
	}
	d->count_traversal = enable != 0;
	return 0;
}

// Host-only self check of the acceleration structures (no device needed): builds the binary tree and its 4-wide, 8-bit
// collapse for a triangle soup and verifies the invariants the traversal kernels rely on.
// Invariants of a pair of acceleration structures over `vertices` (9 floats per triangle): see risltc_cuda_check_bvh in the header
static void check_trees(const float* vertices, uint64_t T, const std::vector<BvhNodeHost>& nodes, const std::vector<uint32_t>& order,
	const std::vector<Qbvh4NodeHost>& n4, uint32_t depth, uint64_t report[6])
{
	// every triangle appears in exactly one leaf slot
	std::vector<uint32_t> seen(T, 0);
	for (uint32_t t : order) if (t < T) seen[t]++;
	uint64_t bad_order = 0;
	for (uint64_t t = 0; t != T; ++t) bad_order += seen[t] != 1;
	// leaf boxes of the binary tree by reference; only nodes reachable from the root count (the device builder leaves unused slots)
	struct LeafBox { float lo[3], hi[3]; };
	std::vector<std::pair<int, LeafBox>> leaves;
	uint64_t bad_binary = 0, reachable = 0;
	auto check_leaf = [&](int ref, const float* lo, const float* hi) {
		const uint32_t r = ~(uint32_t) ref, first = r >> 4, count = (r & 15u) + 1u;
		for (uint32_t s2 = first; s2 != first + count; ++s2) {
			if (s2 >= order.size() || order[s2] >= T) { ++bad_binary; continue; }
			for (int v = 0; v != 3; ++v)
				for (int k = 0; k != 3; ++k) {
					const float x = vertices[9 * (size_t) order[s2] + 3 * v + k];
					if (!(x >= lo[k] && x <= hi[k])) ++bad_binary;
				}
		}
		LeafBox b; memcpy(b.lo, lo, 12); memcpy(b.hi, hi, 12);
		leaves.push_back({ ref, b });
	};
	std::vector<int> todo(1, 0);
	std::vector<uint8_t> visited(nodes.size(), 0);
	while (!todo.empty()) {
		const int i = todo.back(); todo.pop_back();
		if (i < 0 || (size_t) i >= nodes.size() || visited[i]) { ++bad_binary; continue; }
		visited[i] = 1; ++reachable;
		const BvhNodeHost& n = nodes[i];
		// an inner child's box must contain both boxes stored in that child
		auto check_inner = [&](int child, const float* lo, const float* hi) {
			if ((size_t) child >= nodes.size()) { ++bad_binary; return; }
			const BvhNodeHost& c = nodes[child];
			for (int k = 0; k != 3; ++k) {
				if (!(lo[k] <= c.left_lo[k] + 2e-5f * fabsf(c.left_lo[k]) + 1e-6f && hi[k] >= c.left_hi[k] - 2e-5f * fabsf(c.left_hi[k]) - 1e-6f)) ++bad_binary;
				if (c.right_lo[0] <= c.right_hi[0] && !(lo[k] <= c.right_lo[k] + 2e-5f * fabsf(c.right_lo[k]) + 1e-6f && hi[k] >= c.right_hi[k] - 2e-5f * fabsf(c.right_hi[k]) - 1e-6f)) ++bad_binary;
			}
			todo.push_back(child);
		};
		if (n.left < 0) check_leaf(n.left, n.left_lo, n.left_hi); else check_inner(n.left, n.left_lo, n.left_hi);
		if (n.right < 0) { if (n.right_lo[0] <= n.right_hi[0]) check_leaf(n.right, n.right_lo, n.right_hi); }
		else check_inner(n.right, n.right_lo, n.right_hi);
	}
	// the leaves partition the slots
	std::vector<uint32_t> slot_seen(order.size(), 0);
	for (const auto& l : leaves) { const uint32_t r = ~(uint32_t) l.first; for (uint32_t s2 = r >> 4; s2 != (r >> 4) + (r & 15u) + 1u && s2 < order.size(); ++s2) slot_seen[s2]++; }
	for (uint32_t c : slot_seen) bad_binary += c != 1;
	std::sort(leaves.begin(), leaves.end(), [](const std::pair<int, LeafBox>& a, const std::pair<int, LeafBox>& b) { return a.first < b.first; });
	// 4-wide tree: every leaf reference appears once with a quantised box that contains the binary tree's box; every inner node is referenced once
	uint64_t bad_wide = 0, children = 0;
	std::vector<uint32_t> node_refs(n4.size(), 0), leaf_refs(leaves.size(), 0);
	const int scale_word[3] = { 3, 10, 11 };
	for (const Qbvh4NodeHost& q : n4) {
		float origin[3]; memcpy(origin, q.w, 12);
		for (int c = 0; c != 4; ++c) {
			const int ref = (int) q.w[12 + c];
			if (ref == 0x7FFFFFFF) continue;
			++children;
			if (ref >= 0) { if ((size_t) ref < n4.size()) node_refs[ref]++; else ++bad_wide; continue; }
			auto it = std::lower_bound(leaves.begin(), leaves.end(), ref, [](const std::pair<int, LeafBox>& a, int r) { return a.first < r; });
			if (it == leaves.end() || it->first != ref) { ++bad_wide; continue; }
			leaf_refs[it - leaves.begin()]++;
			for (int k = 0; k != 3; ++k) {
				float scale; memcpy(&scale, &q.w[scale_word[k]], 4);
				const double step = (double) scale / 32768.0;
				const double lo = origin[k] + ((q.w[4 + k] >> (8 * c)) & 0xFFu) * step, hi = origin[k] + ((q.w[7 + k] >> (8 * c)) & 0xFFu) * step;
				if (!(lo <= it->second.lo[k] && hi >= it->second.hi[k])) ++bad_wide;
			}
		}
	}
	for (size_t i = 1; i < n4.size(); ++i) bad_wide += node_refs[i] != 1;
	for (uint32_t r : leaf_refs) bad_wide += r != 1;
	report[0] = bad_order; report[1] = bad_binary; report[2] = bad_wide; report[3] = reachable; report[4] = n4.size(); report[5] = ((uint64_t) depth << 32) | (uint64_t) (children * 100 / (n4.size() ? n4.size() : 1));
}

extern "C" int risltc_cuda_check_bvh(const float* vertices, uint64_t T, uint32_t max_leaf, uint64_t report[6]) {
	if (!vertices || T == 0 || !report) return fail("check_bvh: no triangles", nullptr);
	std::vector<BvhNodeHost> nodes; std::vector<uint32_t> order;
	build_bvh(vertices, T, nodes, order, max_leaf);
	std::vector<Qbvh4NodeHost> n4;
	const uint32_t depth = build_qbvh4(nodes, n4);
	check_trees(vertices, T, nodes, order, n4, depth, report);
	return 0;
}

// FNV-1a over the host builder's binary tree, triangle order and 4-wide tree: equal for equal trees (host-only)
extern "C" int risltc_cuda_bvh_checksum(const float* vertices, uint64_t T, uint32_t max_leaf, uint64_t* checksum) {
	if (!vertices || T == 0 || !checksum) return fail("bvh_checksum: no triangles", nullptr);
	std::vector<BvhNodeHost> nodes; std::vector<uint32_t> order;
	build_bvh(vertices, T, nodes, order, max_leaf);
	std::vector<Qbvh4NodeHost> n4;
	build_qbvh4(nodes, n4);
	uint64_t h = 1469598103934665603ull;
	auto mix = [&](const void* data, size_t bytes) { const unsigned char* p = (const unsigned char*) data; for (size_t i = 0; i != bytes; ++i) { h ^= p[i]; h *= 1099511628211ull; } };
	mix(nodes.data(), nodes.size() * sizeof(BvhNodeHost)); mix(order.data(), order.size() * sizeof(uint32_t)); mix(n4.data(), n4.size() * sizeof(Qbvh4NodeHost));
	*checksum = h;
	return 0;
}

// The same invariants for the structures upload_scene left on the device (whichever builder made them)
extern "C" int risltc_cuda_check_scene_bvh(risltc_device_t* d, uint64_t report[6]) {
	if (use(d)) return 1;
	if (!d->nodes || !d->nodes4 || !d->tris || !report) return fail("check_scene_bvh: upload_scene first", nullptr);
	const uint64_t T = d->view.triangle_count;
	std::vector<BvhNode> dn(d->node_count); std::vector<BvhTri> dt(T); std::vector<Qbvh4NodeHost> n4(d->node4_count); std::vector<uint32_t> qp(6 * T);
	CU(cudaMemcpy(dn.data(), d->nodes, dn.size() * sizeof(BvhNode), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(dt.data(), d->tris, dt.size() * sizeof(BvhTri), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(n4.data(), d->nodes4, n4.size() * sizeof(Qbvh4Node), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(qp.data(), d->positions, qp.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	std::vector<float> verts;
	dequantize_mesh_for_bvh(qp.data(), T, d->dequant_factor, d->dequant_summand, verts);
	std::vector<BvhNodeHost> nodes(dn.size()); std::vector<uint32_t> order(T);
	uint64_t bad_records = 0;
	for (size_t i = 0; i != dn.size(); ++i) {
		BvhNodeHost& n = nodes[i];
		n.left_lo[0] = dn[i].a.x; n.left_lo[1] = dn[i].a.y; n.left_lo[2] = dn[i].a.z; n.left_hi[0] = dn[i].a.w; n.left_hi[1] = dn[i].b.x; n.left_hi[2] = dn[i].b.y;
		n.right_lo[0] = dn[i].b.z; n.right_lo[1] = dn[i].b.w; n.right_lo[2] = dn[i].c.x; n.right_hi[0] = dn[i].c.y; n.right_hi[1] = dn[i].c.z; n.right_hi[2] = dn[i].c.w;
		n.left = dn[i].d.x; n.right = dn[i].d.y;
	}
	for (uint64_t slot = 0; slot != T; ++slot) {
		uint32_t id; memcpy(&id, &dt[slot].v0.w, 4);
		const uint32_t t = id & 0x7FFFFFFFu;
		order[slot] = t;
		if (t >= T) { ++bad_records; continue; }
		// the triangle record holds v0 and the two edges of triangle t, rounded as the host builder rounds them
		const float* v = verts.data() + 9 * (size_t) t;
		volatile float e1x = v[3] - v[0], e1y = v[4] - v[1], e1z = v[5] - v[2], e2x = v[6] - v[0], e2y = v[7] - v[1], e2z = v[8] - v[2];
		const float want[9] = { v[0], v[1], v[2], e1x, e1y, e1z, e2x, e2y, e2z };
		const float got[9] = { dt[slot].v0.x, dt[slot].v0.y, dt[slot].v0.z, dt[slot].e1.x, dt[slot].e1.y, dt[slot].e1.z, dt[slot].e2.x, dt[slot].e2.y, dt[slot].e2.z };
		bad_records += memcmp(want, got, sizeof(want)) != 0;
		bad_records += (id >> 31) != (qp[6 * (size_t) t + 1] >> 31);
	}
	check_trees(verts.data(), T, nodes, order, n4, (uint32_t) d->bvh_stats[7] & 0xFFFFu, report);
	report[0] += bad_records;
	return 0;
}

extern "C" int risltc_cuda_set_bvh_builder(risltc_device_t* d, uint32_t builder) {
	if (use(d)) return 1;
	if (builder > RISLTC_BVH_BUILDER_DEVICE_RADIX) return fail("set_bvh_builder: unknown builder", nullptr);
	d->bvh_builder = builder;
	return 0;
}

extern "C" int risltc_cuda_bvh_stats(risltc_device_t* d, double stats[8]) {
	if (use(d)) return 1;
	if (!stats || !d->nodes) return fail("bvh_stats: upload_scene first", nullptr);
	memcpy(stats, d->bvh_stats, sizeof(d->bvh_stats));
	return 0;
}

extern "C" void* risltc_cuda_stream(risltc_device_t* d) { return d ? (void*) d->stream : nullptr; }

// ---- known-answer entry point of the shadow-ray kernels (the other risltc_cuda_kat_* live in kat.cu)
template <typename T>
struct DeviceArray {
	T* p = nullptr; size_t n = 0;
	int init(const T* host, size_t count) {
		n = count;
		if (cudaMalloc(&p, sizeof(T) * (count ? count : 1)) != cudaSuccess) return 1;
		if (host && cudaMemcpy(p, host, sizeof(T) * count, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
		return 0;
	}
	int fetch(T* host) { return cudaMemcpy(host, p, sizeof(T) * n, cudaMemcpyDeviceToHost) != cudaSuccess; }
	~DeviceArray() { cudaFree(p); }
};
#define KAT_GRID(count) ((count) + 127) / 128, 128

// The production shadow-ray kernels (kind 8: trace4p_kernel, 4: trace4_kernel, 2: trace_kernel) on an array of rays with
// t_min = 1e-3: the rays are laid out as one ray slot of `count` pixels, exactly what the shading kernels leave behind; for
// kind 8 as two ray slots of count / 2 pixels (ray i and ray i + count / 2 must share their origin, count must be even).
__global__ void kat_trace_fill_kernel(const float* rays, float4* origin, float4* ray_a, float4* ray_b, uint32_t count) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const float* r = rays + 8 * (size_t) i;
	origin[i] = make_float4(r[0], r[1], r[2], 0.0f);
	ray_a[i] = make_float4(r[4], r[5], r[6], r[7]);
	ray_b[i] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
}
__global__ void kat_trace_read_kernel(const float4* ray_b, uint32_t* hits, uint32_t count) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) hits[i] = (ray_b[i].w == 2.0f) ? 1u : 0u;
}
extern "C" int risltc_cuda_kat_trace(risltc_device_t* d, const float* rays, uint32_t* hits, uint32_t count, uint32_t kind) {
	if (use(d)) return 1;
	if (!d->nodes || !d->nodes4) return fail("kat_trace: upload_scene first", nullptr);
	if (kind != 2 && kind != 4 && kind != 8) return fail("kat_trace: kind must be 2, 4 or 8", nullptr);
	if (kind == 8 && (count & 1u)) return fail("kat_trace: kind 8 takes an even number of rays (two slots of count / 2 pixels)", nullptr);
	DeviceArray<float> r; DeviceArray<uint32_t> h; DeviceArray<float4> og, ra, rb; DeviceArray<unsigned int> ticket;
	if (r.init(rays, (size_t) count * 8) || h.init(nullptr, count) || og.init(nullptr, count) || ra.init(nullptr, count) || rb.init(nullptr, count) || ticket.init(nullptr, 4))
		return fail("kat_trace: allocation failed", nullptr);
	CU(cudaMemset(ticket.p, 0, 4 * sizeof(unsigned int)));
	PixelBuffers px = {};
	px.origin = og.p; px.ray_a = ra.p; px.ray_b = rb.p; px.ticket = ticket.p; px.pixel_count = (kind == 8) ? count / 2u : count;
	kat_trace_fill_kernel<<<KAT_GRID(count)>>>(r.p, og.p, ra.p, rb.p, count);
	if (kind == 8) trace4p_kernel<false><<<d->sm_count * d->trace4p_resident, 128>>>(d->view, px, count / 2u, d->tri_vote, d->refill, 0x3F800000u);
	else if (kind == 4) trace4_kernel<false><<<d->sm_count * d->trace4_resident, 128>>>(d->view, px, count, d->tri_vote, d->refill, 0x3F800000u);
	else trace_kernel<<<d->sm_count * d->trace_resident, 128>>>(d->view, px, count, d->tri_vote);
	kat_trace_read_kernel<<<KAT_GRID(count)>>>(rb.p, h.p, count);
	CU(cudaDeviceSynchronize());
	if (h.fetch(hits)) return fail("kat_trace: read-back failed", nullptr);
	return 0;
}

