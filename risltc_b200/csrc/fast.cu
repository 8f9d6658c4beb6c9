// fast.cu -- the production kernel set: same device code as api.cu's `exact` set, compiled with fused
// multiply-adds, MUFU reciprocals / rsqrt and flush-to-zero, plus the specialised persistent shading
// kernel of shade_fast.cuh for the default estimator.
#define RL_NS fast
#define RL_FAST 1
#include "shade_fast.cuh"
#include "launch.h"
#include "clip_rotation_table.inc"

namespace fast {

static int g_sm_count = 148;
static int g_resident[2] = { 0, 0 };   // resident CTAs per SM of shade_ris_ltc3_kernel<false / true>

cudaError_t initialize(int device_ordinal) {
	cudaError_t e = cudaMemcpyToSymbol(c_clip_rotation, h_clip_rotation, sizeof(h_clip_rotation));
	if (e != cudaSuccess) return e;
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties(&prop, device_ordinal);
	if (e != cudaSuccess) return e;
	g_sm_count = prop.multiProcessorCount;
	e = cudaFuncSetAttribute(shade_ris_ltc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	return e;
}

// Largest light table staged in shared memory: 48 bytes per triangle light, keeping >= 2 CTAs per SM
static const uint32_t kMaxSmemLights = 2048;

int launch_frame(cudaStream_t stream, const SceneView& view, const FrameUniforms& f, const Variant& variant, const Stripes& stripes,
	const PixelBuffers& px, cudaEvent_t* ev)
{
	const uint32_t tiles_x = (f.width + 15) / 16, tiles_y = (stripes.owned_rows + 7) / 8;
	dim3 grid(tiles_x, tiles_y);
	const bool defer = variant.polygon_technique != TECH_TURK && variant.polygon_technique != TECH_BASELINE && variant.mis_heuristic != MIS_OPTIMAL;
	if (ev) cudaEventRecord(ev[0], stream);
	gbuffer_kernel<<<grid, 128, 0, stream>>>(view, f, stripes, px);
	if (ev) cudaEventRecord(ev[1], stream);
	const bool specialised = variant.light_sampling == 1u && variant.polygon_technique == TECH_LTC_CP && variant.mis_heuristic == MIS_OPTIMAL_CLAMPED
		&& variant.sample_count == 1u && variant.light_samples == 1u && variant.fast_atan == 0u
		&& variant.min_light_vertices == 3u && variant.max_light_vertices == 3u && view.lights_tri != nullptr;
	if (specialised) {
		const bool smem = view.light_count <= kMaxSmemLights;
		const size_t bytes = smem ? (size_t) view.light_count * 48 : 0;
		int& resident = g_resident[smem ? 1 : 0];
		if (smem) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, shade_ris_ltc3_kernel<true>, 128, bytes);
		else if (!resident) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, shade_ris_ltc3_kernel<false>, 128, 0);
		const uint32_t tile_count = tiles_x * tiles_y;
		uint32_t ctas = (uint32_t) (g_sm_count * (resident > 0 ? resident : 1));
		if (ctas > tile_count) ctas = tile_count;
		if (smem) shade_ris_ltc3_kernel<true><<<ctas, 128, bytes, stream>>>(view, f, stripes, px, tiles_x, tile_count);
		else shade_ris_ltc3_kernel<false><<<ctas, 128, 0, stream>>>(view, f, stripes, px, tiles_x, tile_count);
	}
	else if (variant.max_light_vertices == 3) {
		if (defer) shade_kernel<3, true><<<grid, 128, 0, stream>>>(view, f, variant, stripes, px);
		else shade_kernel<3, false><<<grid, 128, 0, stream>>>(view, f, variant, stripes, px);
	}
	else {
		if (defer) shade_kernel<4, true><<<grid, 128, 0, stream>>>(view, f, variant, stripes, px);
		else shade_kernel<4, false><<<grid, 128, 0, stream>>>(view, f, variant, stripes, px);
	}
	if (ev) cudaEventRecord(ev[2], stream);
	resolve_kernel<<<grid, 128, 0, stream>>>(view, f, variant, stripes, px);
	if (ev) cudaEventRecord(ev[3], stream);
	return 3;
}

}  // namespace fast
