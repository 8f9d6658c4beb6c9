// internal.h -- what the translation units of librisltc_cuda.so share besides the kernels' headers. The device object
// itself stays private to api.cu; the other units get at it through these accessors.
//   api.cu      C ABI, device object, frame loop, the production kernels of the default estimator
//   generic.cu  the exactly rounded generic shading kernel for every variant and 3..7 vertex lights (kernels.cuh)
//   kat.cu      known-answer entry points
#pragma once
#include "../../include/risltc_cuda.h"
#include "common.cuh"

int rl_fail(const char* what, const char* detail);
int rl_use(risltc_device_t* d);
const SceneView& rl_view(const risltc_device_t* d);
int rl_sm_count(const risltc_device_t* d);
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return rl_fail(#call, cudaGetErrorString(e_)); } while (0)

// generic.cu: shade_kernel<V, DEFER> for V = variant.max_light_vertices (3..7); defer = rays are recorded for kernel (3)
int rl_launch_generic_shade(const SceneView& s, const FrameUniforms& f, const Variant& v, const Stripes& st, const PixelBuffers& px,
	dim3 grid, bool defer, cudaStream_t stream);

// winner_cr.cu: winner_kernel<384, 768> with correctly rounded atan / acos / sin / cos
int rl_launch_winner_cr(const SceneView& s, const FrameUniforms& f, const Stripes& st, const PixelBuffers& px, uint32_t tiles_x, uint32_t tile_count,
	uint32_t ctas, cudaStream_t stream);

// bvh_gpu.cu: the acceleration structures of upload_scene built on the device from the quantised positions already there.
// counts = {binary node slots, 4-wide nodes}, depths = {binary levels, 4-wide levels}, ms = {sort + hierarchy, boxes + records, collapse}
int rl_build_bvh_gpu(const uint2* positions, uint64_t triangle_count, const float factor[3], const float summand[3], uint32_t max_leaf, uint32_t ploc_radius,
	BvhNode** nodes, BvhTri** tris, Qbvh4Node** nodes4, uint64_t counts[2], uint32_t depths[2], float ms[3]);
