// launch.h -- entry points of the `fast` translation unit (fast.cu) for api.cu.
#pragma once
#include "common.cuh"

namespace fast {
// uploads the clip rotation table into this translation unit's constant memory and sizes the persistent grid
cudaError_t initialize(int device_ordinal);
// One frame: (1) G-buffer, (2) fused RIS + shading, (3)+(4) shadow rays + MIS sum + accumulation.
// Events (may be null) are recorded before (1) and after each kernel. Returns the number of kernels launched.
int launch_frame(cudaStream_t stream, const SceneView& view, const FrameUniforms& f, const Variant& variant, const Stripes& stripes,
	const PixelBuffers& px, cudaEvent_t* events4);
}
