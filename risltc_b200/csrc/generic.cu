// generic.cu -- the generic, exactly rounded shading kernel (kernels.cuh: shade_kernel) instantiated for every light
// vertex count the reference supports (MAX_POLYGONAL_LIGHT_VERTEX_COUNT 3..7, main.c:191-204, clip table
// polygon_clipping.glsl:161-216). Serves every variant of the comparison matrix (experiment_list.c:316-396) and
// RISLTC_PRECISION_EXACT; a translation unit of its own so that it compiles in parallel with api.cu.
// a namespace of its own: the host-side stubs of the shared __device__ functions must not collide with api.cu's
#define RL_NS generic
#define RL_CR_LIBM 1   // transcendental functions correctly rounded, like the oracle (common.cuh)
#include "internal.h"
#include "kernels.cuh"

using namespace RL_NS;

template <int V>
static void launch(const SceneView& s, const FrameUniforms& f, const Variant& v, const Stripes& st, const PixelBuffers& px, dim3 grid, bool defer, cudaStream_t stream) {
	if (defer) shade_kernel<V, true><<<grid, 128, 0, stream>>>(s, f, v, st, px);
	else shade_kernel<V, false><<<grid, 128, 0, stream>>>(s, f, v, st, px);
}

int rl_launch_generic_shade(const SceneView& s, const FrameUniforms& f, const Variant& v, const Stripes& st, const PixelBuffers& px,
	dim3 grid, bool defer, cudaStream_t stream)
{
	switch (v.max_light_vertices) {
	case 3: launch<3>(s, f, v, st, px, grid, defer, stream); break;
	case 4: launch<4>(s, f, v, st, px, grid, defer, stream); break;
	case 5: launch<5>(s, f, v, st, px, grid, defer, stream); break;
	case 6: launch<6>(s, f, v, st, px, grid, defer, stream); break;
	case 7: launch<7>(s, f, v, st, px, grid, defer, stream); break;
	default: return rl_fail("shade: lights have 3 to 7 vertices", nullptr);
	}
	return 0;
}
