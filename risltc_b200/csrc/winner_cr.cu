// winner_cr.cu -- winner_kernel (shade_fast.cuh) compiled a second time with correctly rounded transcendental functions
// (RL_CR_LIBM, common.cuh): the estimator's projected solid angles are differences of atan terms, so 1-2 ulp of atanf move
// some samples by more than the parity tolerance. RISLTC_WINNER=cr (or risltc_cuda_set_precision's FAST_CR) selects it.
// A namespace of its own: the host-side stubs of the shared __device__ functions must not collide with api.cu's.
#define RL_NS winner_cr
#define RL_CR_LIBM 1
#define RL_PSA_ATTR __forceinline__   // the production kernels call the PSA functions once per loop body (shading.cuh)
#include "internal.h"
#include "shade_fast.cuh"

using namespace RL_NS;

int rl_launch_winner_cr(const SceneView& s, const FrameUniforms& f, const Stripes& st, const PixelBuffers& px, uint32_t tiles_x, uint32_t tile_count,
	uint32_t ctas, cudaStream_t stream)
{
	winner_kernel<384, 768><<<ctas, 384, 0, stream>>>(s, f, st, px, tiles_x, tile_count, 3u);
	return 0;
}
