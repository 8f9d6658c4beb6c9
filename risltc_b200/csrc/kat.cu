// kat.cu -- known-answer entry points: the device functions of the shading kernels run on arrays, one thread per item,
// so that tests can compare them with the oracle function by function (include/risltc_cuda.h, risltc_cuda_kat_*).
// risltc_cuda_kat_trace, which runs the production shadow-ray kernels, lives in api.cu next to them.
// a namespace of its own: the host-side stubs of the shared __device__ functions must not collide with api.cu's
#define RL_NS kat
#define RL_CR_LIBM 1   // transcendental functions correctly rounded, like the oracle (common.cuh)
#include "internal.h"
#include "shading.cuh"
#include "bvh.cuh"
#include <cstdio>

using namespace RL_NS;
#define use rl_use
#define fail rl_fail

template <int V>
__global__ void kat_clip_kernel(float* polygons, uint32_t* counts, uint32_t count, uint32_t min_vertices) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	float3 v[V + 1];
	for (int k = 0; k != V + 1; ++k) v[k] = mk3(polygons[i * 24 + 3 * k], polygons[i * 24 + 3 * k + 1], polygons[i * 24 + 3 * k + 2]);
	uint32_t vc = clip_to_horizon<V + 1>(counts[i], v, min_vertices);
	for (int k = 0; k != V + 1; ++k) { polygons[i * 24 + 3 * k] = v[k].x; polygons[i * 24 + 3 * k + 1] = v[k].y; polygons[i * 24 + 3 * k + 2] = v[k].z; }
	counts[i] = vc;
}

__global__ void kat_form_factor_kernel(const float* polygons, const uint32_t* counts, float* out, uint32_t count) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	float3 v[RL_MAX_P];
	for (int k = 0; k != RL_MAX_P; ++k) v[k] = mk3(polygons[i * 24 + 3 * k], polygons[i * 24 + 3 * k + 1], polygons[i * 24 + 3 * k + 2]);
	out[i] = polygon_form_factor<RL_MAX_P>(counts[i], v);
}

// out_polygons: 44 floats per polygon = {vc, v[8][2], e[8][2], inner0[2], sector[8], total}
template <int P>
__global__ void kat_psa_kernel(const float* polygons, const uint32_t* counts, const float* randoms, float* out_polygons, float* out_dirs, uint32_t count, uint32_t fast, uint32_t biased) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	float3 v[P];
	for (int k = 0; k != P; ++k) v[k] = mk3(polygons[i * 24 + 3 * k], polygons[i * 24 + 3 * k + 1], polygons[i * 24 + 3 * k + 2]);
	PsaPolygon<P> p;
	for (int k = 0; k != P; ++k) { p.v[k] = mk2(0.0f, 0.0f); p.e[k] = mk2(0.0f, 0.0f); p.sector[k] = 0.0f; }
	psa_prepare_rt<P>(p, counts[i], v, fast != 0);
	float* o = out_polygons + (size_t) i * 44;
	for (int k = 0; k != 44; ++k) o[k] = 0.0f;
	o[0] = (float) p.vc;
	for (int k = 0; k != P; ++k) {
		if ((uint32_t) k >= p.vc) break;
		o[1 + 2 * k] = p.v[k].x; o[2 + 2 * k] = p.v[k].y;
		o[17 + 2 * k] = p.e[k].x; o[18 + 2 * k] = p.e[k].y;
		o[35 + k] = p.sector[k];
	}
	o[33] = p.inner0.x; o[34] = p.inner0.y; o[43] = p.total;
	float3 d = psa_sample_rt<P>(p, randoms[2 * i], randoms[2 * i + 1], fast != 0, biased != 0);
	out_dirs[3 * i] = d.x; out_dirs[3 * i + 1] = d.y; out_dirs[3 * i + 2] = d.z;
}

__global__ void kat_noise_kernel(uint32_t width, uint32_t height, uint32_t frame_word, uint32_t draws, float* out) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= width * height) return;
	uint32_t seed = noise_seed(i % width, i / width, width, frame_word);
	for (uint32_t k = 0; k != draws; ++k) out[(size_t) i * draws + k] = noise_next(seed);
}

__global__ void kat_ltc_kernel(SceneView s, const float* in, float c0, float c1, float c2, float c3, float c4, float c5, float* out, uint32_t count) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const float* a = in + 11 * (size_t) i;
	float c[6] = { c0, c1, c2, c3, c4, c5 };
	LtcFrame l = make_ltc_frame(s, a[0], a[1], mk3(a[2], a[3], a[4]), mk3(a[5], a[6], a[7]), mk3(a[8], a[9], a[10]), c);
	float* o = out + 33 * (size_t) i;
	// world_to_shading as mat4x3 [column][row]
	const float3 rows[3] = { l.rx, l.ry, l.rz };
	for (int r = 0; r != 3; ++r) { o[0 + r] = rows[r].x; o[3 + r] = rows[r].y; o[6 + r] = rows[r].z; }
	o[9] = l.t.x; o[10] = l.t.y; o[11] = l.t.z;
	// shading_to_cosine / cosine_to_shading as mat3 [column][row]
	float s2c[9] = { l.s00, 0.0f, l.s20, 0.0f, l.s11, 0.0f, l.s02, 0.0f, l.s22 };
	float c2s[9] = { l.c00, 0.0f, l.c20, 0.0f, l.c11, 0.0f, l.c02, 0.0f, l.c22 };
	for (int k = 0; k != 9; ++k) { o[12 + k] = s2c[k]; o[21 + k] = c2s[k]; }
	o[30] = l.albedo; o[31] = l.det; o[32] = 0.0f;
}

__global__ void kat_any_hit_kernel(SceneView s, const float* rays, uint32_t* hits, uint32_t count) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const float* r = rays + 8 * (size_t) i;
	hits[i] = bvh_any_hit(s, mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3], r[7]) ? 1u : 0u;
}

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

extern "C" int risltc_cuda_kat_clip(risltc_device_t* d, float* polygons, uint32_t* vertex_counts, uint32_t count, uint32_t max_light_vertices) {
	if (use(d)) return 1;
	DeviceArray<float> p; DeviceArray<uint32_t> c;
	if (p.init(polygons, (size_t) count * 24) || c.init(vertex_counts, count)) return fail("kat_clip: allocation failed", nullptr);
	uint32_t min_vertices = 3;
	switch (max_light_vertices) {
	case 3: kat_clip_kernel<3><<<KAT_GRID(count)>>>(p.p, c.p, count, min_vertices); break;
	case 4: kat_clip_kernel<4><<<KAT_GRID(count)>>>(p.p, c.p, count, min_vertices); break;
	case 5: kat_clip_kernel<5><<<KAT_GRID(count)>>>(p.p, c.p, count, min_vertices); break;
	case 6: kat_clip_kernel<6><<<KAT_GRID(count)>>>(p.p, c.p, count, min_vertices); break;
	case 7: kat_clip_kernel<7><<<KAT_GRID(count)>>>(p.p, c.p, count, min_vertices); break;
	default: return fail("kat_clip: max_light_vertices must be 3..7", nullptr);
	}
	CU(cudaDeviceSynchronize());
	if (p.fetch(polygons) || c.fetch(vertex_counts)) return fail("kat_clip: read-back failed", nullptr);
	return 0;
}

extern "C" int risltc_cuda_kat_ltc_integral(risltc_device_t* d, const float* polygons, const uint32_t* vertex_counts, float* out, uint32_t count) {
	if (use(d)) return 1;
	DeviceArray<float> p, o; DeviceArray<uint32_t> c;
	if (p.init(polygons, (size_t) count * 24) || c.init(vertex_counts, count) || o.init(nullptr, count)) return fail("kat_ltc_integral: allocation failed", nullptr);
	kat_form_factor_kernel<<<KAT_GRID(count)>>>(p.p, c.p, o.p, count);
	CU(cudaDeviceSynchronize());
	if (o.fetch(out)) return fail("kat_ltc_integral: read-back failed", nullptr);
	return 0;
}

extern "C" int risltc_cuda_kat_psa(risltc_device_t* d, const float* polygons, const uint32_t* vertex_counts, const float* randoms,
	float* out_polygons, float* out_dirs, uint32_t count, uint32_t max_polygon_vertices, uint32_t fast_atan, uint32_t biased)
{
	if (use(d)) return 1;
	DeviceArray<float> p, r, op, od; DeviceArray<uint32_t> c;
	if (p.init(polygons, (size_t) count * 24) || c.init(vertex_counts, count) || r.init(randoms, (size_t) count * 2) || op.init(nullptr, (size_t) count * 44) || od.init(nullptr, (size_t) count * 3))
		return fail("kat_psa: allocation failed", nullptr);
	switch (max_polygon_vertices) {
	case 4: kat_psa_kernel<4><<<KAT_GRID(count)>>>(p.p, c.p, r.p, op.p, od.p, count, fast_atan, biased); break;
	case 5: kat_psa_kernel<5><<<KAT_GRID(count)>>>(p.p, c.p, r.p, op.p, od.p, count, fast_atan, biased); break;
	case 6: kat_psa_kernel<6><<<KAT_GRID(count)>>>(p.p, c.p, r.p, op.p, od.p, count, fast_atan, biased); break;
	case 7: kat_psa_kernel<7><<<KAT_GRID(count)>>>(p.p, c.p, r.p, op.p, od.p, count, fast_atan, biased); break;
	case 8: kat_psa_kernel<8><<<KAT_GRID(count)>>>(p.p, c.p, r.p, op.p, od.p, count, fast_atan, biased); break;
	default: return fail("kat_psa: max_polygon_vertices must be 4..8", nullptr);
	}
	CU(cudaDeviceSynchronize());
	if (op.fetch(out_polygons) || od.fetch(out_dirs)) return fail("kat_psa: read-back failed", nullptr);
	return 0;
}

extern "C" int risltc_cuda_kat_noise(risltc_device_t* d, uint32_t width, uint32_t height, uint32_t frame_word, uint32_t draws, float* out) {
	if (use(d)) return 1;
	DeviceArray<float> o;
	size_t n = (size_t) width * height;
	if (o.init(nullptr, n * draws)) return fail("kat_noise: allocation failed", nullptr);
	kat_noise_kernel<<<KAT_GRID((uint32_t) n)>>>(width, height, frame_word, draws, o.p);
	CU(cudaDeviceSynchronize());
	if (o.fetch(out)) return fail("kat_noise: read-back failed", nullptr);
	return 0;
}

extern "C" int risltc_cuda_kat_ltc_coefficients(risltc_device_t* d, const float* inputs, const float c[6], float* out, uint32_t count) {
	if (use(d)) return 1;
	if (!rl_view(d).ltc_rgba) return fail("kat_ltc_coefficients: upload_ltc first", nullptr);
	DeviceArray<float> in, o;
	if (in.init(inputs, (size_t) count * 11) || o.init(nullptr, (size_t) count * 33)) return fail("kat_ltc_coefficients: allocation failed", nullptr);
	kat_ltc_kernel<<<KAT_GRID(count)>>>(rl_view(d), in.p, c[0], c[1], c[2], c[3], c[4], c[5], o.p, count);
	CU(cudaDeviceSynchronize());
	if (o.fetch(out)) return fail("kat_ltc_coefficients: read-back failed", nullptr);
	return 0;
}

extern "C" int risltc_cuda_kat_any_hit(risltc_device_t* d, const float* rays, uint32_t* hits, uint32_t count) {
	if (use(d)) return 1;
	if (!rl_view(d).nodes) return fail("kat_any_hit: upload_scene first", nullptr);
	DeviceArray<float> r; DeviceArray<uint32_t> h;
	if (r.init(rays, (size_t) count * 8) || h.init(nullptr, count)) return fail("kat_any_hit: allocation failed", nullptr);
	kat_any_hit_kernel<<<KAT_GRID(count)>>>(rl_view(d), r.p, h.p, count);
	CU(cudaDeviceSynchronize());
	if (h.fetch(hits)) return fail("kat_any_hit: read-back failed", nullptr);
	return 0;
}

// Exhaustive checks of the hand-written exactly rounded sequences of common.cuh against the IEEE operations they replace:
// out[0] = floats in [2^-64, 2^64) whose inversesqrt differs from 1 / sqrt, out[1] = values of 0..65535 whose unorm16
// differs from x / 65535, out[2] = floats outside the fast range (or special) whose inversesqrt differs bitwise.
__global__ void kat_exact_math_kernel(unsigned long long* out) {
	const uint32_t stride = gridDim.x * blockDim.x;
	uint32_t bad_fast = 0, bad_unorm = 0, bad_other = 0;
	for (uint64_t b = blockIdx.x * blockDim.x + threadIdx.x; b < 0x100000000ull; b += stride) {
		const float x = __uint_as_float((uint32_t) b);
		const uint32_t got = __float_as_uint(inversesqrt(x)), want = __float_as_uint(__frcp_rn(__fsqrt_rn(x)));
		const bool fast = ((uint32_t) b - 0x1F800000u) < 0x40000000u;
		const bool same = got == want || (isnan(__uint_as_float(got)) && isnan(__uint_as_float(want)));
		if (!same) { if (fast) ++bad_fast; else ++bad_other; unsigned long long k = atomicAdd(&out[3], 1ull); if (k < 12) out[4 + k] = ((unsigned long long) got << 32) | (uint32_t) b; }
		if (b < 65536u && __float_as_uint(unorm16((uint32_t) b)) != __float_as_uint(__fdiv_rn((float) (uint32_t) b, 65535.0f))) ++bad_unorm;
	}
	if (bad_fast) atomicAdd(&out[0], (unsigned long long) bad_fast);
	if (bad_unorm) atomicAdd(&out[1], (unsigned long long) bad_unorm);
	if (bad_other) atomicAdd(&out[2], (unsigned long long) bad_other);
}
extern "C" int risltc_cuda_kat_exact_math(risltc_device_t* d, uint64_t mismatches[3]) {
	if (use(d)) return 1;
	DeviceArray<unsigned long long> o;
	if (o.init(nullptr, 16)) return fail("kat_exact_math: allocation failed", nullptr);
	CU(cudaMemset(o.p, 0, 16 * sizeof(unsigned long long)));
	kat_exact_math_kernel<<<rl_sm_count(d) * 8, 256>>>(o.p);
	CU(cudaDeviceSynchronize());
	unsigned long long h[16];
	if (o.fetch(h)) return fail("kat_exact_math: read-back failed", nullptr);
	for (int i = 0; i != 3; ++i) mismatches[i] = h[i];
	for (int i = 4; i != 16; ++i) if (h[i]) printf("kat_exact_math: x = 0x%08x, inversesqrt gives 0x%08x\n", (unsigned) h[i], (unsigned) (h[i] >> 32));
	return 0;
}
