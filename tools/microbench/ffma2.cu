// ffma2.cu -- how fast is Blackwell's packed FP32 FMA (fma.rn.f32x2 -> SASS FFMA2) next to FFMA, per SM and per issue slot?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/ffma2 tools/microbench/ffma2.cu
// Prints Gfma/s (scalar FMAs counted, a packed instruction = 2) for: FFMA alone, FFMA2 alone, FFMA + independent ALU work,
// FFMA2 + the same ALU work (does packing free issue slots for the other pipe?).
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
	unsigned long long d;
	asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) kernel(float* out, int iters, float seed) {
	float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const float m = 1.0000001f, c = 1e-9f;
	unsigned long long p0, p1, p2, p3, pm, pc;
	{
		float2 t;
		t = make_float2(a0, a1); p0 = *reinterpret_cast<unsigned long long*>(&t);
		t = make_float2(a2, a3); p1 = *reinterpret_cast<unsigned long long*>(&t);
		t = make_float2(a4, a5); p2 = *reinterpret_cast<unsigned long long*>(&t);
		t = make_float2(a6, a7); p3 = *reinterpret_cast<unsigned long long*>(&t);
		t = make_float2(m, m); pm = *reinterpret_cast<unsigned long long*>(&t);
		t = make_float2(c, c); pc = *reinterpret_cast<unsigned long long*>(&t);
	}
	unsigned u0 = threadIdx.x, u1 = u0 * 3u, u2 = u0 * 5u, u3 = u0 * 7u;
	for (int i = 0; i < iters; ++i) {
		#pragma unroll
		for (int k = 0; k < 8; ++k) {
			if (MODE == 0 || MODE == 2) {
				a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
				a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
			}
			else {
				p0 = fma2(p0, pm, pc); p1 = fma2(p1, pm, pc); p2 = fma2(p2, pm, pc); p3 = fma2(p3, pm, pc);
			}
			if (MODE >= 2) {   // 8 independent ALU-pipe instructions (LOP3 / SHF) per 8 scalar FMAs
				u0 = (u0 ^ u1) + 0x9E3779B9u; u1 = __funnelshift_l(u1, u2, 7); u2 = (u2 & u3) | 0x55u; u3 = __funnelshift_l(u3, u0, 13);
				u0 = (u0 ^ u2) + 0x7F4A7C15u; u1 = __funnelshift_l(u1, u3, 9); u2 = (u2 | u0) ^ 0x33u; u3 = __funnelshift_l(u3, u1, 5);
			}
		}
	}
	float2 q0 = *reinterpret_cast<float2*>(&p0), q1 = *reinterpret_cast<float2*>(&p1), q2 = *reinterpret_cast<float2*>(&p2), q3 = *reinterpret_cast<float2*>(&p3);
	out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + q0.x + q0.y + q1.x + q1.y + q2.x + q2.y + q3.x + q3.y + (float) (u0 ^ u1 ^ u2 ^ u3);
}

template <int MODE>
static void run(const char* name, float* out, int sms) {
	const int iters = 4096, blocks = sms * 8;
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	kernel<MODE><<<blocks, 256>>>(out, 16, 1.0f);
	cudaEventRecord(a);
	kernel<MODE><<<blocks, 256>>>(out, iters, 1.0f);
	cudaEventRecord(b); cudaEventSynchronize(b);
	float ms = 0; cudaEventElapsedTime(&ms, a, b);
	const double fmas = (double) blocks * 256 * iters * 8 * 8;
	printf("%-28s %8.3f ms  %9.1f Gfma/s  (%.2f TFLOP/s)\n", name, ms, fmas / ms / 1e6, 2.0 * fmas / ms / 1e9);
}

int main() {
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	float* out; cudaMalloc(&out, (size_t) p.multiProcessorCount * 8 * 256 * sizeof(float));
	printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
	run<0>("FFMA", out, p.multiProcessorCount);
	run<1>("FFMA2", out, p.multiProcessorCount);
	run<2>("FFMA + ALU (1:1)", out, p.multiProcessorCount);
	run<3>("FFMA2 + ALU (1:2)", out, p.multiProcessorCount);
	return 0;
}
