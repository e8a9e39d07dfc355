// dmma_probe.cu -- FP64 tensor-core (mma.sync f64) throughput on B200, alone and next to vector DFMA work.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double &d0, double &d1, double &d2, double &d3, double a0, double a1, double a2, double a3, double b0, double b1)
{
	asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
	             : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}

// V: 0 m8n8k4 x8 independent accumulators; 1 m16n8k8 x4 accumulators; 2 = V0 interleaved with 8 vector DFMA per mma
template <int V>
__global__ void __launch_bounds__(128) probe(double *out, int iters, double seed)
{
	double d[16];
#pragma unroll
	for (int i = 0; i < 16; ++i) d[i] = 0.0;
	double a = seed + threadIdx.x * 1e-6, b = seed * 0.5 + threadIdx.x * 1e-7;
	double v[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) v[i] = seed + i;
	for (int it = 0; it < iters; ++it) {
		if (V == 0 || V == 2) {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				dmma884(d[2 * i], d[2 * i + 1], a, b);
				if (V == 2) {
#pragma unroll
					for (int j = 0; j < 8; ++j) v[j] = fma(v[j], 1.0000001, 1e-9);
				}
			}
		} else {
#pragma unroll
			for (int i = 0; i < 4; ++i) dmma1688(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3], a, b, a, b, b, a);
		}
		a += 1e-12;
	}
	double r = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) r += d[i];
#pragma unroll
	for (int i = 0; i < 8; ++i) r += v[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int V>
void run(const char *name, int warps_per_sm, double fma_per_warp_iter, double vec_per_warp_iter)
{
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	int threads = 128, blocks = prop.multiProcessorCount * warps_per_sm / 4, iters = 20000;
	double *d;
	cudaMalloc(&d, sizeof(double) * blocks * threads);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 4; ++rep) {
		cudaEventRecord(e0);
		probe<V><<<blocks, threads>>>(d, iters, 0.5);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (rep && ms < best) best = ms;
	}
	double warps = (double)blocks * threads / 32;
	double tf = 2.0 * fma_per_warp_iter * iters * warps / (best * 1e-3) / 1e12;
	double vtf = 2.0 * vec_per_warp_iter * iters * warps / (best * 1e-3) / 1e12;
	printf("%-34s warps/SM %2d %8.3f ms  tensor %6.2f TFLOP/s  vector %6.2f TFLOP/s (%s)\n", name, warps_per_sm, best, tf, vtf, cudaGetErrorString(cudaGetLastError()));
	cudaFree(d);
}

int main()
{
	for (int w : {4, 8, 16}) {
		run<0>("mma m8n8k4 f64", w, 8 * 256.0, 0);
		run<1>("mma m16n8k8 f64", w, 4 * 1024.0, 0);
		run<2>("m8n8k4 + 8 DFMA per mma", w, 8 * 256.0, 8 * 8 * 32.0);
	}
	return 0;
}
