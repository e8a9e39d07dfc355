// fp64_mix_probe2.cu -- FP64 pipe rates on B200 per instruction kind and for DFMA/DADD mixes.
#include <cstdio>
#include <cuda_runtime.h>

// V: 0 DFMA (reg, const, const)  1 DFMA (reg, reg, reg) independent  2 DADD (reg, reg) independent
//    3 DMUL (reg, reg)           4 alternate DFMA/DADD independent    5 2 DADD : 1 DFMA   6 DADD (reg,const)
template <int V>
__global__ void __launch_bounds__(128) probe(double *out, int iters, double seed)
{
	constexpr int C = 24;
	double acc[C], b[4], c[4];
#pragma unroll
	for (int j = 0; j < C; ++j) acc[j] = seed + j * 1e-3 + threadIdx.x * 1e-6;
#pragma unroll
	for (int j = 0; j < 4; ++j) { b[j] = 1.0 + (seed + j) * 1e-9 + threadIdx.x * 1e-12; c[j] = (seed + j) * 1e-12 + threadIdx.x * 1e-15; }
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int j = 0; j < C; ++j) {
			if (V == 0) acc[j] = fma(acc[j], 1.0000001, 1e-9);
			if (V == 1) acc[j] = fma(acc[j], b[j & 3], c[(j + 1) & 3]);
			if (V == 2) acc[j] = acc[j] + c[j & 3];
			if (V == 3) acc[j] = acc[j] * b[j & 3];
			if (V == 4) acc[j] = (j & 1) ? fma(acc[j], b[j & 3], c[(j + 1) & 3]) : acc[j] + c[j & 3];
			if (V == 5) acc[j] = (j % 3 == 0) ? fma(acc[j], b[j & 3], c[(j + 1) & 3]) : acc[j] + c[j & 3];
			if (V == 6) acc[j] = acc[j] + 1e-9;
		}
	}
	double r = 0;
#pragma unroll
	for (int j = 0; j < C; ++j) r += acc[j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int V>
void run(const char *name, int warps_per_sm)
{
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	int threads = 128, blocks = prop.multiProcessorCount * warps_per_sm / 4, iters = 40000;
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
	double lane_ops = 24.0 * iters * (double)blocks * threads;
	double rate = lane_ops / (best * 1e-3);
	printf("%-40s warps/SM %2d  %8.3f ms  %6.2f T lane-ops/s = %5.1f lanes/clk/SM @1.9GHz  (%s)\n", name, warps_per_sm, best, rate / 1e12,
	       rate / prop.multiProcessorCount / 1.9e9, cudaGetErrorString(cudaGetLastError()));
	cudaFree(d);
}

int main()
{
	for (int w : {4, 16}) {
		run<0>("DFMA reg,const,const", w);
		run<1>("DFMA reg,reg,reg", w);
		run<2>("DADD reg,reg", w);
		run<6>("DADD reg,const", w);
		run<3>("DMUL reg,reg", w);
		run<4>("alternate DFMA / DADD", w);
		run<5>("1 DFMA : 2 DADD", w);
	}
	return 0;
}
