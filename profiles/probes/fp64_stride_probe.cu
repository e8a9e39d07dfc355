// fp64_stride_probe.cu -- stride-s Chebyshev recurrence: x_{m+s} = 2cos(s d) x_m - x_{m-s}.  The s interleaved
// sub-chains of one axis share the multiplier register, so consecutive DFMAs can hit the operand-reuse cache
// and issue at the 2-register rate.  Variants: S = 1 (current kernel), 2, 4.  C = 44 accumulators, 3 axes.
#include <cstdio>
#include <cuda_runtime.h>

template <int S, int C>
__global__ void __launch_bounds__(128, 2) probe(double *out, int iters, double seed)
{
	double acc[C];
#pragma unroll
	for (int j = 0; j < C; ++j) acc[j] = 0.0;
	double t[3] = { seed * 1.01 + threadIdx.x * 1e-9, seed * 1.02 + threadIdx.x * 1e-9, seed * 1.03 + threadIdx.x * 1e-9 };
	for (int it = 0; it < iters; ++it) {
		double x[3][S], xp[3][S];
#pragma unroll
		for (int a = 0; a < 3; ++a)
#pragma unroll
			for (int r = 0; r < S; ++r) { x[a][r] = 1.0 + 0.001 * r + t[a] * 1e-3; xp[a][r] = t[a] * (0.5 + 0.001 * r); }
#pragma unroll
		for (int j = 0; j < C / S; ++j) {
			double s[S];
#pragma unroll
			for (int r = 0; r < S; ++r) s[r] = 0.0;
#pragma unroll
			for (int a = 0; a < 3; ++a) {
#pragma unroll
				for (int r = 0; r < S; ++r) { // S consecutive DFMAs share t[a]
					const double xn = fma(t[a], x[a][r], -xp[a][r]);
					xp[a][r] = x[a][r]; x[a][r] = xn;
				}
			}
#pragma unroll
			for (int r = 0; r < S; ++r) acc[j * S + r] += (x[0][r] + x[1][r]) + x[2][r];
		}
		t[0] += x[0][0] * 1e-30; t[1] += x[1][0] * 1e-30; t[2] += x[2][0] * 1e-30;
	}
	double r = 0;
#pragma unroll
	for (int j = 0; j < C; ++j) r += acc[j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int S, int C>
void run(const char *name, int warps_per_sm)
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
		probe<S, C><<<blocks, threads>>>(d, iters, 0.5);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (rep && ms < best) best = ms;
	}
	double lane_ops = 6.0 * C * iters * (double)blocks * threads;
	double rate = lane_ops / (best * 1e-3);
	printf("%-28s warps/SM %2d %8.3f ms %6.2f T lane-ops/s = %5.1f lanes/clk/SM @1.9GHz (%s)\n", name, warps_per_sm, best, rate / 1e12,
	       rate / prop.multiProcessorCount / 1.9e9, cudaGetErrorString(cudaGetLastError()));
	cudaFree(d);
}

int main()
{
	for (int w : {4, 8}) {
		run<1, 44>("stride 1 (current)", w);
		run<2, 44>("stride 2", w);
		run<4, 44>("stride 4", w);
	}
	return 0;
}
