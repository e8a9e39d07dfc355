// fp64_reuse_probe.cu -- can the ring hot loop reach the full FP64 issue rate by arranging for operand reuse
// (accumulate DADD immediately followed by the chain's DFMA that reads the same x_j)?
#include <cstdio>
#include <cuda_runtime.h>

template <int V, int C>
__global__ void __launch_bounds__(128, 3) probe(double *out, int iters, double seed)
{
	double acc[C + 2];
#pragma unroll
	for (int j = 0; j < C + 2; ++j) acc[j] = 0.0;
	double t0 = seed * 1.01 + threadIdx.x * 1e-9, t1 = seed * 1.02 + threadIdx.x * 1e-9, t2 = seed * 1.03 + threadIdx.x * 1e-9;
	for (int i = 0; i < iters; ++i) {
		double p0 = t0 * 0.5, p1 = t1 * 0.5, p2 = t2 * 0.5, q0 = 1, q1 = 1, q2 = 1;
		if (V == 0) {
#pragma unroll
			for (int j = 0; j < C; ++j) {
				const double x0 = fma(t0, q0, -p0), x1 = fma(t1, q1, -p1), x2 = fma(t2, q2, -p2);
				p0 = q0; q0 = x0; p1 = q1; q1 = x1; p2 = q2; q2 = x2;
				acc[j] += (x0 + x1) + x2;
			}
		} else if (V == 1) {
			// accumulate x_j of each chain right before the DFMA that consumes x_j
#pragma unroll
			for (int j = 0; j < C; ++j) {
				acc[j] += q0; { const double x = fma(t0, q0, -p0); p0 = q0; q0 = x; }
				acc[j] += q1; { const double x = fma(t1, q1, -p1); p1 = q1; q1 = x; }
				acc[j] += q2; { const double x = fma(t2, q2, -p2); p2 = q2; q2 = x; }
			}
		} else if (V == 2) {
			// same, chains skewed by one step so that the three adds of a step hit different accumulators
			{ const double x = fma(t1, q1, -p1); p1 = q1; q1 = x; }
			{ const double x = fma(t2, q2, -p2); p2 = q2; q2 = x; }
			{ const double x = fma(t2, q2, -p2); p2 = q2; q2 = x; }
#pragma unroll
			for (int j = 0; j < C; ++j) {
				acc[j] += q0; { const double x = fma(t0, q0, -p0); p0 = q0; q0 = x; }
				acc[j + 1] += q1; { const double x = fma(t1, q1, -p1); p1 = q1; q1 = x; }
				acc[j + 2] += q2; { const double x = fma(t2, q2, -p2); p2 = q2; q2 = x; }
			}
		}
		t0 += q0 * 1e-30; t1 += q1 * 1e-30; t2 += q2 * 1e-30;
	}
	double r = 0;
#pragma unroll
	for (int j = 0; j < C + 2; ++j) r += acc[j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int V, int C>
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
		probe<V, C><<<blocks, threads>>>(d, iters, 0.5);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (rep && ms < best) best = ms;
	}
	double lane_ops = 6.0 * C * iters * (double)blocks * threads;
	double rate = lane_ops / (best * 1e-3);
	printf("%-44s warps/SM %2d %8.3f ms %6.2f T lane-ops/s = %5.1f lanes/clk/SM @1.9GHz (%s)\n", name, warps_per_sm, best, rate / 1e12,
	       rate / prop.multiProcessorCount / 1.9e9, cudaGetErrorString(cudaGetLastError()));
	cudaFree(d);
}

int main()
{
	for (int w : {4, 12}) {
		run<0, 43>("V0 axis pre-sum (current kernel)", w);
		run<1, 43>("V1 per-chain add adjacent to its DFMA", w);
		run<2, 43>("V2 V1 + chains skewed by one step", w);
	}
	return 0;
}
