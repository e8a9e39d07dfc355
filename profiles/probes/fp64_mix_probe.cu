// fp64_mix_probe.cu -- what FP64 issue rate does the ring kernel's hot-loop instruction mix reach on its own?
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix_probe fp64_mix_probe.cu
// Variants: 0 pure DFMA (2 constant operands, the peak microbenchmark)   1 DFMA with 3 distinct operands
//           2 hot loop as in self_ring_kernel (3 chains: 3 DFMA + 3 DADD per step, C=43 accumulators)
//           3 hot loop with 6 chains (two particles interleaved)            4 pure DADD
#include <cstdio>
#include <cuda_runtime.h>

template <int V, int C>
__global__ void __launch_bounds__(128) probe(double *out, int iters, double seed)
{
	double acc[C];
#pragma unroll
	for (int j = 0; j < C; ++j) acc[j] = 0.0;
	double t0 = seed * 1.01 + threadIdx.x * 1e-9, t1 = seed * 1.02, t2 = seed * 1.03;
	double u0 = seed * 1.04, u1 = seed * 1.05, u2 = seed * 1.06;
	double r = 0;
	for (int i = 0; i < iters; ++i) {
		if (V == 0) {
#pragma unroll
			for (int j = 0; j < C; ++j) acc[j] = fma(acc[j], 1.0000001, 1e-9);
		} else if (V == 1) {
#pragma unroll
			for (int j = 0; j < C; ++j) acc[j] = fma(acc[(j + 1) % C], acc[(j + 2) % C], acc[j]);
		} else if (V == 4) {
#pragma unroll
			for (int j = 0; j < C; ++j) acc[j] = acc[j] + acc[(j + 1) % C];
		} else if (V == 2) {
			double p0 = t0, p1 = t1, p2 = t2, q0 = 1, q1 = 1, q2 = 1;
#pragma unroll
			for (int j = 0; j < C; ++j) {
				const double x0 = fma(t0, q0, -p0), x1 = fma(t1, q1, -p1), x2 = fma(t2, q2, -p2);
				p0 = q0; q0 = x0; p1 = q1; q1 = x1; p2 = q2; q2 = x2;
				acc[j] += (x0 + x1) + x2;
			}
			t0 += q0 * 1e-30; t1 += q1 * 1e-30; t2 += q2 * 1e-30;
		} else if (V == 3) {
			double p0 = t0, p1 = t1, p2 = t2, q0 = 1, q1 = 1, q2 = 1;
			double a0 = u0, a1 = u1, a2 = u2, b0 = 1, b1 = 1, b2 = 1;
#pragma unroll
			for (int j = 0; j < C; ++j) {
				const double x0 = fma(t0, q0, -p0), x1 = fma(t1, q1, -p1), x2 = fma(t2, q2, -p2);
				const double y0 = fma(u0, b0, -a0), y1 = fma(u1, b1, -a1), y2 = fma(u2, b2, -a2);
				p0 = q0; q0 = x0; p1 = q1; q1 = x1; p2 = q2; q2 = x2;
				a0 = b0; b0 = y0; a1 = b1; b1 = y1; a2 = b2; b2 = y2;
				acc[j] += ((x0 + x1) + (x2 + y0)) + (y1 + y2);
			}
			t0 += q0 * 1e-30; t1 += q1 * 1e-30; t2 += q2 * 1e-30; u0 += b0 * 1e-30;
		}
	}
#pragma unroll
	for (int j = 0; j < C; ++j) r += acc[j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int V, int C>
void run(const char *name, double instr_per_iter, int warps_per_sm)
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
	double lane_ops = instr_per_iter * iters * (double)blocks * threads;
	double rate = lane_ops / (best * 1e-3);
	double per_sm_clk = rate / prop.multiProcessorCount / 1.9e9;
	printf("%-34s warps/SM %2d  %8.3f ms  %7.2f T lane-ops/s  = %5.1f lanes/clk/SM @1.9GHz (64 = full rate)\n", name, warps_per_sm, best, rate / 1e12, per_sm_clk);
	cudaFree(d);
}

int main()
{
	for (int w : {4, 8, 12, 16}) {
		if (w == 4) { run<0, 43>("DFMA const operands", 43, 4); run<1, 43>("DFMA 3 distinct operands", 43, 4); run<4, 43>("DADD", 43, 4);
		              run<2, 43>("ring hot loop 3 chains C=43", 6 * 43, 4); run<3, 43>("ring hot loop 6 chains C=43", 12 * 43, 4); }
		if (w == 8) { run<0, 43>("DFMA const operands", 43, 8); run<1, 43>("DFMA 3 distinct operands", 43, 8); run<4, 43>("DADD", 43, 8);
		              run<2, 43>("ring hot loop 3 chains C=43", 6 * 43, 8); run<3, 43>("ring hot loop 6 chains C=43", 12 * 43, 8); }
		if (w == 12) { run<0, 43>("DFMA const operands", 43, 12); run<1, 43>("DFMA 3 distinct operands", 43, 12); run<4, 43>("DADD", 43, 12);
		               run<2, 43>("ring hot loop 3 chains C=43", 6 * 43, 12); run<3, 43>("ring hot loop 6 chains C=43", 12 * 43, 12); }
		if (w == 16) { run<0, 43>("DFMA const operands", 43, 16); run<1, 43>("DFMA 3 distinct operands", 43, 16); run<4, 43>("DADD", 43, 16);
		               run<2, 43>("ring hot loop 3 chains C=43", 6 * 43, 16); }
	}
	return 0;
}
