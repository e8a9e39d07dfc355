// dmma_phase_probe.cu -- how the FP64 pipe of one B200 SM sub-partition is shared between warps that alternate a
// "generate Chebyshev rows into shared memory" phase (vector DFMA/DADD, STS) and a "fragment loads + DMMA" phase,
// the structure of self_ring_mma_kernel.  Prints SM sub-partition cycles per (generate + 24 DMMA) unit;
// the pipe alone needs 24 x 16 (DMMA) + ~50 x 2.4 (vector) = ~505.
//   V0 natural (every warp free-running)        V1 phases aligned over the block with __syncthreads (1 block = all warps of the SM)
//   V2 generate only                             V3 DMMA only
//   V4 natural, 6 accumulators (even/odd k-block) V5 one warp-specialised block: even warps generate for odd warps (double buffer)
#include <cstdio>
#include <cuda_runtime.h>
#define PITCH 36
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void gen(double *su, double *sv, double c, double TS, int lane)
{
	const double t = c + c;
	const double T2 = fma(t, c, -1.0), T3 = fma(t, T2, -c);
	const double d2 = T2 + T2, d3 = T3 + T3;
	const double T4 = fma(d2, T2, -1.0), T5 = fma(d2, T3, -c), T6 = fma(d3, T3, -1.0), T7 = fma(d3, T4, -c);
	sv[0 * PITCH + lane] = 1.0; sv[1 * PITCH + lane] = c; sv[2 * PITCH + lane] = T2; sv[3 * PITCH + lane] = T3;
	sv[4 * PITCH + lane] = T4; sv[5 * PITCH + lane] = T5; sv[6 * PITCH + lane] = T6; sv[7 * PITCH + lane] = T7;
	const double e1 = TS + TS;
	const double U2 = fma(e1, TS, -1.0), U3 = fma(e1, U2, -TS);
	const double e2 = U2 + U2, e3 = U3 + U3;
	const double U4 = fma(e2, U2, -1.0), U5 = fma(e2, U3, -TS), U6 = fma(e3, U3, -1.0), U7 = fma(e3, U4, -TS);
	su[0 * PITCH + lane] = 1.0; su[1 * PITCH + lane] = TS; su[2 * PITCH + lane] = U2; su[3 * PITCH + lane] = U3;
	su[4 * PITCH + lane] = U4; su[5 * PITCH + lane] = U5; su[6 * PITCH + lane] = U6; su[7 * PITCH + lane] = U7;
	const double t4 = T4 + T4;
	double p0 = 1.0, p1 = c, p2 = T2, p3 = T3, x0 = T4, x1 = T5, x2 = T6, x3 = T7;
#pragma unroll
	for (int r4 = 8; r4 < 24; r4 += 4) {
		const double n0 = fma(t4, x0, -p0), n1 = fma(t4, x1, -p1), n2 = fma(t4, x2, -p2), n3 = fma(t4, x3, -p3);
		p0 = x0; p1 = x1; p2 = x2; p3 = x3; x0 = n0; x1 = n1; x2 = n2; x3 = n3;
		double *row = sv + r4 * PITCH + lane;
		row[0] = n0; row[PITCH] = n1; row[2 * PITCH] = n2; row[3 * PITCH] = n3;
	}
}
template <int NACC>
__device__ __forceinline__ void mma_phase(double (&D)[NACC][3][2], const double *su, const double *sv, int lane)
{
	const int arow = lane >> 2, acol = lane & 3;
	const double *pu = su + arow * PITCH + acol, *pv = sv + arow * PITCH + acol;
	double af = pu[0], bf[3];
#pragma unroll
	for (int j = 0; j < 3; ++j) bf[j] = pv[8 * j * PITCH];
#pragma unroll
	for (int kb = 0; kb < 8; ++kb) {
		double an = 0, bn[3] = {0, 0, 0};
		if (kb < 7) {
			an = pu[4 * (kb + 1)];
#pragma unroll
			for (int j = 0; j < 3; ++j) bn[j] = pv[8 * j * PITCH + 4 * (kb + 1)];
		}
#pragma unroll
		for (int j = 0; j < 3; ++j) dmma884(D[kb % NACC][j][0], D[kb % NACC][j][1], af, bf[j]);
		af = an;
#pragma unroll
		for (int j = 0; j < 3; ++j) bf[j] = bn[j];
	}
}

template <int V>
__global__ void __launch_bounds__(512) probe(double *out, int iters, double seed)
{
	extern __shared__ double smem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	constexpr int NACC = V == 4 ? 2 : 1;
	double D[NACC][3][2] = {};
	double c = seed + threadIdx.x * 1e-4, TS = seed * 0.7 - threadIdx.x * 1e-4;
	if (V != 5) {
		double *su = smem + warp * 32 * PITCH, *sv = su + 8 * PITCH;
		for (int it = 0; it < iters; ++it) {
			c = fma(c, 0.999, 1e-4); TS = fma(TS, 0.999, -1e-4);
			if (V != 3) gen(su, sv, c, TS, lane);
			else if (it == 0) gen(su, sv, c, TS, lane);
			if (V == 1) __syncthreads(); else __syncwarp();
			if (V != 2) mma_phase<NACC>(D, su, sv, lane);
			if (V == 1) __syncthreads(); else __syncwarp();
		}
	} else {
		// pair p = warp/2: even warp generates into buffer (it&1), odd warp consumes; named barriers per pair and buffer
		const int pair = warp >> 1, role = warp & 1;
		double *base = smem + pair * 2 * 32 * PITCH;
		for (int it = 0; it < iters; ++it) {
			double *su = base + (it & 1) * 32 * PITCH, *sv = su + 8 * PITCH;
			const int full = 1 + pair * 4 + (it & 1), empty = 1 + pair * 4 + 2 + (it & 1);
			if (role == 0) {
				c = fma(c, 0.999, 1e-4); TS = fma(TS, 0.999, -1e-4);
				if (it >= 2) asm volatile("bar.sync %0, 64;" ::"r"(empty) : "memory");
				gen(su, sv, c, TS, lane);
				asm volatile("bar.arrive %0, 64;" ::"r"(full) : "memory");
			} else {
				asm volatile("bar.sync %0, 64;" ::"r"(full) : "memory");
				mma_phase<NACC>(D, su, sv, lane);
				if (it + 2 < iters) asm volatile("bar.arrive %0, 64;" ::"r"(empty) : "memory");
			}
		}
	}
	double r = c;
#pragma unroll
	for (int i = 0; i < NACC; ++i)
#pragma unroll
		for (int j = 0; j < 3; ++j) r += D[i][j][0] + D[i][j][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int V>
void run(const char *name, int warps_per_sm, int warps_per_block)
{
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	int threads = 32 * warps_per_block, blocks = prop.multiProcessorCount * warps_per_sm / warps_per_block, iters = 4000;
	size_t sm = (size_t)warps_per_block * 32 * PITCH * 8;
	cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
	double *d;
	cudaMalloc(&d, sizeof(double) * blocks * threads);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 4; ++rep) {
		cudaEventRecord(e0);
		probe<V><<<blocks, threads, sm>>>(d, iters, 0.3);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (rep && ms < best) best = ms;
	}
	int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	// units per SM sub-partition: V5 has one producing and one consuming warp per unit
	double units = (double)iters * warps_per_sm / 4 / (V == 5 ? 2 : 1);
	printf("%-44s warps/SM %2d block %2d  %8.3f ms  %7.1f clk per unit per sub-partition (%s)\n", name, warps_per_sm, warps_per_block, best,
	       best * 1e-3 * 1.965e9 / units, cudaGetErrorString(cudaGetLastError()));
	cudaFree(d);
}

int main()
{
	for (int w : {4, 8, 16}) {
		run<2>("generate only", w, 4);
		run<3>("DMMA only", w, 4);
		run<0>("natural", w, 4);
		run<4>("natural, 2 accumulator sets", w, 4);
		run<1>("phases aligned by __syncthreads", w, w);
		run<5>("producer/consumer warp pairs", w, w);
	}
	run<0>("natural", 20, 4);
	run<0>("natural", 24, 4);
	return 0;
}
