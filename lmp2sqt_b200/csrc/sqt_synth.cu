// sqt_synth.cu -- on-device synthetic LJ-liquid-like frame source and the FP64 roofline probe.
//
// The reference can only read LAMMPS text dumps (snapshot.c:166-382); at the benchmark shapes
// (1 M - 16 M atoms x 5 000 frames) a text dump would be terabytes, so the bench workloads come
// from a deterministic generator behind the same frame interface:
// fcc lattice at reduced density 0.8442, Gaussian jitter 0.10, unwrapped Gaussian random walk
// 0.077 per axis per frame.  Randomness is counter based -- Philox4x32-10 keyed by the seed with
// counter (atom, frame) -- so any particle shard generates exactly the numbers the full system
// would, independent of the number of GPUs.
#include "../../include/sqt.h"
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

struct sqt_synth {
	int device, n_cells, n_local;
	long long first_atom;
	unsigned long long seed;
	long long frame; // next frame to generate
	double L, a;
	double *d_pos;   // SoA [3][n_local]
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4])
{
#pragma unroll
	for (int r = 0; r < 10; ++r) {
		const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
		const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
		const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// three independent standard normals for (seed, atom, frame)
__device__ __forceinline__ void normal3(unsigned long long seed, long long atom, long long frame, double g[3])
{
	uint32_t r[4], s[4];
	philox4x32_10((uint32_t)atom, (uint32_t)(atom >> 32), (uint32_t)frame, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
	philox4x32_10((uint32_t)atom, (uint32_t)(atom >> 32), (uint32_t)frame, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), s);
	const double u1 = ((double)r[0] + 0.5) * (1.0 / 4294967296.0), u2 = ((double)r[1] + 0.5) * (1.0 / 4294967296.0);
	const double u3 = ((double)r[2] + 0.5) * (1.0 / 4294967296.0), u4 = ((double)s[0] + 0.5) * (1.0 / 4294967296.0);
	const double m1 = sqrt(-2.0 * log(u1)), m2 = sqrt(-2.0 * log(u3));
	double sn, cs;
	sincospi(2.0 * u2, &sn, &cs);
	g[0] = m1 * cs; g[1] = m1 * sn;
	sincospi(2.0 * u4, &sn, &cs);
	g[2] = m2 * cs;
}

__global__ void synth_init_kernel(double *pos, int n, long long first, int n_cells, double a, unsigned long long seed)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const long long j = first + i;
	const int b = (int)(j & 3);
	const long long cell = j >> 2;
	const int cz = (int)(cell % n_cells), cy = (int)((cell / n_cells) % n_cells), cx = (int)(cell / ((long long)n_cells * n_cells));
	const double bx = (b == 1 || b == 2) ? 0.5 : 0.0, by = (b == 1 || b == 3) ? 0.5 : 0.0, bz = (b == 2 || b == 3) ? 0.5 : 0.0;
	double g[3];
	normal3(seed, j, -1, g);
	pos[i] = (cx + bx) * a + 0.10 * g[0];
	pos[(long long)n + i] = (cy + by) * a + 0.10 * g[1];
	pos[2LL * n + i] = (cz + bz) * a + 0.10 * g[2];
}

__global__ void synth_step_kernel(double *pos, int n, long long first, long long frame, unsigned long long seed)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double g[3];
	normal3(seed, first + i, frame, g);
	pos[i] += 0.077 * g[0];
	pos[(long long)n + i] += 0.077 * g[1];
	pos[2LL * n + i] += 0.077 * g[2];
}

extern "C" int sqt_synth_create(int device, int n_cells, long long first_atom, int n_atoms_local, unsigned long long seed,
                                sqt_synth **out)
{
	if (!out || n_cells < 1 || n_atoms_local < 1) return SQT_ERR_ARG;
	if (cudaSetDevice(device) != cudaSuccess) return SQT_ERR_CUDA;
	sqt_synth *s = (sqt_synth *)calloc(1, sizeof(*s));
	s->device = device; s->n_cells = n_cells; s->n_local = n_atoms_local; s->first_atom = first_atom; s->seed = seed;
	const double n_full = 4.0 * n_cells * (double)n_cells * n_cells;
	s->L = cbrt(n_full / 0.8442);
	s->a = s->L / n_cells;
	if (cudaMalloc(&s->d_pos, sizeof(double) * 3 * (size_t)n_atoms_local) != cudaSuccess) { free(s); return SQT_ERR_NOMEM; }
	*out = s;
	return SQT_OK;
}

extern "C" const double *sqt_synth_next(sqt_synth *s)
{
	if (!s) return NULL;
	cudaSetDevice(s->device);
	const int nb = (s->n_local + 255) / 256;
	// legacy default stream: implicitly ordered against the engine's (blocking) streams
	if (s->frame == 0) synth_init_kernel<<<nb, 256>>>(s->d_pos, s->n_local, s->first_atom, s->n_cells, s->a, s->seed);
	else synth_step_kernel<<<nb, 256>>>(s->d_pos, s->n_local, s->first_atom, s->frame, s->seed);
	if (cudaGetLastError() != cudaSuccess) return NULL;
	s->frame++;
	return s->d_pos;
}

extern "C" double sqt_synth_box_length(const sqt_synth *s) { return s ? s->L : 0.0; }

extern "C" int sqt_synth_copy_to_host(sqt_synth *s, double *dst)
{
	if (!s || !dst) return SQT_ERR_ARG;
	cudaSetDevice(s->device);
	return cudaMemcpy(dst, s->d_pos, sizeof(double) * 3 * (size_t)s->n_local, cudaMemcpyDeviceToHost) == cudaSuccess ? SQT_OK : SQT_ERR_CUDA;
}

extern "C" int sqt_synth_copy_to_device(sqt_synth *s, double *d_dst)
{
	if (!s || !d_dst) return SQT_ERR_ARG;
	cudaSetDevice(s->device);
	return cudaMemcpyAsync(d_dst, s->d_pos, sizeof(double) * 3 * (size_t)s->n_local, cudaMemcpyDeviceToDevice, 0) == cudaSuccess ? SQT_OK : SQT_ERR_CUDA;
}

extern "C" void sqt_synth_destroy(sqt_synth *s)
{
	if (!s) return;
	cudaSetDevice(s->device);
	cudaFree(s->d_pos);
	free(s);
}

// ---------------------------------------------------------------------------------------------
// FP64 roofline denominator.  MEASURED_PEAKS.json carries only HBM and bf16-tensor peaks; the ring
// kernel is bound by the non-tensor FP64 pipe, so its peak is measured the same way the kernel
// uses it: register-resident dependent DFMA chains, 8 independent chains per thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed)
{
	double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
	const double m = 1.0000001, c = 1e-9;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u = 0; u < 8; ++u) {
			a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
			a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" int sqt_measure_dfma_peak(int device, double *tflops, double *sm_clock_mhz_est)
{
	if (cudaSetDevice(device) != cudaSuccess) return SQT_ERR_CUDA;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SQT_ERR_CUDA;
	const int blocks = prop.multiProcessorCount * 4, threads = 256, iters = 1024;
	double *d = NULL;
	if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return SQT_ERR_NOMEM;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	double best = 0.0;
	for (int rep = 0; rep < 6; ++rep) {
		cudaEventRecord(e0);
		dfma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0 + rep);
		cudaEventRecord(e1);
		if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return SQT_ERR_CUDA; }
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		const double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
		const double tf = fl / (ms * 1e-3) / 1e12;
		if (rep >= 1 && tf > best) best = tf;
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaFree(d);
	if (tflops) *tflops = best;
	// 64 DFMA lanes per SM per clock (16 per SM sub-partition) => implied clock
	if (sm_clock_mhz_est) *sm_clock_mhz_est = best * 1e12 / (2.0 * 64.0 * prop.multiProcessorCount) / 1e6;
	return SQT_OK;
}
