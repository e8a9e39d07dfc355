// sqt_vel.cu -- velocity autocorrelation over the origin ring (-V / flag_velocity).
//
// Reference: EvalTwoTimeEach, lmp2sqt.c:1397-1401 and :1441 -- per live buffer and lag, rrCvv += sum_n v_n(t) . v_n(t0),
// with the origin velocities orgV captured by SetWaitedTimeCorr (:1283-1287).  Here the origin velocities live in a
// second particle-minor ring [slot][axis][particle] next to the position ring; one pass per frame reads it once
// (24 B per particle and live origin: HBM bound, K-independent) and leaves one partial sum per (slot, particle chunk),
// which vel_reduce_kernel adds in a fixed order.  The k-space current correlators of -V (C2_v_v, C2_rho_v, C2_v_rho)
// need no per-particle history: they come from the weighted pass of rho_mma_kernel and collective_ring_kernel.
#include "sqt_internal.h"
#include "sqt_lane.cuh"

__global__ void __launch_bounds__(256) vel_ring_kernel(const VelArgs a)
{
	const int slot = blockIdx.y, ch = blockIdx.x;
	int tau = 0;
	if (!mma_slot_state(slot, a.p, a.T, a.B, &tau)) return;
	const bool capture = tau == 0;
	const double *vx = a.vcur, *vy = vx + a.cur_pitch, *vz = vy + a.cur_pitch;
	double *rx = a.vring + (long long)slot * 3 * a.RP, *ry = rx + a.RP, *rz = ry + a.RP;
	const long long i0 = (long long)ch * a.chunk;
	long long i1 = i0 + a.chunk;
	if (i1 > a.n_local) i1 = a.n_local;
	double acc0 = 0.0, acc1 = 0.0;
	long long i = i0 + threadIdx.x;
	for (; i + blockDim.x < i1; i += 2 * blockDim.x) { // two independent chains: the loads of both are in flight together
		const long long j = i + blockDim.x;
		const double ax = __ldg(vx + i), ay = __ldg(vy + i), az = __ldg(vz + i);
		const double bx = __ldg(vx + j), by = __ldg(vy + j), bz = __ldg(vz + j);
		double ox, oy, oz, px, py, pz;
		if (capture) { // SetWaitedTimeCorr, lmp2sqt.c:1283-1287
			ox = ax; oy = ay; oz = az; px = bx; py = by; pz = bz;
			rx[i] = ax; ry[i] = ay; rz[i] = az; rx[j] = bx; ry[j] = by; rz[j] = bz;
		} else {
			ox = __ldcs(rx + i); oy = __ldcs(ry + i); oz = __ldcs(rz + i);
			px = __ldcs(rx + j); py = __ldcs(ry + j); pz = __ldcs(rz + j);
		}
		// un-fused like the reference's x86-64 build (lmp2sqt.c:1398-1400)
		acc0 += __dadd_rn(__dadd_rn(__dmul_rn(ax, ox), __dmul_rn(ay, oy)), __dmul_rn(az, oz));
		acc1 += __dadd_rn(__dadd_rn(__dmul_rn(bx, px), __dmul_rn(by, py)), __dmul_rn(bz, pz));
	}
	if (i < i1) {
		const double ax = __ldg(vx + i), ay = __ldg(vy + i), az = __ldg(vz + i);
		double ox = ax, oy = ay, oz = az;
		if (capture) { rx[i] = ax; ry[i] = ay; rz[i] = az; }
		else { ox = __ldcs(rx + i); oy = __ldcs(ry + i); oz = __ldcs(rz + i); }
		acc0 += __dadd_rn(__dadd_rn(__dmul_rn(ax, ox), __dmul_rn(ay, oy)), __dmul_rn(az, oz));
	}
	double v = acc0 + acc1;
	__shared__ double s_w[8];
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x == 0) {
		double t = 0.0;
		for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
		a.partial[(long long)slot * a.nchunk + ch] = t;
	}
}

__global__ void vel_reduce_kernel(const VelArgs a)
{
	const int slot = blockIdx.x;
	int tau = 0;
	if (!mma_slot_state(slot, a.p, a.T, a.B, &tau)) return;
	// 32 threads, each a strided share of the chunks, then a fixed butterfly: bit-reproducible
	double v = strided_sum(a.partial + (long long)slot * a.nchunk, 1, threadIdx.x, a.nchunk, 32);
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
	if (threadIdx.x == 0) a.bufCvv[slot * a.T + tau] = v;
}

int sqt_launch_vel(const VelArgs &a, cudaStream_t st)
{
	dim3 grid(a.nchunk, a.B);
	vel_ring_kernel<<<grid, 256, 0, st>>>(a);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return (int)e;
	vel_reduce_kernel<<<a.B, 32, 0, st>>>(a);
	return (int)cudaGetLastError();
}
