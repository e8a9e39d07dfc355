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

// ---- -s (flag_s): per-axis MSD over the origin ring and the frame's sum_n v (x) r ------------------------------------------
//
// Reference: EvalTwoTimeEach adds dx^2, dy^2, dz^2 of every particle to rrMSR1_R[lag] (lmp2sqt.c:1434-1436; the three
// columns msdx msdy msdz of the Dt file) and EvalOneTimeSumVR forms sumVR_ct = sum_n v_n (x) r_n (:1122-1142), whose
// squared change since the origin is the stress-like rrMSR2_VR (:1461-1476).  The self-term kernels keep only the sum of
// the three squares (MSD rides as ONE column of the blocked product), so -s pays one more pass over the origin ring: 24 B
// per particle and live origin, HBM bound.  The ring is addressed through three strides so that both layouts serve
// ([slot][axis][particle] of the particle-per-lane kernels, [particle][axis][slot] of the recurrence kernel).
__global__ void __launch_bounds__(256) axis_msd_kernel(const AxisMsdArgs a)
{
	const int slot = blockIdx.y, ch = blockIdx.x;
	int tau = 0;
	if (!mma_slot_state(slot, a.p, a.T, a.B, &tau)) return;
	double acc[3] = {0.0, 0.0, 0.0};
	if (tau != 0) { // lag 0: the origin is this very frame (captured by the self-term kernel of the same frame): all zero
		const long long i0 = (long long)ch * a.chunk;
		long long i1 = i0 + a.chunk;
		if (i1 > a.n_local) i1 = a.n_local;
		const double *rs = a.ring + (long long)slot * a.s_slot;
		for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
#pragma unroll
			for (int ax = 0; ax < 3; ++ax) {
				const double d = __dsub_rn(__ldg(a.cur + ax * a.cur_pitch + i), __ldcs(rs + ax * a.s_axis + i * a.s_ptl));
				acc[ax] += __dmul_rn(d, d);
			}
		}
	}
	__shared__ double s_w[8][3];
#pragma unroll
	for (int ax = 0; ax < 3; ++ax) {
		double v = acc[ax];
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
		if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5][ax] = v;
	}
	__syncthreads();
	if (threadIdx.x < 3) {
		double t = 0.0;
		for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w][threadIdx.x];
		a.partial[((long long)slot * a.nchunk + ch) * 3 + threadIdx.x] = t;
	}
}

__global__ void axis_msd_reduce_kernel(const AxisMsdArgs a)
{
	const int slot = blockIdx.x;
	int tau = 0;
	if (!mma_slot_state(slot, a.p, a.T, a.B, &tau)) return;
	for (int ax = 0; ax < 3; ++ax) { // 32 threads, a strided share of the chunks each, then a fixed butterfly: bit-reproducible
		double v = strided_sum(a.partial + (long long)slot * a.nchunk * 3 + ax, 3, threadIdx.x, a.nchunk, 32);
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
		if (threadIdx.x == 0) a.bufMSR1[((long long)slot * a.T + tau) * 3 + ax] = v;
	}
}

int sqt_launch_axis_msd(const AxisMsdArgs &a, cudaStream_t st)
{
	dim3 grid(a.nchunk, a.B);
	axis_msd_kernel<<<grid, 256, 0, st>>>(a);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return (int)e;
	axis_msd_reduce_kernel<<<a.B, 32, 0, st>>>(a);
	return (int)cudaGetLastError();
}

__global__ void __launch_bounds__(256) sum_vr_kernel(const SumVRArgs a)
{
	double acc[9];
#pragma unroll
	for (int j = 0; j < 9; ++j) acc[j] = 0.0;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_local; i += (long long)gridDim.x * blockDim.x) {
		double r[3], v[3];
#pragma unroll
		for (int c = 0; c < 3; ++c) { r[c] = __ldg(a.cur + c * a.cur_pitch + i); v[c] = __ldg(a.cur + (3 + c) * a.cur_pitch + i); }
#pragma unroll
		for (int x = 0; x < 3; ++x)
#pragma unroll
			for (int y = 0; y < 3; ++y) acc[3 * x + y] += __dmul_rn(v[x], r[y]); // real_tensor_product_r2_r1r1(VR, vel, r), common.c:67-80
	}
	__shared__ double s_w[8][9];
#pragma unroll
	for (int j = 0; j < 9; ++j) {
		double v = acc[j];
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
		if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5][j] = v;
	}
	__syncthreads();
	if (threadIdx.x < 9) {
		double t = 0.0;
		for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w][threadIdx.x];
		a.partial[(long long)blockIdx.x * 9 + threadIdx.x] = t;
	}
}

__global__ void sum_vr_reduce_kernel(const SumVRArgs a)
{
	for (int j = 0; j < 9; ++j) {
		double v = strided_sum(a.partial + j, 9, threadIdx.x, a.nblocks, 32);
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
		if (threadIdx.x == 0) a.out[j] = v;
	}
}

int sqt_launch_sum_vr(const SumVRArgs &a, cudaStream_t st)
{
	sum_vr_kernel<<<a.nblocks, 256, 0, st>>>(a);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return (int)e;
	sum_vr_reduce_kernel<<<1, 32, 0, st>>>(a);
	return (int)cudaGetLastError();
}
