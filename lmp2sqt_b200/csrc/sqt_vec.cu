// sqt_vec.cu -- self terms for SMALL nCSpatial (<= 32: the shipped decks in.lmp2sqt.150 / .150.smallQ, BASELINE configs
// 1 and 2) with a particle on every lane.
//
// At small nCSpatial neither the recurrence kernel of sqt_kernels.cu (a thread per origin slot: every lane pays the
// ~57 FP64 instructions of displacement + three cos(2 pi u) for nCSpatial recurrence steps that are shorter than that,
// one particle per warp iteration, 6 virtual rho slots on 32 lanes) nor the DMMA product of sqt_mma.cu (its shared-memory
// staging costs more than nCSpatial multiply-adds) fits.  Here a warp owns SPW origin slots like the tensor-core kernel, a
// LANE owns a particle, and the nCSpatial running sums of F_s live in the lane's registers:
//     per (particle, origin, axis):  c = cos(2 pi u);  x_{m+1} = 2 c x_m - x_{m-1};  acc[m] += x_m      (lmp2sqt.c:1188-1196)
// -- the reference's own recurrence, three independent chains per lane, all 32 lanes busy, origin positions read from
// the particle-minor ring as one 256-byte access.  MSD, MQD and the van Hove bin are taken per (particle, origin) as in
// the other kernels.  At the end the lane sums are combined by a fixed butterfly and one row per warp and slot is left
// for vec_reduce_kernel.  rho(k) comes from rho_mma_kernel (sqt_mma.cu), which works at any nCSpatial.
#include "sqt_internal.h"
#include "sqt_lane.cuh"
#include <stdio.h>
#include <stdlib.h>

#ifdef SQT_PHASE_TIMING
// diagnostic build (make timing; profiles/probes/vec_phase.py): per-warp clocks of the kernel's phases, summed over warps
// [0] prologue [1] particle loop [2] butterfly + row store [3] fence + group counter [4] group sum (last warp only)
// [5] histogram flush [6] warps [7] last warps [8] steps [9] max whole-warp cycles
__device__ unsigned long long g_vec_tim[16];
extern "C" void sqt_debug_vec_tim(unsigned long long *out16)
{
	if (out16) cudaMemcpyFromSymbol(out16, g_vec_tim, sizeof(unsigned long long) * 16);
	unsigned long long z[16] = {0};
	cudaMemcpyToSymbol(g_vec_tim, z, sizeof(z));
}
#define VT(i) do { const long long t_ = clock64(); if (lane == 0) atomicAdd(&g_vec_tim[i], (unsigned long long)(t_ - vt_)); vt_ = t_; } while (0)
#else
#define VT(i) do {} while (0)
#endif

template <int C, int SPW, bool VARBOX, int MINB>
__global__ void __launch_bounds__(128, MINB) self_ring_vec_kernel(const MmaArgs a)
{
#ifdef SQT_PHASE_TIMING
	long long vt_ = clock64();
	const long long vt0_ = vt_;
	int vsteps_ = 0;
#endif
	extern __shared__ long long s_mem[];
	const int wpb = blockDim.x >> 5;
	const int SB = wpb * SPW;
	long long *s_thr = s_mem;                          // [nCS + 2]
	int *s_hist = (int *)(s_mem + a.nCS + 2);          // [SB][nCS]
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int gw = blockIdx.x * wpb + warp;
	const int wsg = gw % a.NWSG, kq = gw / a.NWSG;
	const int slot0 = wsg * SPW;
	const int nCS = a.nCS;

	for (int e = threadIdx.x; e < nCS * SB; e += blockDim.x) s_hist[e] = 0;
	for (int e = threadIdx.x; e < nCS + 2; e += blockDim.x) s_thr[e] = a.thr[e];

	// fused small-shape path: the slot groups beyond NWSG_real serve the six virtual rho(k) slots (axis x {cos, sin}) -- the
	// reference's own tables summed over particles (lmp2sqt.c:1176-1217), one chain per slot, no ring, no histogram
	const bool virt = wsg >= a.NWSG_real;
	const int v0 = (wsg - a.NWSG_real) * SPW;
	unsigned live_m = 0, cap_m = 0;
#pragma unroll
	for (int s = 0; s < SPW; ++s) {
		int tau = 0;
		const bool lv = virt ? (v0 + s < SQT_NVIRT) : (slot0 + s < a.B && mma_slot_state(slot0 + s, a.p, a.T, a.B, &tau));
		if (lv) live_m |= 1u << s;
		if (!virt && lv && tau == 0) cap_m |= 1u << s;
	}
	const double invL = a.invL;
	double acc[SPW][C], msd[SPW], mqd[SPW];
#pragma unroll
	for (int s = 0; s < SPW; ++s) {
		msd[s] = 0.0; mqd[s] = 0.0;
#pragma unroll
		for (int m = 0; m < C; ++m) acc[s][m] = 0.0;
	}
	__syncthreads();

	const double *cx = a.cur, *cy = cx + a.cur_pitch, *cz = cy + a.cur_pitch;
	const long long RP = a.RP;
	const int nws = (a.total_warps - wsg + a.NWSG - 1) / a.NWSG;
	// Tile hand-out in two halves so that the counter's round trip to L2 is never waited for: the request for the tile
	// after the current one is issued when a tile starts and only resolved when that tile's last step is prepared.
	int jgrab = 0;
	auto grab_issue = [&]() -> int {
		int t = 0;
		if (a.dynamic) { if (lane == 0) t = atomicAdd(&a.tile_ctr[wsg], 1); }
		else { t = kq + jgrab * nws; ++jgrab; }
		return t;
	};
	auto grab_finish = [&](int t) -> int {
		t = __reduce_max_sync(0xffffffffu, t);
		return t < a.ntiles ? t : -1;
	};

	VT(0);
	if (live_m) {
		// The loop runs over 32-particle steps.  ncu (profiles/r2_ncu_full_vec_c1.txt) showed the first version waiting on its
		// own loads (long_scoreboard 3.2 warps per issue, FP64 pipe 34 % busy at 4 warps per scheduler): every step's current
		// positions and origin rows are therefore fetched one step ahead, while the previous step's chains run.
		int pend = grab_issue();
		int tile = grab_finish(pend);
		int i0 = 0, tbase = 0, tlen = 0;
		bool have = tile >= 0;
		if (have) { tbase = tile * a.tile; tlen = a.n_local - tbase; tlen = tlen < a.tile ? tlen : a.tile; pend = grab_issue(); }
		bool pvalid = false;
		long long pn = 0;
		double prx = 0.0, pry = 0.0, prz = 0.0, pox[SPW], poy[SPW], poz[SPW];
#pragma unroll
		for (int s = 0; s < SPW; ++s) { pox[s] = 0.0; poy[s] = 0.0; poz[s] = 0.0; }
		auto load_step = [&](int tb, int tl, int i) {
			pvalid = i + lane < tl;
			pn = tb + (pvalid ? i + lane : tl - 1);
			prx = __ldg(cx + pn); pry = __ldg(cy + pn); prz = __ldg(cz + pn);
			if (!virt) {
				const double *rrow = a.ring + (long long)slot0 * 3 * RP + pn; // [slot][axis][particle]
#pragma unroll
				for (int s = 0; s < SPW; ++s)
					if (((live_m >> s) & 1) && !((cap_m >> s) & 1)) {
						const double *rs = rrow + (long long)s * 3 * RP;
						pox[s] = __ldg(rs); poy[s] = __ldg(rs + RP); poz[s] = __ldg(rs + 2 * RP);
					}
			}
		};
		if (have) load_step(tbase, tlen, 0);
		while (have) {
			const bool valid = pvalid;
			const long long n = pn;
			const double rx = prx, ry = pry, rz = prz;
			double cox[SPW], coy[SPW], coz[SPW];
#pragma unroll
			for (int s = 0; s < SPW; ++s) { cox[s] = pox[s]; coy[s] = poy[s]; coz[s] = poz[s]; }
			// the step after this one: its loads go out now
			int ni0 = i0 + 32, ntbase = tbase, ntlen = tlen;
			bool nhave = true;
			if (ni0 >= tlen) {
				const int nt = grab_finish(pend);
				if (nt < 0) nhave = false;
				else { ntbase = nt * a.tile; ntlen = a.n_local - ntbase; ntlen = ntlen < a.tile ? ntlen : a.tile; ni0 = 0; pend = grab_issue(); }
			}
			if (nhave) load_step(ntbase, ntlen, ni0);

			if (virt) {
#pragma unroll
				for (int s = 0; s < SPW; ++s) {
					if (!((live_m >> s) & 1)) continue;
					const int ax = (v0 + s) >> 1, im = (v0 + s) & 1;
					const double xa = ax == 0 ? rx : (ax == 1 ? ry : rz);
					const double u = xa * invL;
					const double c = mma_cos2pi(u);
					double x = im ? mma_cos2pi(u - 0.25) : c;  // sin(k x) / cos(k x): the seeds of lmp2sqt.c:1180-1186
					double pv = im ? 0.0 : 1.0;                 // x_0 = sin 0 / cos 0
					if (!valid) { x = 0.0; pv = 0.0; }
					const double t = c + c;
#pragma unroll
					for (int m = 0; m < C; ++m) {
						acc[s][m] += x;
						const double nx = fma(t, x, -pv);
						pv = x; x = nx;
					}
					if (!im && valid) msd[s] += xa;            // the cos slots also carry sum_n r_a (for MSDCM)
				}
			} else {
				double *rrow = a.ring + (long long)slot0 * 3 * RP + n; // [slot][axis][particle]
#pragma unroll
				for (int s = 0; s < SPW; ++s) {
					if (!((live_m >> s) & 1)) continue;
					const bool capture = (cap_m >> s) & 1;
					double *rs = rrow + (long long)s * 3 * RP;
					double ox, oy, oz;
					if (capture) {
						ox = rx; oy = ry; oz = rz;
						if (valid) { rs[0] = rx; rs[RP] = ry; rs[2 * RP] = rz; } // SetWaitedTimeCorr, lmp2sqt.c:1279-1282
					} else { ox = cox[s]; oy = coy[s]; oz = coz[s]; }
					const double dx = rx - ox, dy = ry - oy, dz = rz - oz; // lmp2sqt.c:1407-1409
					// un-fused like the reference's x86-64 build so that dr2, hence the bin, is bit-identical
					const double dr2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
					{
						int bi = lane_vanhove_bin(dr2, a.inv_rVal, nCS, s_thr);
						if (!valid) bi = nCS; // not counted
						const unsigned peers = __match_any_sync(0xffffffffu, bi);
						if (bi < nCS && lane == (__ffs(peers) - 1)) s_hist[(warp * SPW + s) * nCS + bi] += __popc(peers);
					}
					if (valid) { msd[s] += dr2; mqd[s] = fma(dr2, dr2, mqd[s]); }
					double ux = dx * invL, uy = dy * invL, uz = dz * invL;
					if (VARBOX && !capture) {
						const double dsl = invL - a.invL0[slot0 + s];
						ux = fma(ox, dsl, ux); uy = fma(oy, dsl, uy); uz = fma(oz, dsl, uz);
					}
					// three chains at once; an invalid lane runs them on zeros (x_0 = 0 keeps every x_m = 0)
					const double c0 = mma_cos2pi(ux), c1 = mma_cos2pi(uy), c2 = mma_cos2pi(uz);
					const double one = valid ? 1.0 : 0.0;
					double p0 = one, p1 = one, p2 = one;                                  // x_0 = T_0
					double x0 = valid ? c0 : 0.0, x1 = valid ? c1 : 0.0, x2 = valid ? c2 : 0.0; // x_1 = T_1
					const double t0 = c0 + c0, t1 = c1 + c1, t2 = c2 + c2;
#pragma unroll
					for (int m = 0; m < C; ++m) {
						acc[s][m] += (x0 + x1) + x2;
						const double n0 = fma(t0, x0, -p0), n1 = fma(t1, x1, -p1), n2 = fma(t2, x2, -p2);
						p0 = x0; p1 = x1; p2 = x2;
						x0 = n0; x1 = n1; x2 = n2;
					}
				}
			}
			i0 = ni0; tbase = ntbase; tlen = ntlen; have = nhave;
#ifdef SQT_PHASE_TIMING
			++vsteps_;
#endif
		}
	}
	VT(1);

	// lane sums -> warp sums by a fixed butterfly; partial rows [slot group][k-th warp][slot][C + 2]
	if (kq < a.KW) {
		double *pb = a.partial + ((long long)(wsg * a.KW + kq) * SPW) * (C + 2);
#pragma unroll
		for (int s = 0; s < SPW; ++s) {
#pragma unroll
			for (int m = 0; m < C + 2; ++m) {
				double v = m < C ? acc[s][m < C ? m : 0] : (m == C ? msd[s] : mqd[s]);
#pragma unroll
				for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
				if (lane == (m & 31)) pb[s * (C + 2) + m] = v;
			}
		}
	}
	VT(2);
	if (a.done_ctr) {
		// fused small-shape path: whichever warp of this slot group posts its rows last also sums the group's rows -- in row
		// order, so the result does not depend on which warp that is -- and writes them to their final place: F_s / MSD / MQD
		// of the slot's current lag (TBuf.F_s_qq2, rrMSD, rrMQD), or rho(k) and sum_n r_a for a virtual group
		__threadfence();
		int last = 0;
		if (lane == 0) last = atomicAdd(&a.done_ctr[wsg], 1) == nws - 1;
		last = __shfl_sync(0xffffffffu, last, 0);
		VT(3);
		if (last) {
			__threadfence();
			const long long stride = (long long)SPW * (C + 2);
#pragma unroll
			for (int s = 0; s < SPW; ++s) {
				const double *p0 = a.partial + ((long long)(wsg * a.KW) * SPW + s) * (C + 2);
				if (virt) {
					const int v = v0 + s;
					if (v >= SQT_NVIRT) continue;
					const int ax = v >> 1, im = v & 1;
					for (int m = lane; m < nCS; m += 32) a.rho_cur[(2 * nCS) * ax + 2 * m + im] = strided_sum(p0 + m, stride, 0, nws, 1); // marker layout lmp2sqt.c:1173-1174
					if (!im && lane == 0) a.rho_cur[6 * nCS + ax] = strided_sum(p0 + C, stride, 0, nws, 1);
				} else {
					const int slot = slot0 + s;
					int tau = 0;
					if (slot >= a.B || !mma_slot_state(slot, a.p, a.T, a.B, &tau)) continue;
					for (int m = lane; m < C + 2; m += 32) {
						const double v = strided_sum(p0 + m, stride, 0, nws, 1);
						if (m == C) a.bufMSD[slot * a.T + tau] = v;
						else if (m == C + 1) a.bufMQD[slot * a.T + tau] = v;
						else if (m < nCS) a.bufFs[((long long)slot * a.T + tau) * nCS + m] = v;
					}
				}
			}
#ifdef SQT_PHASE_TIMING
			VT(4);
			if (lane == 0) atomicAdd(&g_vec_tim[7], 1ull);
#endif
		}
	}
	__syncthreads();
	for (int e = threadIdx.x; e < nCS * SB; e += blockDim.x) {
		const int cntv = s_hist[e];
		if (cntv) {
			const int col = e / nCS;
			const int gwc = blockIdx.x * wpb + col / SPW;
			const int sl = (gwc % a.NWSG) * SPW + col % SPW; // (virtual slot groups never count: cntv is 0 there)
			atomicAdd(&a.hist_cur[sl * nCS + e % nCS], cntv);
		}
	}
#ifdef SQT_PHASE_TIMING
	VT(5);
	if (lane == 0) {
		atomicAdd(&g_vec_tim[6], 1ull);
		atomicAdd(&g_vec_tim[8], (unsigned long long)vsteps_);
		atomicMax(&g_vec_tim[9], (unsigned long long)(clock64() - vt0_));
	}
#endif
}

// one block per origin slot: fixed-order sum of the per-warp rows, scatter to the TBuf layout (F_s_qq2 / rrMSD / rrMQD /
// DrTable), re-arm the histogram and the tile counters
__global__ void vec_reduce_kernel(const MmaReduceArgs a)
{
	const int slot = blockIdx.x;
	const int W = a.NT; // row width C + 2 (the field is reused)
	int tau = 0;
	const bool live = slot < a.B && mma_slot_state(slot, a.p, a.T, a.B, &tau);
	if (slot < a.B) {
		const int SPW = a.SPW, wsg = slot / SPW, s = slot % SPW;
		if (live) {
			for (int m = threadIdx.x; m < W; m += blockDim.x) {
				const double *p = a.partial + ((long long)(wsg * a.KW) * SPW + s) * W + m;
				const long long stride = (long long)SPW * W;
				const double v = strided_sum(p, stride, 0, a.KW, 1);
				if (m == W - 2) a.bufMSD[slot * a.T + tau] = v;
				else if (m == W - 1) a.bufMQD[slot * a.T + tau] = v;
				else if (m < a.nCS) a.bufFs[((long long)slot * a.T + tau) * a.nCS + m] = v; // (fewer chains than nCSpatial: F_s not wanted)
			}
			if (threadIdx.x == 0 && tau == 0) a.invL0[slot] = a.invL;
			for (int bin = threadIdx.x; bin < a.nCS; bin += blockDim.x)
				a.bufHist[((long long)slot * a.T + tau) * a.nCS + bin] = a.hist_cur[slot * a.nCS + bin];
		}
		__syncthreads();
		for (int bin = threadIdx.x; bin < a.nCS; bin += blockDim.x) a.hist_cur[slot * a.nCS + bin] = 0;
	}
	if (slot == 0)
		for (int i = threadIdx.x; i < a.NWSG; i += blockDim.x) a.tile_ctr[i] = 0;
}

// ---------------------------------------------------------------------------------------------
// vec_finish_kernel: everything after the particle loop of a frame in ONE launch, for the small shapes (shipped decks:
// 4,000 / 32,000 atoms, nCSpatial 30 / 10) whose frames are a few tens of microseconds and were dominated by the launch
// latency of five tiny kernels.  One block per origin slot: it moves the slot's van Hove counts to their lag, does for its
// slot what collective_ring_kernel does, and AccumSpacetimeCorr (lmp2sqt.c:368-444) for the buffer that completes in this
// frame is shared by all blocks.  (The sums over the per-warp rows -- F_s, MSD, MQD, rho(k) -- are made by the last warp of
// each slot group inside self_ring_vec_kernel.)  Only without a communicator: sharded runs need the allreduces in between.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) vec_finish_kernel(const VecFinishArgs a)
{
	extern __shared__ double s_fin[]; // rho [6 nCS] + sum r [3]
	const MmaReduceArgs &r = a.r;
	const int slot = blockIdx.x;
	const int nCS = r.nCS, T = r.T;
	double *s_rho = s_fin;
	int tau = 0;
	const bool live = mma_slot_state(slot, r.p, r.T, r.B, &tau);
	// rho(k), sum_n r_a, F_s, MSD and MQD of this frame are final already: the last warp of every slot group of
	// self_ring_vec_kernel summed the group's rows
	for (int e = threadIdx.x; e < 6 * nCS + 3; e += blockDim.x) s_rho[e] = a.rho_cur[e];
	if (live) {
		if (threadIdx.x == 0 && tau == 0) r.invL0[slot] = r.invL;
		for (int bin = threadIdx.x; bin < nCS; bin += blockDim.x) {
			r.bufHist[((long long)slot * T + tau) * nCS + bin] = r.hist_cur[slot * nCS + bin];
			r.hist_cur[slot * nCS + bin] = 0;
		}
	} else
		for (int bin = threadIdx.x; bin < nCS; bin += blockDim.x) r.hist_cur[slot * nCS + bin] = 0;
	if (slot == 0)
		for (int i = threadIdx.x; i < r.NWSG; i += blockDim.x) { r.tile_ctr[i] = 0; a.done_ctr[i] = 0; } // re-arm the scheduler
	__syncthreads();
	// AccumSpacetimeCorr (lmp2sqt.c:368-444) for the buffer that completes in this frame, part 1: its lags < T - 1 were
	// written in earlier frames, so all blocks share them (one element per thread); lag T - 1 follows below, by the block
	// that writes it
	const int nbd = a.nb_done;
	if (nbd >= 0) {
		const int TS = T * nCS;
		const long long base = (long long)nbd * TS;
		for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < TS; idx += gridDim.x * blockDim.x) {
			const int nk = idx / T, nt = idx % T, src = nt * nCS + nk; // output [nk][nt] like avF_qq2, buffers [lag][nk]
			if (nt == T - 1 || !a.flag_f) continue;
			const double F = a.bufF[base + src];
			a.av[0][idx] += F; a.av[1][idx] += F * F;
			if (a.flag_F) {
				const double Fs = r.bufFs[base + src];
				const double Fd = 2.0 * (F - Fs);
				a.av[2][idx] += Fs; a.av[3][idx] += Fs * Fs;
				a.av[4][idx] += Fd; a.av[5][idx] += Fd * Fd;
			}
			if (a.flag_t) a.av[6][idx] += (double)r.bufHist[base + src];
		}
		if (a.flag_t && blockIdx.x == gridDim.x - 1)
			for (int idx = threadIdx.x; idx < T - 1; idx += blockDim.x) {
				a.av1[0][idx] += r.bufMSD[nbd * T + idx];
				a.av1[1][idx] += r.bufMQD[nbd * T + idx];
				a.av1[2][idx] += a.bufMSDCM[nbd * T + idx];
			}
	}
	if (!live) return;
	// collective_ring_kernel's part for this slot (lmp2sqt.c:1584-1586, 1298-1299, 1447-1450)
	const bool capture = tau == 0;
	for (int m = threadIdx.x; m < nCS; m += blockDim.x) {
		double F = 0.0;
		double *r0 = a.rho0 + (long long)slot * 6 * nCS;
		for (int ax = 0; ax < 3; ++ax) {
			const int mR = (2 * nCS) * ax + 2 * m, mI = mR + 1;
			const double R = s_rho[mR], I = s_rho[mI];
			if (capture) { r0[mR] = R; r0[mI] = I; }
			const double R0 = capture ? R : r0[mR], I0 = capture ? I : r0[mI];
			F = __dadd_rn(F, __dadd_rn(__dmul_rn(R, R0), __dmul_rn(I, I0)));
		}
		a.bufF[((long long)slot * T + tau) * nCS + m] = F;
	}
	if (threadIdx.x == 0) {
		const double *sr = s_rho + 6 * nCS;
		double *s0 = a.sumr0 + slot * 3;
		if (capture) { s0[0] = sr[0]; s0[1] = sr[1]; s0[2] = sr[2]; }
		const double ex = capture ? 0.0 : sr[0] - s0[0], ey = capture ? 0.0 : sr[1] - s0[1], ez = capture ? 0.0 : sr[2] - s0[2];
		double d = __dmul_rn(ex, ex);
		d = __dadd_rn(d, __dmul_rn(ey, ey));
		d = __dadd_rn(d, __dmul_rn(ez, ez));
		a.bufMSDCM[slot * T + tau] += d * a.inv_n2; // never zeroed: cumulative over buffer reuse, like the reference
	}
	if (slot != nbd) return;
	// part 2: the last lag of the completing buffer, written by this very block just above
	__syncthreads();
	const int TS = T * nCS, nt = T - 1;
	const long long base = (long long)slot * TS;
	if (a.flag_f)
		for (int nk = threadIdx.x; nk < nCS; nk += blockDim.x) {
			const int idx = nk * T + nt, src = nt * nCS + nk;
			const double F = a.bufF[base + src];
			a.av[0][idx] += F; a.av[1][idx] += F * F;
			if (a.flag_F) {
				const double Fs = r.bufFs[base + src];
				const double Fd = 2.0 * (F - Fs);
				a.av[2][idx] += Fs; a.av[3][idx] += Fs * Fs;
				a.av[4][idx] += Fd; a.av[5][idx] += Fd * Fd;
			}
			if (a.flag_t) a.av[6][idx] += (double)r.bufHist[base + src];
		}
	if (a.flag_t && threadIdx.x == 0) {
		a.av1[0][nt] += r.bufMSD[slot * T + nt];
		a.av1[1][nt] += r.bufMQD[slot * T + nt];
		a.av1[2][nt] += a.bufMSDCM[slot * T + nt];
	}
}

int sqt_launch_vec_finish(const VecFinishArgs &a, cudaStream_t st)
{
	vec_finish_kernel<<<a.r.B, 128, sizeof(double) * (6 * a.r.nCS + 3), st>>>(a);
	return (int)cudaGetLastError();
}

typedef void (*vec_fn)(const MmaArgs);
// MINB = resident blocks per SM the register allocation is bounded for (launch bounds)
static vec_fn vec_variant(int C, int SPW, int varbox, int minb)
{
#define VV(C_, SPW_, M_) if (C == C_ && SPW == SPW_ && minb == M_) return varbox ? (vec_fn)self_ring_vec_kernel<C_, SPW_, true, M_> : (vec_fn)self_ring_vec_kernel<C_, SPW_, false, M_>;
	VV(4, 2, 5) VV(8, 2, 5) VV(12, 2, 4) VV(12, 2, 5)
	VV(4, 1, 8) VV(8, 1, 8) VV(12, 1, 6) VV(12, 1, 8) VV(12, 1, 5) VV(12, 1, 4) VV(8, 1, 6)
	VV(16, 1, 5) VV(20, 1, 5) VV(24, 1, 4) VV(28, 1, 4) VV(32, 1, 4) VV(32, 1, 3)
	VV(16, 1, 4) VV(20, 1, 4) VV(24, 1, 3) VV(28, 1, 3) VV(32, 1, 2) VV(4, 1, 6)
#undef VV
	return nullptr;
}
// measured on B200 (profiles/README.md, small shapes): C = 32 needs 168 registers to keep its 32 running sums and the
// prefetched step out of local memory (3 blocks per SM beat 4 spilling ones, 30.9 vs 37.7 us at config 1); C <= 12 runs best
// with ONE slot per warp at 80 registers, 6 blocks per SM (91 vs 98 us at config 2; 8 blocks at 64 registers spill: 103 us)
// One slot per warp: the largest block count whose register budget keeps the running sums and the prefetched step out of
// local memory (cuobjdump -res-usage: no stack) -- spills sit in the chain loop and cost more than the lost warps
// (config 2, C = 12: 80 registers x 6 blocks with 48 B of stack 90.7 us per kernel, 94 x 5 without 80.3 us).
static int vec_default_minb(int C, int SPW)
{
	if (SPW != 1) return C * SPW > 20 ? 4 : 5;
	return C <= 4 ? 8 : (C <= 8 ? 6 : (C <= 12 ? 5 : (C <= 20 ? 4 : 3)));
}

static int vec_env_int(const char *name, int dflt)
{
	const char *s = getenv(name);
	return (s && *s) ? atoi(s) : dflt;
}

int sqt_pick_vec(int nCS, int nChain, int B, int n_local, int n_sm, int fuse, VecChoice *vc)
{
	vc->on = 0; vc->fused = 0;
	const int want = vec_env_int("SQT_VEC", -1);
	if (want == 0 || B < 1 || nCS < 1 || nChain < 1 || nChain > 32) return 0;
	const int C = (nChain + 3) / 4 * 4;
	int SPW = 1;               // origin slots per warp (2 C SPW registers of running sums per lane)
	if (C <= 12 && vec_env_int("SQT_VEC_SPW", 0) == 2) SPW = 2;
	int minb = vec_env_int("SQT_VEC_MINB", vec_default_minb(C, SPW));
	if (!vec_variant(C, SPW, 0, minb)) minb = vec_default_minb(C, SPW);
	const int W = 4;
	const size_t sm = sizeof(long long) * (nCS + 2) + sizeof(int) * (size_t)nCS * W * SPW;
	for (int vb = 0; vb < 2; ++vb) {
		cudaError_t e = cudaFuncSetAttribute((const void *)vec_variant(C, SPW, vb, minb), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
		if (e != cudaSuccess) return (int)e;
	}
	int occ = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)vec_variant(C, SPW, 0, minb), W * 32, sm);
	if (e != cudaSuccess) return (int)e;
	if (occ < 1) return 0;
	cudaFuncAttributes fa;
	cudaFuncGetAttributes(&fa, (const void *)vec_variant(C, SPW, 0, minb));
	vc->C = C; vc->SPW = SPW; vc->minb = minb; vc->WPB = W; vc->occ = occ; vc->regs = fa.numRegs; vc->smem = sm;
	vc->NWSG_real = (B + SPW - 1) / SPW;
	vc->fused = fuse && C >= nCS && vec_env_int("SQT_FUSE", 1) != 0; // the virtual rho slots need all nCSpatial chains
	vc->NWSG = vc->NWSG_real + (vc->fused ? (SQT_NVIRT + SPW - 1) / SPW : 0);
	int blocks = n_sm * occ;
	long long warps_per_group = ((long long)blocks * W + vc->NWSG - 1) / vc->NWSG;
	long long ts = n_local / (4 * warps_per_group); // about 4 tiles per warp, a multiple of 32, 32 .. 256 particles
	ts = (ts / 32) * 32;
	if (ts < 32) ts = 32;
	if (ts > 128) ts = 128;    // (the hand-out of the next tile is requested a tile ahead, so short tiles cost nothing)
	const int fT = vec_env_int("SQT_TILE", 0);
	if (fT > 0) ts = ((fT + 31) / 32) * 32;
	vc->tile = (int)ts;
	vc->ntiles = (int)((n_local + ts - 1) / ts);
	long long need = (long long)vc->ntiles * vc->NWSG;
	if ((long long)blocks * W > need) blocks = (int)((need + W - 1) / W);
	if (blocks < 1) blocks = 1;
	vc->blocks = blocks;
	vc->total_warps = blocks * W;
	vc->KW = (vc->total_warps + vc->NWSG - 1) / vc->NWSG;
	vc->dynamic = vec_env_int("SQT_DETERMINISTIC", 0) ? 0 : 1;
	vc->on = 1;
	return 0;
}

int sqt_launch_vec(const VecChoice &vc, const MmaArgs &a, cudaStream_t st)
{
	vec_fn fn = vec_variant(vc.C, vc.SPW, a.varbox, vc.minb);
	if (!fn) return -1;
	fn<<<vc.blocks, vc.WPB * 32, vc.smem, st>>>(a);
	return (int)cudaGetLastError();
}

int sqt_launch_vec_reduce(const MmaReduceArgs &a, cudaStream_t st)
{
	vec_reduce_kernel<<<a.B, 64, 0, st>>>(a);
	return (int)cudaGetLastError();
}
