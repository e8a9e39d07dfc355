// sqt_internal.h -- shared declarations between the C-ABI layer (sqt_api.cu), the kernels
// (sqt_kernels.cu), the NCCL shim (sqt_comm.cpp) and the synthetic source (sqt_synth.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SQT_NVIRT 6          // virtual origin slots: (axis x,y,z) x (Re, Im) of rho(k)
// running sums on the device (d_avblk): SQT_N2D arrays [nCS][T], then SQT_N1D arrays [T]
//   2-D: 0 avF 1 sdF 2 avFs 3 sdFs 4 avFd 5 sdFd 6 avHist 7 C2_v_rho 8 C2_rho_v 9 C2_v_v 10 C2_v_mu 11 C2_mu_v 12 C2_mu_rho 13 C2_rho_mu 14 C2_mu_mu
//   1-D: 0 msd 1 mqd 2 msdcm 3 cvv 4 cvcmvcm 5 cmm (never added to, like the reference) 6-8 MSR1_R x y z 9-17 MSR2_VR xx..zz
#define SQT_N2D 15
#define SQT_N1D 18
#define SQT_NC2 8            // planes of bufC2, in the order of 2-D arrays 7..14
#define SQT_MAX_TYPES 8      // per-type partials: species a = 1..n_types
// with n_types > 1 the running sums continue behind the SQT_N1D arrays:
//   F_ab [n_types^2] x [nCS][T], F_s,a [n_types] x [nCS][T], van Hove_a [n_types] x [nCS][T], MSD_a [n_types] x [T], MQD_a [n_types] x [T]
#define SQT_MAX_WPB 4        // warps per block of the ring kernel

// Everything the per-frame ring kernel needs.  Passed by value.
struct RingArgs {
	const double *cur;        // current frame, SoA [3][cur_pitch]
	double *ring;             // origin positions [n][3][RP]  (slot-minor: a warp reads 32 origins of one particle/axis contiguously)
	double *partial;          // per-warp partial sums [slot group][k-th warp][NV][O]
	int *tile_ctr;            // [NWSG] next particle tile per slot group (dynamic scheduling)
	int *hist_cur;            // this frame's van Hove counts [B][nCS]   (zeroed by the reduce kernel)
	const double *invL0;      // [B] 1/L of each buffer's origin frame
	const long long *thr;     // [nCS+2] bit patterns of the exact bin thresholds (see host: build_thresholds)
	long long cur_pitch;
	int n_local;
	int tile, ntiles;         // particles per tile, tiles in the shard
	int NWSG, KW, total_warps; // slot groups, warps per slot group (upper bound), warps launched
	int dynamic;              // 1: tiles handed out by atomic counter, 0: static round robin (bit-reproducible)
	int p;                    // processed-frame index within the current input file
	int T, B, nCS, RP;
	int V0;                   // first virtual slot (B rounded up to even)
	int SPAD, NV, CT;         // padded slot count, values per slot (CT + 2 G), padded chain count (C*G)
	double invL;              // 1 / L of the current frame
	float inv_rVal;
	int varbox;               // some live origin has a different box length than the current frame
};

struct KernelChoice {
	int C, G, minb;           // chains per thread, lane groups per warp, launch-bounds variant
	int WPB;                  // warps per block
	int blocks;               // persistent 1-D grid
	int NWSG, KW, total_warps;
	int tile, ntiles, dynamic;
	int SPAD, NV, CT, NSLOT, V0;
	size_t smem;
	int regs;                 // registers per thread reported by the runtime
	int occ;                  // resident blocks per SM
};

// ---- blocked-Chebyshev / FP64 tensor-core variant of the ring kernel (large nCSpatial) ---------------
// m = S*b + r (b < 8, r < S):  sum_n T_m = 2 G[b][r] - sum_n T_{|S b - r|},  G[b][r] = sum_n T_{S b}(c_n) T_r(c_n)
// G is an (8 x 3N)(3N x 8 NT) product per origin slot, accumulated by mma.sync m8n8k4 f64.
struct MmaArgs {
	const double *cur;        // current frame, SoA [3][cur_pitch]
	double *ring;             // origin positions [slot][3][RP], RP = particle pitch
	double *partial;          // per-warp G fragments [slot group][k-th warp][SPW][NT][32 lanes][2]
	double *extra;            // SPOUT: per-warp (MSD, MQD) sums [slot group][k-th warp][SPW][2]
	int *tile_ctr;            // [NWSG]
	int *hist_cur;            // [B][nCS]
	const double *invL0;      // [B]
	const long long *thr;     // [nCS+2]
	long long cur_pitch;
	int n_local, tile, ntiles, NWSG, KW, total_warps, dynamic;
	int NWSG_real;            // vector kernel, fused small-shape path: slot groups >= NWSG_real serve the six virtual rho(k) slots
	// fused small-shape path only (done_ctr != nullptr): the LAST warp of a slot group to finish sums the group's per-warp
	// rows (fixed order, whichever warp it is) and writes them where the reference keeps them
	int *done_ctr;            // [NWSG] warps of the group that have posted their rows (re-armed by vec_finish_kernel)
	double *bufFs, *bufMSD, *bufMQD, *rho_cur;
	int p, T, B, nCS;
	long long RP;
	int S, NT;                // stride of the block decomposition, column tiles (8 NT >= S + 2)
	double invL;
	float inv_rVal;
	int varbox;
};
struct MmaChoice {
	int on;                   // 1: use the mma path for the real slots (the classic kernel then serves only rho)
	int geo;                  // 1: S / NT / SPOUT are valid (the rho kernel can run) even if on == 0
	int S, NT, SPW, SPOUT, WPB, blocks, occ, regs, NWSG, KW, total_warps, tile, ntiles, dynamic; // SPW = origin slots per warp, SPOUT = MSD/MQD summed outside the product
	size_t smem;
};
int sqt_pick_mma(int nCS, int B, int n_local, int n_sm, MmaChoice *mc);
int sqt_launch_mma(const MmaChoice &mc, const MmaArgs &a, cudaStream_t st);
struct MmaReduceArgs {
	const double *partial, *extra;
	int KW, NWSG, S, NT, SPW, spout, n_local;
	int p, T, B, nCS;
	int *tile_ctr;
	int *hist_cur, *bufHist;
	double *bufFs, *bufMSD, *bufMQD, *invL0;
	double invL;
};
int sqt_launch_mma_reduce(const MmaReduceArgs &a, cudaStream_t st);
// ---- sqt_vec.cu: particle-per-lane kernel for small nCSpatial (<= 32); takes MmaArgs / MmaReduceArgs (NT = row width C + 2)
struct VecChoice {
	int on;
	int C, SPW, WPB, blocks, occ, regs, NWSG, KW, total_warps, tile, ntiles, dynamic; // C = nCSpatial rounded up to 4
	int minb;                 // launch-bounds variant (resident blocks per SM the registers are bounded for)
	int fused, NWSG_real;     // fused: rho(k) rides in the same kernel (six virtual slots) and ONE kernel finishes the frame
	size_t smem;
};
// nChain = recurrence steps per (particle, origin, axis): nCSpatial, or 1 when F_s is not wanted (-f / -t without -F): the
// kernel then only serves MSD, MQD and the van Hove histogram, whatever nCSpatial is
int sqt_pick_vec(int nCS, int nChain, int B, int n_local, int n_sm, int fuse, VecChoice *vc);
// the fused frame epilogue of the small-shape path (no communicator): per origin slot the fixed-order sum of the per-warp
// rows, rho(k) from the virtual slots' rows, the F(q,t) update / origin copies / MSDCM of collective_ring_kernel, and
// AccumSpacetimeCorr for the buffer that completes in this frame
struct VecFinishArgs {
	MmaReduceArgs r;          // as for vec_reduce_kernel (NT = row width C + 2)
	int NWSG_real;
	int *done_ctr;
	double *rho_cur, *rho0, *sumr0, *bufF, *bufMSDCM;
	double inv_n2;
	int nb_done;              // buffer completing in this frame (-1: none)
	int flag_f, flag_F, flag_t;
	double *av[7], *av1[3];
};
int sqt_launch_vec_finish(const VecFinishArgs &a, cudaStream_t st);
int sqt_launch_vec(const VecChoice &vc, const MmaArgs &a, cudaStream_t st);
int sqt_launch_vec_reduce(const MmaReduceArgs &a, cudaStream_t st);

struct RhoMmaArgs {
	const double *cur;        // current frame, SoA [3][cur_pitch]
	const double *wt;         // nullptr: rho(k).  Else per-particle weights, SoA [3][cur_pitch] (the velocities): the pass then
	                          // yields the current density kvel_q1 (lmp2sqt.c:1203-1208) = m kVal sum_n v_a (-sin, +cos)(m k0 x_a)
	double kfac;              // kVal of the current frame (weighted pass only)
	double *partial;          // per-block fragments [block][axis][cos|sin][NT][32 lanes][2]
	double *rho_cur;          // [3][nCS][Re,Im] then sum_n r_a [3]   (weighted pass: kvel, then sum_n v_a)
	long long cur_pitch;
	double *extra;            // SPOUT: per-block sum_n r_a [block][3]
	int n_local, total_warps, n_sets, nCS, S, NT, spout; // n_sets = blocks = partial sets
	double invL;
};
struct RhoMmaChoice { int on, blocks, WPB, total_warps, occ; size_t smem; };
int sqt_pick_rho_mma(const MmaChoice &mc, int n_local, int n_sm, RhoMmaChoice *rc); // needs only mc.S / NT / SPOUT (valid whenever mc.geo)
int sqt_launch_rho_mma(const MmaChoice &mc, const RhoMmaChoice &rc, const RhoMmaArgs &a, cudaStream_t st);

// ---- kernels.cu ---------------------------------------------------------------------------
// Picks (C, G, grid) for a deck; returns 0 or a cudaError_t.
int sqt_pick_kernel(int nCS, int B, int n_local, int n_sm, KernelChoice *kc);
int sqt_launch_ring(const KernelChoice &kc, const RingArgs &a, cudaStream_t st);

struct ReduceArgs {
	const double *partial;
	int KW, NWSG, NV, SPAD, CT, V0, G;
	int *tile_ctr;
	int p, T, B, nCS;
	int *hist_cur;            // [B][nCS] -> moved into bufHist, then zeroed
	int *bufHist;             // [B][T][nCS]
	double *bufFs;            // [B][T][nCS]
	double *bufMSD, *bufMQD;  // [B][T]
	double *rho_cur;          // [6 nCS + 3]: rho (marker layout) then sum of positions
	double *invL0;            // [B]
	double invL;
};
int sqt_launch_reduce(const ReduceArgs &a, cudaStream_t st);

struct CollectiveArgs {
	const double *rho_cur;    // reduced over shards: rho [6 nCS], sum r [3]; with -V then kvel [6 nCS], sum v [3]
	double *rho0;             // [B][6 nCS]
	double *sumr0;            // [B][3]
	double *bufF;             // [B][T][nCS]
	double *bufMSDCM;         // [B][T]   (never zeroed: cumulative, lmp2sqt.c:1450)
	int p, T, B, nCS;
	double inv_n2;            // 1 / N^2
	// -V (flag_velocity): current correlators, lmp2sqt.c:1542-1555, 1452-1457
	int flag_V;
	double *kvel0;            // [B][6 nCS]
	double *sumv0;            // [B][3]
	double *bufC2;            // [SQT_NC2][B][T][nCS]: C2_v_rho, C2_rho_v, C2_v_v, C2_v_mu, C2_mu_v, C2_mu_rho, C2_rho_mu, C2_mu_mu
	double *bufCvcm;          // [B][T]
	// -M (flag_magnet): the dipole-weighted density kmu_q1 and its correlators, lmp2sqt.c:1556-1576
	int flag_M;
	double *kmu0;             // [B][6 nCS]
	int rho_n;                // 6 nCS + 3: stride of the field blocks of rho_cur
	int blk_v, blk_m;         // block index of kvel / kmu inside rho_cur (block 0 = rho)
	// -s (flag_s): sumVR_ct = sum_n v (x) r of the frame (9 doubles at rho_cur + off_vr, reduced over shards),
	// orgSumVR, rrMSR2_VR (lmp2sqt.c:1122-1142, 1273-1275, 1461-1476)
	int flag_s;
	int off_vr;
	int vr_cumulative;        // no -f: the reference never zeroes sumVR_ct (ZeroOneTimeCorr runs under -f only, :1246-1249), it keeps growing
	int vr_parity;            // sumVR_state[vr_parity] is this frame's value, [1 - vr_parity] the previous frame's
	double *sumVR_state;      // [2][9]
	double *orgSumVR;         // [B][9]
	double *bufMSR2;          // [B][T][9]
	// per-type partials: rho_a(k) of the frame in blocks 1 .. n_types of rho_cur (block 0 = their sum), origin copies,
	// F_ab[lag] = sum over axes of rho_a(t) . rho_b*(t0)
	int n_types;              // 0 / 1: off
	double *rho0_t;           // [B][n_types][6 nCS]
	double *bufFab;           // [n_types * n_types][B][T][nCS], plane a * n_types + b
};
struct TypeSumArgs { double *rho_cur; int rho_n, n_types; }; // block 0 = sum of blocks 1 .. n_types, fixed order
int sqt_launch_type_sum(const TypeSumArgs &a, cudaStream_t st);
// frame in dump order -> grouped by species (each species' range starts at a multiple of 32 particles): dst[c][i] = src[c][perm[i]]
int sqt_launch_gather(double *dst, long long dst_pitch, const double *src, long long src_pitch, const int *perm, int n_sorted, int ncomp, cudaStream_t st);
int sqt_launch_collective(const CollectiveArgs &a, cudaStream_t st);

struct AccumArgs {
	// completed buffer nb
	int nb, T, nCS;
	const double *bufF, *bufFs, *bufMSD, *bufMQD, *bufMSDCM;
	const int *bufHist;
	double *cblk;             // packed block [Fs T*nCS | MSD T | MQD T | hist T*nCS | Cvv T | MSR1_R 3T]  (allreduced between pack and accumulate)
	double *av[SQT_N2D];      // [nCS][T] each, order above
	double *av1[SQT_N1D];     // [T] each
	// what AccumSpacetimeCorr adds depends on the getopt flags (lmp2sqt.c:381-432): F under -f, F_s / F_d under -F,
	// DrTable under -f and -t (EvalTwoTimeEach sits inside `if (flag_t)`, :1370), MSD / MQD / MSDCM under -t
	int flag_f, flag_F, flag_t, flag_V, flag_M, flag_s;
	const double *bufC2, *bufCvv, *bufCvcm; // -V / -M: [SQT_NC2][B][T][nCS], [B][T], [B][T]
	const double *bufMSR1, *bufMSR2;        // -s: [B][T][3] (this shard's particles), [B][T][9]
	int B;
	// per-type partials: the reduce kernels wrote per-species buffers [n_types][B][T][nCS] / [n_types][B][T] instead of the plain
	// ones; pack sums them into the plain block and appends one block per species [Fs TS | MSD T | MQD T | hist TS] to cblk
	int n_types;
	const double *bufFs_t, *bufMSD_t, *bufMQD_t, *bufFab;
	const int *bufHist_t;
	double *tav;              // type sums behind the SQT_N1D arrays (layout above)
};
int sqt_launch_pack(const AccumArgs &a, cudaStream_t st);
int sqt_launch_accumulate(const AccumArgs &a, cudaStream_t st);

int sqt_launch_soa_copy(double *dst, long long dst_pitch, const double *src, long long src_pitch, int n, int ncomp, cudaStream_t st);

// ---- sqt_vel.cu: velocity autocorrelation over the origin ring (-V: rrCvv, lmp2sqt.c:1397-1401, 1441) ----------------
struct VelArgs {
	const double *vcur;       // current velocities, SoA [3][cur_pitch]
	double *vring;            // origin velocities [slot][3][RP]
	double *partial;          // [B][nchunk]
	double *bufCvv;           // [B][T]
	long long cur_pitch, RP;
	int n_local, nchunk, chunk;
	int p, T, B;
};
int sqt_launch_vel(const VelArgs &a, cudaStream_t st);   // vel_ring_kernel + vel_reduce_kernel

// -s: per-axis mean-square displacement over the origin ring (rrMSR1_R, lmp2sqt.c:1434-1436) and the frame's sum_n v (x) r
struct AxisMsdArgs {
	const double *cur;        // current positions, SoA [3][cur_pitch]
	const double *ring;       // origin positions: element (slot, axis, particle) at slot*s_slot + axis*s_axis + particle*s_ptl
	long long s_slot, s_axis, s_ptl;
	double *partial;          // [B][nchunk][3]
	double *bufMSR1;          // [B][T][3]
	long long cur_pitch;
	int n_local, nchunk, chunk;
	int p, T, B;
};
int sqt_launch_axis_msd(const AxisMsdArgs &a, cudaStream_t st); // axis_msd_kernel + axis_msd_reduce_kernel
struct SumVRArgs {
	const double *cur;        // SoA: x y z at components 0..2, vx vy vz at components 3..5
	double *partial;          // [nblocks][9]
	double *out;              // [9]: sum_n v_a r_b, index 3 a + b (Rank2R3 order xx xy xz yx ..)
	long long cur_pitch;
	int n_local, nblocks;
};
int sqt_launch_sum_vr(const SumVRArgs &a, cudaStream_t st);     // sum_vr_kernel + sum_vr_reduce_kernel

// ---- sqt_api.cu internals used by sqt_multi.cu -------------------------------------------------
struct sqt_handle;
cudaStream_t sqt_internal_copy_stream(sqt_handle *h); // the stream the handle's H2D copies are issued on

// ---- sqt_comm.cpp -------------------------------------------------------------------------
struct SqtComm;
int sqt_comm_get_unique_id(void *out128, char *err, int errlen);
int sqt_comm_create(SqtComm **out, const void *id128, int rank, int nranks, char *err, int errlen);
int sqt_comm_allreduce_f64(SqtComm *c, double *buf, size_t count, cudaStream_t st, char *err, int errlen);
void sqt_comm_destroy(SqtComm *c);
