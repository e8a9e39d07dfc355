// sqt_api.cu -- the C-ABI of include/sqt.h: ring control (the reference's InitSpacetimeCorr /
// EvalSpacetimeCorr / AccumSpacetimeCorr state machine, lmp2sqt.c:446-475, 1599-1650, 368-444,
// kept on the host as pure integer bookkeeping) and the stream choreography around the kernels
// of sqt_kernels.cu.  All floating-point work of the hot path happens on the device.
#include "../../include/sqt.h"
#include "sqt_internal.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static char g_create_err[512] = "";

struct sqt_handle {
	sqt_params p;
	int S, T, B;
	int device, n_sm;
	char err[512];
	char cfg[1024];
	bool flag_f, flag_F, flag_t, flag_V, flag_M, flag_s;
	int ncomp;                // arrays per frame: x y z, then vx vy vz (-V or -s), then mux muy muz (-M)
	int comp_v, comp_m;       // first component of the velocities / dipoles (-1: not carried)

	// ---- ring bookkeeping (host mirror of TBuf.count etc.) ----
	int limitCorrAv, countCorrAv, n_print;
	int nSkip, p_file;
	std::vector<int> count;
	std::vector<double> h_invL0;
	long long nPtls;
	int n_local;
	bool sized, first_done, broken;
	double L, vol, kVal, first_vol;
	int first_limit;

	// ---- device state ----
	double *d_cur[2];
	long long cur_pitch;
	double *h_stage[2];
	int stage_cap;
	cudaEvent_t ev_copy[2], ev_ring[2];
	bool copy_pending[2];
	int stage_next;
	double *d_ring;
	long long ring_np;        // particle pitch of the [slot][axis][particle] ring (tensor-core path)
	int RP;
	double *d_partial;
	int *d_tile_ctr;
	int *d_hist_cur, *d_bufHist;
	double *d_bufFs, *d_bufF, *d_bufMSD, *d_bufMQD, *d_bufMSDCM;
	double *d_rho0, *d_sumr0, *d_invL0, *d_rho_cur, *d_cblk;
	int rho_n;                // 6 nCS + 3: one field block of d_rho_cur (rho + sum r; then kvel + sum v with -V, kmu + sum mu with -M)
	int n_fields, blk_v, blk_m; // field blocks in use and where kvel / kmu sit; then 9 doubles sum_n v (x) r (-s), then a scratch block
	int off_vr, off_scratch, coll_n; // offsets into d_rho_cur; doubles the per-frame allreduce covers
	long long *d_thr;
	double *d_avblk, *d_latblk; // [10*S*T + 5*T]
	size_t avblk_n;
	bool latched;
	int latched_count, latched_limit;
	int self_kind;            // which kernel serves the per-particle terms: SELF_NONE / SELF_MMA / SELF_VEC / SELF_CLASSIC
	bool fused;               // small-shape path: rho(k) inside the vector kernel, one kernel finishes the frame (no communicator)
	KernelChoice kc;
	MmaChoice mc;             // blocked-Chebyshev / DMMA path for the real slots (large nCSpatial)
	double *d_mma_partial;
	RhoMmaChoice rmc;
	VecChoice vc;             // particle-per-lane vector kernel (nCSpatial <= 32)
	double *d_rho_partial;
	double *d_mma_extra, *d_rho_extra; // per-warp sums kept outside the DMMA product (MmaChoice.SPOUT)
	int *d_mma_ctr;
	// -V
	double *d_vring, *d_vpartial, *d_bufC2, *d_bufCvv, *d_bufCvcm, *d_kvel0, *d_sumv0;
	int vel_nchunk, vel_chunk;
	// -M, -s
	double *d_kmu0, *d_sumVR_state, *d_orgSumVR, *d_bufMSR1, *d_bufMSR2, *d_axis_partial, *d_vr_partial;
	long long vr_frames;      // processed frames so far (parity of the double-buffered sum_n v (x) r state)
	// per-type partials: the frame is regrouped by species on the device (each species a contiguous
	// range starting at a multiple of 32 particles), every per-particle pass runs once per range
	int n_types;              // 0 / 1: off
	std::vector<int> h_types; // species (1 .. n_types) of this shard's atoms, dump order (sqt_set_types)
	int type_start[SQT_MAX_TYPES], type_count[SQT_MAX_TYPES];
	long long type_total[SQT_MAX_TYPES]; // atoms of each species over all shards
	int n_sorted;             // length of the grouped frame (sum of the ranges, each padded to 32)
	int *d_perm;              // [n_sorted]: dump-order index of the particle at a grouped position, -1 in the padding
	double *d_raw[2];         // host frames land here in dump order before they are regrouped into d_cur
	double *d_bufFs_t, *d_bufMSD_t, *d_bufMQD_t, *d_rho0_t, *d_bufFab;
	int *d_bufHist_t;
	size_t tav_off;           // offset of the species' running sums inside d_avblk
	cudaStream_t s_main, s_copy, s_coll, s_aux;
	cudaEvent_t ev_red, ev_coll, ev_rho, ev_rhoc, ev_prod, ev_t0, ev_t1;
	bool coll_pending, rhoc_pending;
	bool rho_reader;          // the caller reads per-frame results (sqt_get_rho / _kvel / _kmu): see run_frame, stage A
	SqtComm *comm;
	int rank, nranks;
	long long launches;
	bool prof;
	std::vector<cudaEvent_t> prof_pool; // pairs, pre-created by sqt_profile_enable so that the timed loop creates nothing
	size_t prof_used;
};
enum { SELF_NONE = 0, SELF_MMA = 1, SELF_VEC = 2, SELF_CLASSIC = 3 };

static int fail(sqt_handle *h, int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(h ? h->err : g_create_err, 512, fmt, ap);
	va_end(ap);
	return code;
}

#define CU(call)                                                                                          \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) return fail(h, SQT_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

extern "C" const char *sqt_last_error(const sqt_handle *h) { return h ? h->err : g_create_err; }

extern "C" int sqt_results_alloc(sqt_results *r, int S, int T)
{
	if (!r || S < 1 || T < 1) return SQT_ERR_ARG;
	memset(r, 0, sizeof(*r));
	r->nCSpatial = S; r->nCTime = T;
	double **a2[SQT_N2D] = { &r->avF_qq2, &r->StdDevF_qq2, &r->avF_s_qq2, &r->StdDevF_s_qq2, &r->avF_d_qq2, &r->StdDevF_d_qq2, &r->avDrTable,
	                         &r->avC2_v_rho, &r->avC2_rho_v, &r->avC2_v_v, &r->avC2_v_mu, &r->avC2_mu_v, &r->avC2_mu_rho, &r->avC2_rho_mu, &r->avC2_mu_mu };
	double **a1[6] = { &r->rrMSDAv, &r->rrMQDAv, &r->rrMSDCMAv, &r->rrCvvAv, &r->rrCvcmvcmAv, &r->rrCmmAv };
	for (int i = 0; i < SQT_N2D; ++i) if (!(*a2[i] = (double *)calloc((size_t)S * T, sizeof(double)))) return SQT_ERR_NOMEM;
	for (int i = 0; i < 6; ++i) if (!(*a1[i] = (double *)calloc(T, sizeof(double)))) return SQT_ERR_NOMEM;
	if (!(r->rrMSR1_R_Av = (double *)calloc((size_t)T * 3, sizeof(double)))) return SQT_ERR_NOMEM;
	if (!(r->rrMSR2_VR_Av = (double *)calloc((size_t)T * 9, sizeof(double)))) return SQT_ERR_NOMEM;
	return SQT_OK;
}

extern "C" int sqt_results_alloc_types(sqt_results *r, int n_types)
{
	if (!r || n_types < 2 || n_types > SQT_MAX_TYPES || r->nCSpatial < 1) return SQT_ERR_ARG;
	const size_t ST = (size_t)r->nCSpatial * r->nCTime, T = r->nCTime, nt = n_types;
	r->avF_ab = (double *)calloc(nt * nt * ST, sizeof(double));
	r->avF_s_a = (double *)calloc(nt * ST, sizeof(double));
	r->avDrTable_a = (double *)calloc(nt * ST, sizeof(double));
	r->rrMSD_a = (double *)calloc(nt * T, sizeof(double));
	r->rrMQD_a = (double *)calloc(nt * T, sizeof(double));
	if (!r->avF_ab || !r->avF_s_a || !r->avDrTable_a || !r->rrMSD_a || !r->rrMQD_a) {
		free(r->avF_ab); free(r->avF_s_a); free(r->avDrTable_a); free(r->rrMSD_a); free(r->rrMQD_a);
		r->avF_ab = r->avF_s_a = r->avDrTable_a = r->rrMSD_a = r->rrMQD_a = nullptr;
		return SQT_ERR_NOMEM;
	}
	r->n_types = n_types;
	return SQT_OK;
}

extern "C" void sqt_results_free(sqt_results *r)
{
	if (!r) return;
	free(r->avF_ab); free(r->avF_s_a); free(r->avDrTable_a); free(r->rrMSD_a); free(r->rrMQD_a);
	free(r->avF_qq2); free(r->StdDevF_qq2); free(r->avF_s_qq2); free(r->StdDevF_s_qq2);
	free(r->avF_d_qq2); free(r->StdDevF_d_qq2); free(r->avDrTable);
	free(r->rrMSDAv); free(r->rrMQDAv); free(r->rrMSDCMAv);
	free(r->avC2_v_rho); free(r->avC2_rho_v); free(r->avC2_v_v); free(r->rrCvvAv); free(r->rrCvcmvcmAv);
	free(r->avC2_v_mu); free(r->avC2_mu_v); free(r->avC2_mu_rho); free(r->avC2_rho_mu); free(r->avC2_mu_mu); free(r->rrCmmAv);
	free(r->rrMSR1_R_Av); free(r->rrMSR2_VR_Av);
	memset(r, 0, sizeof(*r));
}

// Exact van Hove bin thresholds: thr[k] = bit pattern of the smallest double x >= 0 with
// floor(sqrt(x)/rVal) >= k, found by bisection on the (monotone) bit pattern with the host's
// correctly rounded IEEE sqrt and divide -- the very operations of lmp2sqt.c:1429.
static long long bin_threshold(int k, double rVal)
{
	if (k <= 0) return 0;
	long long lo = 0, hi = 0x7fe0000000000000LL; // f(lo) false (k >= 1), f(hi) true for sane rVal
	while (hi - lo > 1) {
		long long mid = lo + (hi - lo) / 2;
		double x;
		memcpy(&x, &mid, sizeof(x));
		volatile double sq = sqrt(x);
		volatile double q = sq / rVal;
		if (floor(q) >= (double)k) hi = mid; else lo = mid;
	}
	return hi;
}

extern "C" int sqt_create(const sqt_params *p, sqt_handle **out)
{
	sqt_handle *h = nullptr;
	if (!p || !out) return fail(h, SQT_ERR_ARG, "sqt_create: null argument");
	if (p->nCSpatial < 1 || p->nCTime < 1) return fail(h, SQT_ERR_ARG, "sqt_create: nCSpatial and nCTime must be >= 1");
	if (p->nCSpatial > SQT_MAX_NCSPATIAL)
		return fail(h, SQT_ERR_ARG, "sqt_create: nCSpatial %d is beyond what the kernels cover (%d)", p->nCSpatial, SQT_MAX_NCSPATIAL);
	if ((p->flag_V || p->flag_M) && p->nCSpatial > SQT_MAX_NCSPATIAL_V)
		return fail(h, SQT_ERR_ARG, "sqt_create: -V / -M need nCSpatial <= %d (got %d)", SQT_MAX_NCSPATIAL_V, p->nCSpatial);
	if (!(p->rVal > 0)) return fail(h, SQT_ERR_ARG, "sqt_create: rVal must be > 0");
	if (p->n_types < 0 || p->n_types > SQT_MAX_TYPES) return fail(h, SQT_ERR_ARG, "sqt_create: n_types must be 0 .. %d", SQT_MAX_TYPES);
	if (p->n_types > 1 && (p->flag_V || p->flag_M || p->flag_s))
		return fail(h, SQT_ERR_ARG, "sqt_create: per-type partials (n_types > 1) do not combine with -V / -M / -s");
	if (p->n_types > 1 && p->nCSpatial > SQT_MAX_NCSPATIAL_V)
		return fail(h, SQT_ERR_ARG, "sqt_create: per-type partials need nCSpatial <= %d (got %d)", SQT_MAX_NCSPATIAL_V, p->nCSpatial);
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev < 1)
		return fail(h, SQT_ERR_CUDA, "sqt_create: no usable CUDA device (%s); this engine has no CPU fallback",
		            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
	if (p->device < 0 || p->device >= ndev) return fail(h, SQT_ERR_ARG, "sqt_create: device %d out of range (%d devices)", p->device, ndev);
	h = new sqt_handle();
	h->err[0] = 0; h->cfg[0] = 0;
	h->p = *p;
	h->S = p->nCSpatial; h->T = p->nCTime;
	h->B = p->nCBuffer;
	if (h->B > h->T) h->B = h->T; // GetNameList tail, lmp2sqt.c:2105
	if (h->B < 1) h->B = 1;       // main, lmp2sqt.c:240-242
	h->device = p->device;
	h->flag_F = p->flag_F != 0;
	h->flag_f = p->flag_f != 0 || h->flag_F; // -F switches -f on as well (lmp2sqt.c:160-165)
	h->flag_s = p->flag_s != 0;
	h->flag_t = p->flag_t != 0 || h->flag_s; // -s switches -t on as well (lmp2sqt.c:156-160)
	h->flag_V = p->flag_V != 0;
	h->flag_M = p->flag_M != 0;
	h->comp_v = (h->flag_V || h->flag_s) ? 3 : -1;
	h->comp_m = h->flag_M ? (h->comp_v >= 0 ? 6 : 3) : -1;
	h->ncomp = 3 + (h->comp_v >= 0 ? 3 : 0) + (h->comp_m >= 0 ? 3 : 0);
	h->limitCorrAv = p->limitCorrAv;
	h->countCorrAv = 0; h->n_print = 0; h->nSkip = 0; h->p_file = 0;
	h->count.assign(h->B, 0);
	h->h_invL0.assign(h->B, 0.0);
	h->nPtls = 0; h->n_local = 0; h->sized = false; h->first_done = false; h->broken = false;
	h->L = h->vol = h->first_vol = 0; h->kVal = p->kVal; h->first_limit = 0;
	h->d_cur[0] = h->d_cur[1] = nullptr; h->h_stage[0] = h->h_stage[1] = nullptr; h->stage_cap = 0;
	h->copy_pending[0] = h->copy_pending[1] = false; h->stage_next = 0;
	h->d_ring = nullptr; h->d_partial = nullptr; h->d_tile_ctr = nullptr; h->d_mma_partial = nullptr; h->d_mma_ctr = nullptr;
	h->mc.on = 0; h->mc.geo = 0; h->vc.on = 0; h->rmc.on = 0; h->self_kind = SELF_NONE; h->fused = false;
	h->d_rho_partial = nullptr; h->d_mma_extra = h->d_rho_extra = nullptr; h->d_hist_cur = h->d_bufHist = nullptr;
	h->d_bufFs = h->d_bufF = h->d_bufMSD = h->d_bufMQD = h->d_bufMSDCM = nullptr;
	h->d_rho0 = h->d_sumr0 = h->d_invL0 = h->d_rho_cur = h->d_cblk = nullptr; h->d_thr = nullptr;
	h->d_avblk = h->d_latblk = nullptr;
	h->d_vring = h->d_vpartial = h->d_bufC2 = h->d_bufCvv = h->d_bufCvcm = h->d_kvel0 = h->d_sumv0 = nullptr;
	h->vel_nchunk = h->vel_chunk = 0;
	h->d_kmu0 = h->d_sumVR_state = h->d_orgSumVR = h->d_bufMSR1 = h->d_bufMSR2 = h->d_axis_partial = h->d_vr_partial = nullptr;
	h->vr_frames = 0;
	h->n_types = p->n_types > 1 ? p->n_types : 0;
	h->n_sorted = 0; h->d_perm = nullptr; h->d_raw[0] = h->d_raw[1] = nullptr; h->tav_off = 0;
	h->d_bufFs_t = h->d_bufMSD_t = h->d_bufMQD_t = h->d_rho0_t = h->d_bufFab = nullptr; h->d_bufHist_t = nullptr;
	for (int t = 0; t < SQT_MAX_TYPES; ++t) { h->type_start[t] = h->type_count[t] = 0; h->type_total[t] = 0; }
	h->latched = false; h->latched_count = 0; h->latched_limit = 0;
	h->comm = nullptr; h->rank = 0; h->nranks = 1; h->launches = 0; h->prof = false; h->prof_used = 0;
	h->coll_pending = h->rhoc_pending = false; h->rho_reader = false;
	h->s_main = h->s_copy = h->s_coll = h->s_aux = nullptr;
	h->ev_copy[0] = h->ev_copy[1] = h->ev_ring[0] = h->ev_ring[1] = nullptr;
	h->ev_red = h->ev_coll = h->ev_rho = h->ev_rhoc = h->ev_prod = h->ev_t0 = h->ev_t1 = nullptr;

#define CUC(call)                                                                                   \
	do {                                                                                            \
		cudaError_t e_ = (call);                                                                    \
		if (e_ != cudaSuccess) {                                                                    \
			fail(nullptr, SQT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                  \
			sqt_destroy(h);                                                                         \
			return SQT_ERR_CUDA;                                                                    \
		}                                                                                           \
	} while (0)
	CUC(cudaSetDevice(h->device));
	cudaDeviceProp prop;
	CUC(cudaGetDeviceProperties(&prop, h->device));
	h->n_sm = prop.multiProcessorCount;
	CUC(cudaStreamCreate(&h->s_main));
	CUC(cudaStreamCreate(&h->s_copy));
	CUC(cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking)); // result read-backs that must not wait for the self-term kernel
	h->s_coll = h->s_main; // a separate collective stream is created when a communicator is attached
	for (int i = 0; i < 2; ++i) {
		CUC(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&h->ev_ring[i], cudaEventDisableTiming));
	}
	CUC(cudaEventCreateWithFlags(&h->ev_red, cudaEventDisableTiming));
	CUC(cudaEventCreateWithFlags(&h->ev_coll, cudaEventDisableTiming));
	CUC(cudaEventCreateWithFlags(&h->ev_rho, cudaEventDisableTiming));
	CUC(cudaEventCreateWithFlags(&h->ev_rhoc, cudaEventDisableTiming));
	CUC(cudaEventCreateWithFlags(&h->ev_prod, cudaEventDisableTiming));
	CUC(cudaEventCreate(&h->ev_t0));
	CUC(cudaEventCreate(&h->ev_t1));

	const size_t S = h->S, T = h->T, B = h->B;
	CUC(cudaMalloc(&h->d_hist_cur, sizeof(int) * B * S));
	CUC(cudaMemset(h->d_hist_cur, 0, sizeof(int) * B * S));
	CUC(cudaMalloc(&h->d_bufHist, sizeof(int) * B * T * S));
	CUC(cudaMemset(h->d_bufHist, 0, sizeof(int) * B * T * S));
	double **bufs[] = { &h->d_bufFs, &h->d_bufF };
	for (auto pp : bufs) { CUC(cudaMalloc(pp, sizeof(double) * B * T * S)); CUC(cudaMemset(*pp, 0, sizeof(double) * B * T * S)); }
	double **bufs1[] = { &h->d_bufMSD, &h->d_bufMQD, &h->d_bufMSDCM };
	for (auto pp : bufs1) { CUC(cudaMalloc(pp, sizeof(double) * B * T)); CUC(cudaMemset(*pp, 0, sizeof(double) * B * T)); }
	CUC(cudaMalloc(&h->d_rho0, sizeof(double) * B * 6 * S));
	CUC(cudaMemset(h->d_rho0, 0, sizeof(double) * B * 6 * S));
	CUC(cudaMalloc(&h->d_sumr0, sizeof(double) * B * 3));
	CUC(cudaMemset(h->d_sumr0, 0, sizeof(double) * B * 3));
	CUC(cudaMalloc(&h->d_invL0, sizeof(double) * B));
	CUC(cudaMemset(h->d_invL0, 0, sizeof(double) * B));
	h->rho_n = 6 * (int)S + 3;
	h->blk_v = h->flag_V ? 1 : -1;
	h->blk_m = h->flag_M ? (h->flag_V ? 2 : 1) : -1;
	h->n_fields = 1 + (h->flag_V ? 1 : 0) + (h->flag_M ? 1 : 0) + h->n_types; // (per-type partials: rho_a(k) in blocks 1 .. n_types)
	h->off_vr = h->n_fields * h->rho_n;
	h->coll_n = h->off_vr + (h->flag_s ? 9 : 0);
	h->off_scratch = h->off_vr + 16;
	CUC(cudaMalloc(&h->d_rho_cur, sizeof(double) * (h->off_scratch + h->rho_n))); // rho + sum r | kvel + sum v (-V) | kmu + sum mu (-M) | sum v (x) r (-s) | scratch
	CUC(cudaMemset(h->d_rho_cur, 0, sizeof(double) * (h->off_scratch + h->rho_n)));
	const size_t cblk_n = 2 * T * S + 6 * T + (size_t)h->n_types * (2 * T * S + 2 * T);
	CUC(cudaMalloc(&h->d_cblk, sizeof(double) * cblk_n));
	CUC(cudaMemset(h->d_cblk, 0, sizeof(double) * cblk_n));
	if (h->n_types) {
		const size_t nt = h->n_types;
		double **tb[] = { &h->d_bufFs_t, &h->d_bufMSD_t, &h->d_bufMQD_t, &h->d_rho0_t, &h->d_bufFab };
		const size_t tn[] = { nt * B * T * S, nt * B * T, nt * B * T, B * nt * 6 * S, nt * nt * B * T * S };
		for (int i = 0; i < 5; ++i) { CUC(cudaMalloc(tb[i], sizeof(double) * tn[i])); CUC(cudaMemset(*tb[i], 0, sizeof(double) * tn[i])); }
		CUC(cudaMalloc(&h->d_bufHist_t, sizeof(int) * nt * B * T * S));
		CUC(cudaMemset(h->d_bufHist_t, 0, sizeof(int) * nt * B * T * S));
	}
	if (h->flag_M) {
		CUC(cudaMalloc(&h->d_kmu0, sizeof(double) * B * 6 * S));
		CUC(cudaMemset(h->d_kmu0, 0, sizeof(double) * B * 6 * S));
	}
	if (h->flag_s) {
		CUC(cudaMalloc(&h->d_sumVR_state, sizeof(double) * 18));
		CUC(cudaMemset(h->d_sumVR_state, 0, sizeof(double) * 18));
		CUC(cudaMalloc(&h->d_orgSumVR, sizeof(double) * B * 9));
		CUC(cudaMemset(h->d_orgSumVR, 0, sizeof(double) * B * 9));
		CUC(cudaMalloc(&h->d_bufMSR1, sizeof(double) * B * T * 3));
		CUC(cudaMemset(h->d_bufMSR1, 0, sizeof(double) * B * T * 3));
		CUC(cudaMalloc(&h->d_bufMSR2, sizeof(double) * B * T * 9));
		CUC(cudaMemset(h->d_bufMSR2, 0, sizeof(double) * B * T * 9));
	}
	if (h->flag_V || h->flag_M) {
		CUC(cudaMalloc(&h->d_bufC2, sizeof(double) * SQT_NC2 * B * T * S));
		CUC(cudaMemset(h->d_bufC2, 0, sizeof(double) * SQT_NC2 * B * T * S));
	}
	if (h->flag_V) {
		double **vb[] = { &h->d_bufCvv, &h->d_bufCvcm };
		for (auto pp : vb) { CUC(cudaMalloc(pp, sizeof(double) * B * T)); CUC(cudaMemset(*pp, 0, sizeof(double) * B * T)); }
		CUC(cudaMalloc(&h->d_kvel0, sizeof(double) * B * 6 * S));
		CUC(cudaMemset(h->d_kvel0, 0, sizeof(double) * B * 6 * S));
		CUC(cudaMalloc(&h->d_sumv0, sizeof(double) * B * 3));
		CUC(cudaMemset(h->d_sumv0, 0, sizeof(double) * B * 3));
	}
	h->avblk_n = SQT_N2D * S * T + SQT_N1D * T;
	h->tav_off = h->avblk_n;
	if (h->n_types) h->avblk_n += (size_t)h->n_types * h->n_types * S * T + 2 * (size_t)h->n_types * S * T + 2 * (size_t)h->n_types * T;
	CUC(cudaMalloc(&h->d_avblk, sizeof(double) * h->avblk_n));
	CUC(cudaMemset(h->d_avblk, 0, sizeof(double) * h->avblk_n)); // ZeroAvSpacetimeCorr, lmp2sqt.c:1921
	CUC(cudaMalloc(&h->d_latblk, sizeof(double) * h->avblk_n));
	CUC(cudaMemset(h->d_latblk, 0, sizeof(double) * h->avblk_n));
	std::vector<long long> thr(S + 2);
	for (size_t k = 0; k <= S; ++k) thr[k] = bin_threshold((int)k, p->rVal);
	thr[S + 1] = 0x7fffffffffffffffLL;
	CUC(cudaMalloc(&h->d_thr, sizeof(long long) * (S + 2)));
	CUC(cudaMemcpy(h->d_thr, thr.data(), sizeof(long long) * (S + 2), cudaMemcpyHostToDevice));
#undef CUC
	sqt_begin_file(h);
	*out = h;
	return SQT_OK;
}

// everything size_for_atoms allocated
static void release_sized(sqt_handle *h)
{
	for (int i = 0; i < 2; ++i) { cudaFree(h->d_cur[i]); h->d_cur[i] = nullptr; }
	cudaFree(h->d_ring); cudaFree(h->d_partial); cudaFree(h->d_tile_ctr); cudaFree(h->d_mma_partial); cudaFree(h->d_mma_ctr);
	cudaFree(h->d_rho_partial); cudaFree(h->d_mma_extra); cudaFree(h->d_rho_extra); cudaFree(h->d_vring); cudaFree(h->d_vpartial);
	cudaFree(h->d_axis_partial); cudaFree(h->d_vr_partial); cudaFree(h->d_perm); cudaFree(h->d_raw[0]); cudaFree(h->d_raw[1]);
	h->d_axis_partial = h->d_vr_partial = nullptr; h->d_perm = nullptr; h->d_raw[0] = h->d_raw[1] = nullptr;
	h->d_ring = h->d_partial = h->d_mma_partial = h->d_rho_partial = h->d_mma_extra = h->d_rho_extra = h->d_vring = h->d_vpartial = nullptr;
	h->d_tile_ctr = h->d_mma_ctr = nullptr;
	h->sized = false;
}

extern "C" void sqt_destroy(sqt_handle *h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	cudaDeviceSynchronize();
	if (h->comm) sqt_comm_destroy(h->comm);
	release_sized(h);
	for (int i = 0; i < 2; ++i) {
		if (h->h_stage[i]) cudaFreeHost(h->h_stage[i]);
		if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
		if (h->ev_ring[i]) cudaEventDestroy(h->ev_ring[i]);
	}
	cudaFree(h->d_hist_cur); cudaFree(h->d_bufHist);
	cudaFree(h->d_bufFs); cudaFree(h->d_bufF); cudaFree(h->d_bufMSD); cudaFree(h->d_bufMQD); cudaFree(h->d_bufMSDCM);
	cudaFree(h->d_rho0); cudaFree(h->d_sumr0); cudaFree(h->d_invL0); cudaFree(h->d_rho_cur); cudaFree(h->d_cblk);
	cudaFree(h->d_thr); cudaFree(h->d_avblk); cudaFree(h->d_latblk);
	cudaFree(h->d_bufC2); cudaFree(h->d_bufCvv); cudaFree(h->d_bufCvcm); cudaFree(h->d_kvel0); cudaFree(h->d_sumv0);
	cudaFree(h->d_bufFs_t); cudaFree(h->d_bufMSD_t); cudaFree(h->d_bufMQD_t); cudaFree(h->d_rho0_t); cudaFree(h->d_bufFab); cudaFree(h->d_bufHist_t);
	cudaFree(h->d_kmu0); cudaFree(h->d_sumVR_state); cudaFree(h->d_orgSumVR); cudaFree(h->d_bufMSR1); cudaFree(h->d_bufMSR2);
	for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
	cudaEvent_t evs[] = { h->ev_red, h->ev_coll, h->ev_rho, h->ev_rhoc, h->ev_prod, h->ev_t0, h->ev_t1 };
	for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
	if (h->s_coll && h->s_coll != h->s_main) cudaStreamDestroy(h->s_coll);
	if (h->s_aux) cudaStreamDestroy(h->s_aux);
	if (h->s_copy) cudaStreamDestroy(h->s_copy);
	if (h->s_main) cudaStreamDestroy(h->s_main);
	delete h;
}

extern "C" int sqt_set_limit(sqt_handle *h, int limitCorrAv)
{
	if (!h) return SQT_ERR_ARG;
	h->limitCorrAv = limitCorrAv;
	return SQT_OK;
}

// InitSpacetimeCorr, lmp2sqt.c:446-475.
extern "C" int sqt_begin_file(sqt_handle *h)
{
	if (!h) return SQT_ERR_ARG;
	h->nSkip = 0;
	h->p_file = 0;
	for (int nb = 0; nb < h->B; ++nb) h->count[nb] = -nb * h->T / h->B; // C truncation toward zero
	return SQT_OK;
}

static int ensure_stage(sqt_handle *h, int n)
{
	if (n <= h->stage_cap) return SQT_OK;
	CU(cudaSetDevice(h->device));
	for (int i = 0; i < 2; ++i) {
		if (h->copy_pending[i]) { CU(cudaEventSynchronize(h->ev_copy[i])); h->copy_pending[i] = false; }
		if (h->h_stage[i]) { CU(cudaFreeHost(h->h_stage[i])); h->h_stage[i] = nullptr; }
		CU(cudaMallocHost(&h->h_stage[i], sizeof(double) * h->ncomp * (size_t)n));
	}
	h->stage_cap = n;
	return SQT_OK;
}

extern "C" double *sqt_stage_acquire(sqt_handle *h, int n_atoms_local)
{
	if (!h || n_atoms_local < 1) return nullptr;
	if (ensure_stage(h, n_atoms_local) != SQT_OK) return nullptr;
	const int s = h->stage_next;
	h->stage_next ^= 1;
	if (h->copy_pending[s]) { // the H2D copy that last read this pinned buffer must be done
		if (cudaEventSynchronize(h->ev_copy[s]) != cudaSuccess) { fail(h, SQT_ERR_CUDA, "stage_acquire: event sync failed"); return nullptr; }
		h->copy_pending[s] = false;
	}
	return h->h_stage[s];
}

// First processed frame: the analogue of Alloc_more (lmp2sqt.c:1926-2012) -- everything whose
// size depends on the atom count.  On failure nothing stays allocated and the handle refuses further frames.
static int size_for_atoms_impl(sqt_handle *h, int n_local, long long n_total)
{
	CU(cudaSetDevice(h->device));
	h->n_local = n_local;
	h->nPtls = n_total;
	h->cur_pitch = ((long long)n_local + 31) & ~31LL;
	if (h->n_types) { // group the shard by species: ranges padded to 32 particles, stable inside a species
		if ((int)h->h_types.size() != n_local)
			return fail(h, SQT_ERR_STATE, "per-type partials: sqt_set_types must give the species of the shard's %d atoms before the first frame (got %zu)",
			            n_local, h->h_types.size());
		for (int t = 0; t < h->n_types; ++t) h->type_count[t] = 0;
		for (int i = 0; i < n_local; ++i) h->type_count[h->h_types[i] - 1]++;
		int pos = 0;
		for (int t = 0; t < h->n_types; ++t) { h->type_start[t] = pos; pos += (h->type_count[t] + 31) & ~31; }
		h->n_sorted = pos > 0 ? pos : 32;
		std::vector<int> perm(h->n_sorted, -1), fill(h->n_types, 0);
		for (int i = 0; i < n_local; ++i) { const int t = h->h_types[i] - 1; perm[h->type_start[t] + fill[t]++] = i; }
		CU(cudaMalloc(&h->d_perm, sizeof(int) * h->n_sorted));
		CU(cudaMemcpy(h->d_perm, perm.data(), sizeof(int) * h->n_sorted, cudaMemcpyHostToDevice));
		for (int i = 0; i < 2; ++i) CU(cudaMalloc(&h->d_raw[i], sizeof(double) * h->ncomp * (size_t)n_local));
		h->cur_pitch = h->n_sorted;
		// atoms of each species over all shards (normalisation of F_s,a and MSD_a)
		double cnt[SQT_MAX_TYPES];
		for (int t = 0; t < h->n_types; ++t) cnt[t] = (double)h->type_count[t];
		if (h->comm) {
			CU(cudaMemcpy(h->d_cblk, cnt, sizeof(double) * h->n_types, cudaMemcpyHostToDevice));
			if (sqt_comm_allreduce_f64(h->comm, h->d_cblk, (size_t)h->n_types, h->s_coll, h->err, sizeof(h->err))) return SQT_ERR_NCCL;
			CU(cudaStreamSynchronize(h->s_coll));
			CU(cudaMemcpy(cnt, h->d_cblk, sizeof(double) * h->n_types, cudaMemcpyDeviceToHost));
			CU(cudaMemset(h->d_cblk, 0, sizeof(double) * h->n_types));
		}
		for (int t = 0; t < h->n_types; ++t) h->type_total[t] = (long long)(cnt[t] + 0.5);
	}
	for (int i = 0; i < 2; ++i) CU(cudaMalloc(&h->d_cur[i], sizeof(double) * h->ncomp * h->cur_pitch));
	// Which kernel serves the per-particle terms.  F_s needs the nCSpatial Chebyshev chains (-F); MSD, MQD and the van Hove
	// table (-t) only the displacement: without -F the particle-per-lane kernel runs with a single chain, and with neither
	// flag nothing per-particle is computed at all (the reference does the same, lmp2sqt.c:1370, 1479, 1527).
	const bool want_self = h->flag_F || h->flag_t;
	int rc = sqt_pick_mma(h->S, h->B, n_local, h->n_sm, &h->mc);
	if (rc != 0) return fail(h, SQT_ERR_CUDA, "mma kernel selection failed (rc=%d)", rc);
	if (!h->flag_F) h->mc.on = 0;
	h->vc.on = 0;
	if (want_self && !h->mc.on && h->mc.geo) { // (rho(k) comes from the tensor-core rho kernel on this path)
		// without a communicator (and without -V) the whole frame is two launches: rho(k) rides in the vector kernel
		const int fuse = !h->comm && !h->flag_V && !h->flag_M && !h->flag_s && !h->n_types && h->flag_F;
		rc = sqt_pick_vec(h->S, h->flag_F ? h->S : 1, h->B, n_local, h->n_sm, fuse, &h->vc);
		if (rc != 0) return fail(h, SQT_ERR_CUDA, "vector kernel selection failed (rc=%d)", rc);
	}
	h->fused = h->vc.on && h->vc.fused;
	h->self_kind = !want_self ? SELF_NONE : (h->mc.on ? SELF_MMA : (h->vc.on ? SELF_VEC : SELF_CLASSIC));
	const bool lane_family = h->self_kind == SELF_MMA || h->self_kind == SELF_VEC; // a particle per lane: particle-minor ring
	// the classic kernel serves its own slots only for SELF_CLASSIC; otherwise at most the six virtual rho(k) slots (B = 0 for it)
	rc = sqt_pick_kernel(h->S, h->self_kind == SELF_CLASSIC ? h->B : 0, n_local, h->n_sm, &h->kc);
	if (rc != 0) return fail(h, SQT_ERR_CUDA, "no ring-kernel variant for nCSpatial=%d (rc=%d)", h->S, rc);
	h->RP = (h->B + 15) & ~15; // 128-byte rows: a warp's 32 origins of one particle/axis are contiguous
	// the tensor-core kernel walks particles along the warp, so its ring is [slot][axis][particle] (pitch: a multiple of 32
	// particles); the classic kernel walks slots along the warp: [particle][axis][slot]
	h->ring_np = h->n_types ? h->n_sorted : (((long long)n_local + 31) & ~31LL);
	size_t ring_bytes = 0;
	if (h->self_kind != SELF_NONE) {
		ring_bytes = lane_family ? sizeof(double) * 3 * (size_t)h->B * (size_t)h->ring_np : sizeof(double) * 3 * (size_t)h->RP * (size_t)h->ring_np;
		cudaError_t e = cudaMalloc(&h->d_ring, ring_bytes);
		if (e != cudaSuccess) return fail(h, SQT_ERR_NOMEM, "origin ring of %.2f GB does not fit on the device: %s", ring_bytes / 1e9, cudaGetErrorString(e));
		CU(cudaMemsetAsync(h->d_ring, 0, ring_bytes, h->s_main));
	}
	const size_t part_bytes = sizeof(double) * (size_t)h->kc.NWSG * h->kc.KW * h->kc.NV * (32 / h->kc.G);
	CU(cudaMalloc(&h->d_partial, part_bytes));
	CU(cudaMemsetAsync(h->d_partial, 0, part_bytes, h->s_main));
	CU(cudaMalloc(&h->d_tile_ctr, sizeof(int) * h->kc.NWSG));
	CU(cudaMemsetAsync(h->d_tile_ctr, 0, sizeof(int) * h->kc.NWSG, h->s_main));
	size_t mma_bytes = 0;
	if (lane_family) {
		const int nwsg = h->mc.on ? h->mc.NWSG : h->vc.NWSG;
		mma_bytes = h->mc.on ? sizeof(double) * (size_t)h->mc.NWSG * h->mc.KW * h->mc.SPW * h->mc.NT * 64
		                     : sizeof(double) * (size_t)h->vc.NWSG * h->vc.KW * h->vc.SPW * (h->vc.C + 2);
		CU(cudaMalloc(&h->d_mma_partial, mma_bytes));
		CU(cudaMemsetAsync(h->d_mma_partial, 0, mma_bytes, h->s_main));
		CU(cudaMalloc(&h->d_mma_ctr, sizeof(int) * 2 * nwsg)); // tile counters, then the fused path's rows-posted counters
		CU(cudaMemsetAsync(h->d_mma_ctr, 0, sizeof(int) * 2 * nwsg, h->s_main));
		const size_t xb = h->mc.on ? sizeof(double) * (size_t)h->mc.NWSG * h->mc.KW * h->mc.SPW * 2 : 16;
		CU(cudaMalloc(&h->d_mma_extra, xb));
		CU(cudaMemsetAsync(h->d_mma_extra, 0, xb, h->s_main));
	}
	h->rmc.on = 0;
	if ((h->self_kind != SELF_CLASSIC || h->flag_V || h->flag_M || h->n_types) && !h->fused) {
		const char *rm = getenv("SQT_RHO_MMA");
		if (!(rm && *rm == '0')) {
			rc = sqt_pick_rho_mma(h->mc, n_local, h->n_sm, &h->rmc);
			if (rc != 0) return fail(h, SQT_ERR_CUDA, "rho_mma_kernel: %s", cudaGetErrorString((cudaError_t)rc));
		}
		if (h->rmc.on) {
			const size_t rb = sizeof(double) * (size_t)h->rmc.total_warps * 6 * h->mc.NT * 64;
			CU(cudaMalloc(&h->d_rho_partial, rb));
			CU(cudaMalloc(&h->d_rho_extra, sizeof(double) * (size_t)h->rmc.total_warps * 3));
			mma_bytes += rb;
		}
	}
	if ((h->flag_V || h->flag_M || h->n_types) && !h->rmc.on)
		return fail(h, SQT_ERR_ARG, "-V / -M / per-type partials need the tensor-core rho(k) kernel (nCSpatial <= %d, SQT_MMA / SQT_RHO_MMA not 0)", SQT_MAX_NCSPATIAL_V);
	if (h->flag_s) {
		int chunk = 8192, nch = (n_local + chunk - 1) / chunk;
		if (nch > 2048) { nch = 2048; chunk = (n_local + nch - 1) / nch; nch = (n_local + chunk - 1) / chunk; }
		h->vel_chunk = chunk; h->vel_nchunk = nch;
		CU(cudaMalloc(&h->d_axis_partial, sizeof(double) * 3 * (size_t)h->B * nch));
		CU(cudaMalloc(&h->d_vr_partial, sizeof(double) * 9 * 1024));
	}
	if (h->flag_V) {
		if (h->flag_t) {
			const size_t vb = sizeof(double) * 3 * (size_t)h->B * (size_t)h->ring_np;
			cudaError_t e = cudaMalloc(&h->d_vring, vb);
			if (e != cudaSuccess) return fail(h, SQT_ERR_NOMEM, "origin-velocity ring of %.2f GB does not fit on the device: %s", vb / 1e9, cudaGetErrorString(e));
			CU(cudaMemsetAsync(h->d_vring, 0, vb, h->s_main));
			ring_bytes += vb;
			int chunk = 8192, nch = (n_local + chunk - 1) / chunk;
			if (nch > 2048) { nch = 2048; chunk = (n_local + nch - 1) / nch; nch = (n_local + chunk - 1) / chunk; }
			h->vel_chunk = chunk; h->vel_nchunk = nch;
			CU(cudaMalloc(&h->d_vpartial, sizeof(double) * (size_t)h->B * nch));
		}
	}
	int off = 0;
	char classic[400];
	snprintf(classic, sizeof(classic),
	         "self_ring_kernel<C=%d,G=%d> persistent grid=%d block=%d regs=%d smem=%zuB occ=%d blocks/SM; %d slot groups x %d warps, "
	         "%s tiles of %d particles (%d tiles)",
	         h->kc.C, h->kc.G, h->kc.blocks, h->kc.WPB * 32, h->kc.regs, h->kc.smem, h->kc.occ, h->kc.NWSG, h->kc.KW,
	         h->kc.dynamic ? "dynamic" : "static", h->kc.tile, h->kc.ntiles);
	if (h->self_kind == SELF_NONE)
		off = snprintf(h->cfg, sizeof(h->cfg), "no per-particle kernel (neither -F nor -t); rho(k): ");
	if (h->self_kind == SELF_VEC)
		off = snprintf(h->cfg, sizeof(h->cfg),
		               "self_ring_vec_kernel<C=%d,SPW=%d> (a particle per lane, sums in registers%s) persistent grid=%d block=%d regs=%d smem=%zuB occ=%d blocks/SM; "
		               "%d slot groups x %d warps, %s tiles of %d particles; + rho(k): %s",
		               h->vc.C, h->vc.SPW, h->flag_F ? "" : "; no -F: MSD / van Hove only", h->vc.blocks, h->vc.WPB * 32, h->vc.regs, h->vc.smem, h->vc.occ, h->vc.NWSG, h->vc.KW,
		               h->vc.dynamic ? "dynamic" : "static", h->vc.tile,
		               h->fused ? "six virtual slots of the same kernel, frame finished by vec_finish_kernel (2 launches per frame); unused: " : "");
	if (h->self_kind == SELF_MMA)
		off = snprintf(h->cfg, sizeof(h->cfg),
		               "self_ring_mma_kernel<NT=%d,SPW=%d> (blocked Chebyshev S=%d on DMMA%s) persistent grid=%d block=%d regs=%d smem=%zuB occ=%d blocks/SM; "
		               "%d slot groups x %d warps, %s tiles of %d particles; + rho(k): ",
		               h->mc.NT, h->mc.SPW, h->mc.S, h->mc.SPOUT ? ", MSD/MQD outside" : "", h->mc.blocks, h->mc.WPB * 32, h->mc.regs, h->mc.smem, h->mc.occ, h->mc.NWSG, h->mc.KW,
		               h->mc.dynamic ? "dynamic" : "static", h->mc.tile);
	if (h->self_kind == SELF_CLASSIC) // the recurrence kernel serves the self terms and, through its six virtual slots, rho(k)
		off = snprintf(h->cfg, sizeof(h->cfg), "%s%s", classic, h->rmc.on ? "; + rho(k), kvel: " : "");
	if (h->rmc.on)
		off += snprintf(h->cfg + off, sizeof(h->cfg) - off, "rho_mma_kernel<NT=%d> grid=%d block=%d smem=%zuB occ=%d blocks/SM, static tiles of 32 particles%s",
		                h->mc.NT, h->rmc.blocks, h->rmc.WPB * 32, h->rmc.smem, h->rmc.occ,
		                h->flag_V && h->flag_M ? " (+ velocity- and dipole-weighted passes, vel_ring_kernel)"
		                : h->flag_V ? " (+ velocity-weighted pass, vel_ring_kernel)" : h->flag_M ? " (+ dipole-weighted pass)" : "");
	if (h->self_kind != SELF_CLASSIC)
		off += snprintf(h->cfg + off, sizeof(h->cfg) - off, "%s%s", h->rmc.on ? "; unused: " : "", classic);
	off += snprintf(h->cfg + off, sizeof(h->cfg) - off, "; slots=%d(+%d virtual) ring=%.3fGB partials=%.1fMB", h->B, SQT_NVIRT, ring_bytes / 1e9,
	                (part_bytes + mma_bytes) / 1e6);
	if (h->n_types && off < (int)sizeof(h->cfg))
		snprintf(h->cfg + off, sizeof(h->cfg) - off, "; per-type partials: %d species, frame regrouped on the device, one pass per species", h->n_types);
	h->sized = true;
	return SQT_OK;
}

static int size_for_atoms(sqt_handle *h, int n_local, long long n_total)
{
	const int rc = size_for_atoms_impl(h, n_local, n_total);
	if (rc != SQT_OK) { // nothing half-allocated survives; the handle stays unusable for frames (the reference exits here, lmp2sqt.c:1651-1657)
		cudaStreamSynchronize(h->s_main);
		release_sized(h);
		h->broken = true;
	}
	return rc;
}

static cudaEvent_t prof_event(sqt_handle *h)
{
	if (h->prof_used == h->prof_pool.size()) { // (only if more frames are profiled than sqt_profile_enable prepared)
		cudaEvent_t e = nullptr;
		cudaEventCreate(&e);
		h->prof_pool.push_back(e);
	}
	return h->prof_pool[h->prof_used++];
}

// rho(k) is final on the device (and, sharded, summed over ranks): the F(q,t) part of the frame (lmp2sqt.c:1584-1586),
// the origin copies of rho and the centre-of-mass terms.  Runs on the collective stream, which only waits for `after`.
static int stage_collective(sqt_handle *h, int p, long long n_total, bool split)
{
	const int S = h->S, T = h->T, B = h->B;
	if (split) {
		CU(cudaEventRecord(h->ev_rho, h->s_main));
		CU(cudaStreamWaitEvent(h->s_coll, h->ev_rho, 0));
	}
	if (h->comm && sqt_comm_allreduce_f64(h->comm, h->d_rho_cur, (size_t)h->coll_n, h->s_coll, h->err, sizeof(h->err)))
		return SQT_ERR_NCCL;
	CollectiveArgs ca;
	ca.rho_cur = h->d_rho_cur; ca.rho0 = h->d_rho0; ca.sumr0 = h->d_sumr0; ca.bufF = h->d_bufF; ca.bufMSDCM = h->d_bufMSDCM;
	ca.p = p; ca.T = T; ca.B = B; ca.nCS = S;
	ca.inv_n2 = 1.0 / ((double)n_total * (double)n_total);
	ca.flag_V = h->flag_V; ca.kvel0 = h->d_kvel0; ca.sumv0 = h->d_sumv0; ca.bufC2 = h->d_bufC2; ca.bufCvcm = h->d_bufCvcm;
	ca.flag_M = h->flag_M; ca.kmu0 = h->d_kmu0; ca.rho_n = h->rho_n; ca.blk_v = h->blk_v; ca.blk_m = h->blk_m;
	ca.flag_s = h->flag_s; ca.off_vr = h->off_vr; ca.vr_cumulative = !h->flag_f; ca.vr_parity = (int)(h->vr_frames & 1);
	ca.sumVR_state = h->d_sumVR_state; ca.orgSumVR = h->d_orgSumVR; ca.bufMSR2 = h->d_bufMSR2;
	h->vr_frames++;
	ca.n_types = h->n_types; ca.rho0_t = h->d_rho0_t; ca.bufFab = h->d_bufFab;
	if (!h->flag_V && !h->flag_M) ca.rho_n = h->rho_n;
	int rc = sqt_launch_collective(ca, h->s_coll);
	if (rc) return fail(h, SQT_ERR_CUDA, "collective_ring_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
	h->launches++;
	CU(cudaEventRecord(h->ev_rhoc, h->s_coll));
	h->rhoc_pending = true;
	return SQT_OK;
}

// AccumSpacetimeCorr (lmp2sqt.c:376-443) on the host mirror of TBuf.count: kernels for every buffer that completed in this
// frame (except `done_on_device`, which the fused finish kernel already added), the counters, the PrintProcess latch
static int accumulate_completed(sqt_handle *h, int done_on_device, int *status)
{
	const int S = h->S, T = h->T, B = h->B;
	const size_t ST = (size_t)S * T;
	int rc;
	for (int nb = 0; nb < B; ++nb) {
		if (h->count[nb] != T) continue;
		if (nb != done_on_device) {
			AccumArgs aa;
			aa.nb = nb; aa.T = T; aa.nCS = S; aa.B = B;
			aa.bufF = h->d_bufF; aa.bufFs = h->d_bufFs; aa.bufMSD = h->d_bufMSD; aa.bufMQD = h->d_bufMQD; aa.bufMSDCM = h->d_bufMSDCM;
			aa.bufHist = h->d_bufHist; aa.cblk = h->d_cblk;
			aa.flag_f = h->flag_f; aa.flag_F = h->flag_F; aa.flag_t = h->flag_t; aa.flag_V = h->flag_V; aa.flag_M = h->flag_M; aa.flag_s = h->flag_s;
			aa.bufC2 = h->d_bufC2; aa.bufCvv = h->d_bufCvv; aa.bufCvcm = h->d_bufCvcm; aa.bufMSR1 = h->d_bufMSR1; aa.bufMSR2 = h->d_bufMSR2;
			aa.n_types = h->n_types; aa.bufFs_t = h->d_bufFs_t; aa.bufMSD_t = h->d_bufMSD_t; aa.bufMQD_t = h->d_bufMQD_t; aa.bufHist_t = h->d_bufHist_t;
			aa.bufFab = h->d_bufFab; aa.tav = h->d_avblk + h->tav_off;
			for (int i = 0; i < SQT_N2D; ++i) aa.av[i] = h->d_avblk + (size_t)i * ST;
			for (int i = 0; i < SQT_N1D; ++i) aa.av1[i] = h->d_avblk + SQT_N2D * ST + (size_t)i * T;
			rc = sqt_launch_pack(aa, h->s_coll);
			if (rc) return fail(h, SQT_ERR_CUDA, "buffer_pack_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
			const size_t cb_n = h->n_types ? 2 * ST + 6 * (size_t)T + (size_t)h->n_types * (2 * ST + 2 * (size_t)T) : 2 * ST + (h->flag_s ? 6 : 3) * (size_t)T;
			if (h->comm && sqt_comm_allreduce_f64(h->comm, h->d_cblk, cb_n, h->s_coll, h->err, sizeof(h->err)))
				return SQT_ERR_NCCL;
			rc = sqt_launch_accumulate(aa, h->s_coll);
			if (rc) return fail(h, SQT_ERR_CUDA, "buffer_accumulate_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
			h->launches += 2;
		}
		h->count[nb] = 0;
		++h->countCorrAv;
		if (h->countCorrAv == h->limitCorrAv) {
			// PrintProcess (lmp2sqt.c:338-367): latch the sums for the host writers, then
			// ZeroAvSpacetimeCorr (:366, :477-540) including countCorrAv = 0.
			CU(cudaMemcpyAsync(h->d_latblk, h->d_avblk, sizeof(double) * h->avblk_n, cudaMemcpyDeviceToDevice, h->s_coll));
			CU(cudaMemsetAsync(h->d_avblk, 0, sizeof(double) * h->avblk_n, h->s_coll));
			h->latched = true;
			h->latched_count = h->countCorrAv;
			h->latched_limit = h->limitCorrAv;
			h->countCorrAv = 0;
			h->n_print++; // Number_call_Print, lmp2sqt.c:674
			*status = SQT_FLUSH;
		}
	}
	return SQT_OK;
}

// -s: this shard's sum_n v (x) r of the frame into d_rho_cur + off_vr (EvalOneTimeSumVR, lmp2sqt.c:1122-1142)
static int launch_sum_vr(sqt_handle *h, const double *cur, long long cur_pitch)
{
	SumVRArgs sa;
	sa.cur = cur; sa.partial = h->d_vr_partial; sa.out = h->d_rho_cur + h->off_vr; sa.cur_pitch = cur_pitch; sa.n_local = h->n_local;
	sa.nblocks = (h->n_local + 255) / 256;
	if (sa.nblocks > 1024) sa.nblocks = 1024;
	const int rc = sqt_launch_sum_vr(sa, h->s_main);
	if (rc) return fail(h, SQT_ERR_CUDA, "sum_vr_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
	h->launches += 2;
	return SQT_OK;
}

// cur: the frame's SoA on the device, component stride cur_pitch (the engine's own copy of a host frame, or the caller's
// device buffer of sqt_push_frame_device read in place)
static int run_frame(sqt_handle *h, int slot, const double *cur, long long cur_pitch, long long n_total, double xlo, double xhi)
{
	const int S = h->S, T = h->T, B = h->B;
	h->L = xhi - xlo;                      // lmp2sqt.c:1620
	h->vol = h->L * h->L * h->L;           // :1621
	h->nPtls = n_total;                    // :1622
	if (!h->first_done) { h->first_vol = h->vol; h->first_limit = h->limitCorrAv; h->first_done = true; } // factorDr inputs, :1991-2004
	h->kVal = 2. * M_PI / h->L;            // :1632
	const double invL = 1.0 / h->L;
	const int p = h->p_file;
	const bool split = h->s_coll != h->s_main;
	int rc;

	int varbox = 0;
	for (int nb = 0; nb < B; ++nb)
		if (h->count[nb] > 0 && h->h_invL0[nb] != invL) { varbox = 1; break; }

	if (h->n_types) { // per-type partials: regroup the frame by species (24 B per atom read and written: HBM bound, negligible)
		rc = sqt_launch_gather(h->d_cur[slot], h->cur_pitch, cur, cur_pitch, h->d_perm, h->n_sorted, h->ncomp, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "gather_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		h->launches++;
		cur = h->d_cur[slot]; cur_pitch = h->cur_pitch;
	}

	if (h->fused) {
		// ---- small shapes, no communicator: two launches per frame -----------------------------------------------------------
		cudaEvent_t pe0 = nullptr, pe1 = nullptr;
		if (h->prof) { pe0 = prof_event(h); pe1 = prof_event(h); CU(cudaEventRecord(pe0, h->s_main)); }
		MmaArgs ma;
		ma.cur = cur; ma.ring = h->d_ring; ma.partial = h->d_mma_partial; ma.extra = h->d_mma_extra; ma.tile_ctr = h->d_mma_ctr; ma.hist_cur = h->d_hist_cur;
		ma.invL0 = h->d_invL0; ma.thr = h->d_thr; ma.cur_pitch = cur_pitch; ma.n_local = h->n_local;
		ma.tile = h->vc.tile; ma.ntiles = h->vc.ntiles; ma.NWSG = h->vc.NWSG; ma.NWSG_real = h->vc.NWSG_real;
		ma.KW = h->vc.KW; ma.total_warps = h->vc.total_warps; ma.dynamic = h->vc.dynamic;
		ma.p = p; ma.T = T; ma.B = B; ma.nCS = S; ma.RP = h->ring_np; ma.S = 0; ma.NT = 0; ma.invL = invL;
		ma.inv_rVal = (float)(1.0 / h->p.rVal); ma.varbox = varbox;
		ma.done_ctr = h->d_mma_ctr + h->vc.NWSG; ma.bufFs = h->d_bufFs; ma.bufMSD = h->d_bufMSD; ma.bufMQD = h->d_bufMQD; ma.rho_cur = h->d_rho_cur;
		rc = sqt_launch_vec(h->vc, ma, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "self_ring_vec_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		if (h->prof) CU(cudaEventRecord(pe1, h->s_main));
		CU(cudaEventRecord(h->ev_ring[slot], h->s_main));
		int nb_done = -1;
		for (int nb = 0; nb < B; ++nb) { // ring state machine, lmp2sqt.c:1638-1647
			if (h->count[nb] == 0) h->h_invL0[nb] = invL;
			if (++h->count[nb] == T && nb_done < 0) nb_done = nb;
		}
		VecFinishArgs fa;
		MmaReduceArgs &mr = fa.r;
		mr.partial = h->d_mma_partial; mr.extra = h->d_mma_extra; mr.spout = 0; mr.n_local = h->n_local;
		mr.KW = h->vc.KW; mr.NWSG = h->vc.NWSG; mr.S = 0; mr.NT = h->vc.C + 2; mr.SPW = h->vc.SPW;
		mr.p = p; mr.T = T; mr.B = B; mr.nCS = S; mr.tile_ctr = h->d_mma_ctr; mr.hist_cur = h->d_hist_cur; mr.bufHist = h->d_bufHist;
		mr.bufFs = h->d_bufFs; mr.bufMSD = h->d_bufMSD; mr.bufMQD = h->d_bufMQD; mr.invL0 = h->d_invL0; mr.invL = invL;
		fa.NWSG_real = h->vc.NWSG_real; fa.done_ctr = h->d_mma_ctr + h->vc.NWSG; fa.rho_cur = h->d_rho_cur; fa.rho0 = h->d_rho0; fa.sumr0 = h->d_sumr0; fa.bufF = h->d_bufF;
		fa.bufMSDCM = h->d_bufMSDCM; fa.inv_n2 = 1.0 / ((double)n_total * (double)n_total); fa.nb_done = nb_done;
		fa.flag_f = h->flag_f; fa.flag_F = h->flag_F; fa.flag_t = h->flag_t;
		const size_t ST = (size_t)S * T;
		for (int i = 0; i < 7; ++i) fa.av[i] = h->d_avblk + (size_t)i * ST;
		for (int i = 0; i < 3; ++i) fa.av1[i] = h->d_avblk + SQT_N2D * ST + (size_t)i * T;
		rc = sqt_launch_vec_finish(fa, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "vec_finish_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		h->launches += 2;
		CU(cudaEventRecord(h->ev_rhoc, h->s_main));
		h->rhoc_pending = true;
		int status = SQT_OK;
		rc = accumulate_completed(h, nb_done, &status);
		if (rc) return rc;
		h->nSkip = 0;   // lmp2sqt.c:1649
		h->p_file++;
		return status;
	}

	// ---- stage A: rho(k) of this frame first, so that its allreduce, the F(q,t) update and a caller waiting in
	// sqt_get_rho are not queued behind the (long) self-term kernel -------------------------------------------------
	if (h->rmc.on) {
		// the reduce rewrites d_rho_cur, which the previous frame's allreduce + collective stage (other stream) still reads
		if (split && h->rhoc_pending) CU(cudaStreamWaitEvent(h->s_main, h->ev_rhoc, 0));
		cudaEvent_t pe0 = nullptr, pe1 = nullptr;
		if (h->prof) { pe0 = prof_event(h); pe1 = prof_event(h); CU(cudaEventRecord(pe0, h->s_main)); }
		RhoMmaArgs qa;
		qa.cur = cur; qa.wt = nullptr; qa.kfac = 0.0; qa.partial = h->d_rho_partial; qa.extra = h->d_rho_extra; qa.spout = h->mc.SPOUT;
		qa.rho_cur = h->d_rho_cur; qa.cur_pitch = cur_pitch;
		qa.n_local = h->n_local; qa.total_warps = h->rmc.total_warps; qa.n_sets = h->rmc.blocks; qa.nCS = S; qa.S = h->mc.S; qa.NT = h->mc.NT; qa.invL = invL;
		if (h->n_types) { // rho_a(k) of every species from its own range, then rho(k) as their sum
			for (int t = 0; t < h->n_types; ++t) {
				double *blk = h->d_rho_cur + (size_t)(1 + t) * h->rho_n;
				if (h->type_count[t] == 0) { CU(cudaMemsetAsync(blk, 0, sizeof(double) * h->rho_n, h->s_main)); continue; }
				qa.cur = cur + h->type_start[t]; qa.n_local = h->type_count[t]; qa.rho_cur = blk;
				rc = sqt_launch_rho_mma(h->mc, h->rmc, qa, h->s_main);
				if (rc) return fail(h, SQT_ERR_CUDA, "rho_mma_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
				h->launches += 2;
			}
			TypeSumArgs ts;
			ts.rho_cur = h->d_rho_cur; ts.rho_n = h->rho_n; ts.n_types = h->n_types;
			rc = sqt_launch_type_sum(ts, h->s_main);
			if (rc) return fail(h, SQT_ERR_CUDA, "type_sum_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
			h->launches++;
		} else {
			rc = sqt_launch_rho_mma(h->mc, h->rmc, qa, h->s_main);
			if (rc) return fail(h, SQT_ERR_CUDA, "rho_mma_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
			h->launches += 2;
		}
		if (h->flag_V) { // kvel_q1 (lmp2sqt.c:1203-1208): the same pass with the velocity component as the particle's weight
			qa.wt = cur + h->comp_v * cur_pitch; qa.kfac = h->kVal; qa.rho_cur = h->d_rho_cur + (size_t)h->blk_v * h->rho_n;
			rc = sqt_launch_rho_mma(h->mc, h->rmc, qa, h->s_main);
			if (rc) return fail(h, SQT_ERR_CUDA, "rho_mma_kernel (velocity-weighted) launch failed: %s", cudaGetErrorString((cudaError_t)rc));
			h->launches += 2;
		}
		if (h->flag_M) { // kmu_q1 (:1209-1214): and with the dipole component
			qa.wt = cur + h->comp_m * cur_pitch; qa.kfac = h->kVal; qa.rho_cur = h->d_rho_cur + (size_t)h->blk_m * h->rho_n;
			rc = sqt_launch_rho_mma(h->mc, h->rmc, qa, h->s_main);
			if (rc) return fail(h, SQT_ERR_CUDA, "rho_mma_kernel (dipole-weighted) launch failed: %s", cudaGetErrorString((cudaError_t)rc));
			h->launches += 2;
		}
		if (h->prof) CU(cudaEventRecord(pe1, h->s_main));
		if (h->self_kind != SELF_CLASSIC) {
			if (h->flag_s) { rc = launch_sum_vr(h, cur, cur_pitch); if (rc) return rc; }
			rc = stage_collective(h, p, n_total, split);
			if (rc) return rc;
			// A caller that waits for every frame's rho(k) must not find its allreduce parked behind the self-term kernel: that
			// kernel is persistent and fills every SM, so an NCCL kernel that becomes ready a moment later gets no SM until it
			// ends (measured at 8 GPUs: sqt_get_rho returned 17 ms late and the next frame's H2D copy lost its overlap, e2e
			// 18.8 vs 17.1 ms per frame).  Once the caller has shown that it reads per-frame results, the self-term kernel is
			// held back until the frame's collective stage is done (an allreduce latency per frame); free-running callers
			// keep the fully asynchronous order.
			if (split && h->rho_reader) CU(cudaStreamWaitEvent(h->s_main, h->ev_rhoc, 0));
		}
	}

	// ---- stage B: the per-particle terms over the origin ring ---------------------------------------------------------
	// One pass over the whole shard, or (per-type partials) one pass per species over its contiguous range of the grouped
	// frame: same kernels, the range's offset added to the frame and ring pointers, the species' own result buffers.
	const bool lane_family = h->self_kind == SELF_MMA || h->self_kind == SELF_VEC;
	const bool classic_launch = h->self_kind == SELF_CLASSIC || !h->rmc.on; // its own slots, or just the six virtual rho(k) slots
	const int Bc = h->self_kind == SELF_CLASSIC ? B : 0; // slots the classic kernel treats as real
	const bool late_collective = classic_launch && (!h->rmc.on || h->self_kind == SELF_CLASSIC); // rho(k) comes out of the recurrence kernel's reduce
	auto launch_self = [&](long long off, int cnt) -> int {
		int lrc;
		if (lane_family) {
			const bool mm = h->self_kind == SELF_MMA;
			const int tile = mm ? h->mc.tile : h->vc.tile;
			MmaArgs ma;
			ma.cur = cur + off; ma.ring = h->d_ring + off; ma.partial = h->d_mma_partial; ma.extra = h->d_mma_extra; ma.tile_ctr = h->d_mma_ctr; ma.hist_cur = h->d_hist_cur;
			ma.invL0 = h->d_invL0; ma.thr = h->d_thr; ma.cur_pitch = cur_pitch; ma.n_local = cnt;
			ma.tile = tile; ma.ntiles = (cnt + tile - 1) / tile; ma.NWSG = mm ? h->mc.NWSG : h->vc.NWSG; ma.NWSG_real = ma.NWSG;
			ma.KW = mm ? h->mc.KW : h->vc.KW; ma.total_warps = mm ? h->mc.total_warps : h->vc.total_warps; ma.dynamic = mm ? h->mc.dynamic : h->vc.dynamic;
			ma.p = p; ma.T = T; ma.B = B; ma.nCS = S; ma.RP = h->ring_np; ma.S = mm ? h->mc.S : 0; ma.NT = mm ? h->mc.NT : 0; ma.invL = invL;
			ma.inv_rVal = (float)(1.0 / h->p.rVal); ma.varbox = varbox;
			ma.done_ctr = nullptr; ma.bufFs = ma.bufMSD = ma.bufMQD = ma.rho_cur = nullptr;
			lrc = mm ? sqt_launch_mma(h->mc, ma, h->s_main) : sqt_launch_vec(h->vc, ma, h->s_main);
			if (lrc) return fail(h, SQT_ERR_CUDA, "%s launch failed: %s", mm ? "self_ring_mma_kernel" : "self_ring_vec_kernel", cudaGetErrorString((cudaError_t)lrc));
			h->launches++;
		}
		if (classic_launch) {
			RingArgs ra;
			ra.cur = cur + off; ra.ring = h->d_ring ? h->d_ring + off * 3 * h->RP : nullptr; ra.partial = h->d_partial; ra.hist_cur = h->d_hist_cur;
			ra.invL0 = h->d_invL0; ra.thr = h->d_thr; ra.cur_pitch = cur_pitch; ra.n_local = cnt;
			ra.tile = h->kc.tile; ra.ntiles = (cnt + h->kc.tile - 1) / h->kc.tile; ra.NWSG = h->kc.NWSG; ra.KW = h->kc.KW; ra.total_warps = h->kc.total_warps;
			ra.dynamic = h->kc.dynamic; ra.tile_ctr = h->d_tile_ctr; ra.p = p; ra.T = T; ra.B = Bc; ra.nCS = S; ra.RP = h->RP; ra.V0 = h->kc.V0;
			ra.SPAD = h->kc.SPAD; ra.NV = h->kc.NV; ra.CT = h->kc.CT; ra.invL = invL; ra.inv_rVal = (float)(1.0 / h->p.rVal);
			ra.varbox = varbox;
			lrc = sqt_launch_ring(h->kc, ra, h->s_main);
			if (lrc) return fail(h, SQT_ERR_CUDA, "self_ring_kernel launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
			h->launches++;
		}
		return SQT_OK;
	};
	auto launch_reduce = [&](int cnt, double *bFs, double *bMSD, double *bMQD, int *bHist) -> int {
		int lrc;
		if (lane_family) {
			const bool mm = h->self_kind == SELF_MMA;
			MmaReduceArgs mr;
			mr.partial = h->d_mma_partial; mr.extra = h->d_mma_extra; mr.spout = mm ? h->mc.SPOUT : 0; mr.n_local = cnt;
			mr.KW = mm ? h->mc.KW : h->vc.KW; mr.NWSG = mm ? h->mc.NWSG : h->vc.NWSG;
			mr.S = mm ? h->mc.S : 0; mr.NT = mm ? h->mc.NT : h->vc.C + 2; mr.SPW = mm ? h->mc.SPW : h->vc.SPW;
			mr.p = p; mr.T = T; mr.B = B; mr.nCS = S; mr.tile_ctr = h->d_mma_ctr; mr.hist_cur = h->d_hist_cur; mr.bufHist = bHist;
			mr.bufFs = bFs; mr.bufMSD = bMSD; mr.bufMQD = bMQD; mr.invL0 = h->d_invL0; mr.invL = invL;
			lrc = mm ? sqt_launch_mma_reduce(mr, h->s_main) : sqt_launch_vec_reduce(mr, h->s_main);
			if (lrc) return fail(h, SQT_ERR_CUDA, "%s launch failed: %s", mm ? "mma_reduce_kernel" : "vec_reduce_kernel", cudaGetErrorString((cudaError_t)lrc));
			h->launches++;
		}
		if (classic_launch) {
			ReduceArgs rr;
			rr.partial = h->d_partial; rr.KW = h->kc.KW; rr.NWSG = h->kc.NWSG; rr.tile_ctr = h->d_tile_ctr; rr.NV = h->kc.NV; rr.SPAD = h->kc.SPAD; rr.CT = h->kc.CT; rr.V0 = h->kc.V0; rr.G = h->kc.G;
			rr.p = p; rr.T = T; rr.B = Bc; rr.nCS = S; rr.hist_cur = h->d_hist_cur; rr.bufHist = bHist; rr.bufFs = bFs;
			rr.bufMSD = bMSD; rr.bufMQD = bMQD; rr.invL0 = h->d_invL0; rr.invL = invL;
			// with the tensor-core rho kernel on (-V / -M / per-type partials on the recurrence path) rho(k) is already final: the virtual slots' copy goes to scratch
			rr.rho_cur = h->rmc.on ? h->d_rho_cur + h->off_scratch : h->d_rho_cur;
			lrc = sqt_launch_reduce(rr, h->s_main);
			if (lrc) return fail(h, SQT_ERR_CUDA, "frame_reduce_kernel launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
			h->launches++;
		}
		return SQT_OK;
	};
	const bool typed = h->n_types > 1;
	const int n_ranges = typed ? h->n_types : 1;
	const size_t BTS = (size_t)B * T * S, BT = (size_t)B * T;
	cudaEvent_t pe0 = nullptr, pe1 = nullptr;
	if (h->prof) { pe0 = prof_event(h); pe1 = prof_event(h); CU(cudaEventRecord(pe0, h->s_main)); }
	for (int ti = 0; ti < n_ranges; ++ti) {
		const long long off = typed ? h->type_start[ti] : 0;
		const int cnt = typed ? h->type_count[ti] : h->n_local;
		if (cnt > 0) { rc = launch_self(off, cnt); if (rc) return rc; }
		if (ti == n_ranges - 1 && h->prof) CU(cudaEventRecord(pe1, h->s_main)); // (the species' reduce kernels in between are counted in)
		if (ti == 0) {
			// what follows rewrites per-buffer arrays the previous frame's accumulate stage (other stream) may still read
			if (split && h->coll_pending) CU(cudaStreamWaitEvent(h->s_main, h->ev_coll, 0));
			if (h->d_vring) { // -V -t: rrCvv over the origin-velocity ring (lmp2sqt.c:1397-1401, 1441)
				VelArgs va;
				va.vcur = cur + h->comp_v * cur_pitch; va.vring = h->d_vring; va.partial = h->d_vpartial; va.bufCvv = h->d_bufCvv;
				va.cur_pitch = cur_pitch; va.RP = h->ring_np; va.n_local = h->n_local; va.nchunk = h->vel_nchunk; va.chunk = h->vel_chunk;
				va.p = p; va.T = T; va.B = B;
				rc = sqt_launch_vel(va, h->s_main);
				if (rc) return fail(h, SQT_ERR_CUDA, "vel_ring_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
				h->launches += 2;
			}
			if (h->flag_s) { // -s: rrMSR1_R, the per-axis squares of the displacements (lmp2sqt.c:1434-1436): one more pass over the origin ring
				AxisMsdArgs xa;
				xa.cur = cur; xa.ring = h->d_ring; xa.partial = h->d_axis_partial; xa.bufMSR1 = h->d_bufMSR1; xa.cur_pitch = cur_pitch;
				xa.s_slot = lane_family ? 3 * h->ring_np : 1; xa.s_axis = lane_family ? h->ring_np : h->RP; xa.s_ptl = lane_family ? 1 : 3 * (long long)h->RP;
				xa.n_local = h->n_local; xa.nchunk = h->vel_nchunk; xa.chunk = h->vel_chunk; xa.p = p; xa.T = T; xa.B = B;
				rc = sqt_launch_axis_msd(xa, h->s_main);
				if (rc) return fail(h, SQT_ERR_CUDA, "axis_msd_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
				h->launches += 2;
				// on the path whose collective stage follows the reduce kernels, sum_n v (x) r is formed here: it reads the frame,
				// so it has to be enqueued before ev_ring releases the frame buffer to the copy of the frame after next
				if (late_collective) { rc = launch_sum_vr(h, cur, cur_pitch); if (rc) return rc; }
			}
		}
		if (ti == n_ranges - 1) CU(cudaEventRecord(h->ev_ring[slot], h->s_main)); // d_cur[slot] may be overwritten by the copy of the frame after next
		if (cnt > 0) {
			rc = typed ? launch_reduce(cnt, h->d_bufFs_t + ti * BTS, h->d_bufMSD_t + ti * BT, h->d_bufMQD_t + ti * BT, h->d_bufHist_t + ti * BTS)
			           : launch_reduce(cnt, h->d_bufFs, h->d_bufMSD, h->d_bufMQD, h->d_bufHist);
			if (rc) return rc;
		}
	}
	if (late_collective) {
		rc = stage_collective(h, p, n_total, split);
		if (rc) return rc;
	}
	if (split) {
		CU(cudaEventRecord(h->ev_red, h->s_main));
		CU(cudaStreamWaitEvent(h->s_coll, h->ev_red, 0));
	}

	// ---- ring state machine, lmp2sqt.c:1638-1647 -------------------------------------------
	for (int nb = 0; nb < B; ++nb) {
		if (h->count[nb] == 0) h->h_invL0[nb] = invL;
		++h->count[nb];
	}
	int status = SQT_OK;
	rc = accumulate_completed(h, -1, &status);
	if (rc) return rc;
	if (split) { CU(cudaEventRecord(h->ev_coll, h->s_coll)); h->coll_pending = true; }
	h->nSkip = 0;   // lmp2sqt.c:1649
	h->p_file++;
	return status;
}

// comp[0..ncomp): host pointers of the frame's arrays (x y z [vx vy vz]), or comp[0] = device SoA when on_device
static int push_common(sqt_handle *h, const double *const *comp, bool on_device, void *producer_stream, int n_local, long long n_total,
                       double xlo, double xhi)
{
	if (!h || !comp || !comp[0] || n_local < 1 || n_total < n_local) return h ? fail(h, SQT_ERR_ARG, "push_frame: bad arguments") : SQT_ERR_ARG;
	if (h->broken) return fail(h, SQT_ERR_STATE, "the handle failed while sizing its arrays for the first frame and cannot take frames");
	if (!on_device)
		for (int a = 0; a < h->ncomp; ++a)
			if (!comp[a]) return fail(h, SQT_ERR_ARG, "push_frame: -V / -s need the velocity arrays and -M the dipole arrays");
	if (!(xhi > xlo)) return fail(h, SQT_ERR_ARG, "push_frame: box xhi must exceed xlo");
	if (h->nSkip < h->p.nCSkip) { h->nSkip++; return SQT_OK; } // lmp2sqt.c:1612-1615
	CU(cudaSetDevice(h->device));
	if (!h->sized) { int rc = size_for_atoms(h, n_local, n_total); if (rc) return rc; }
	if (n_local != h->n_local || n_total != h->nPtls)
		return fail(h, SQT_ERR_STATE, "atom count changed between frames (%d/%lld -> %d/%lld): the reference sizes its arrays once (lmp2sqt.c:1623-1630)",
		            h->n_local, h->nPtls, n_local, n_total);
	int slot;
	if (on_device) {
		// engine-owned copy, so the producer may refill its buffer.  Ordering against the producer: the legacy default stream
		// synchronises with the engine's (blocking) streams by itself; any other producer stream is named by the caller
		// The kernels read the caller's buffer in place (component stride n_local): no engine-side copy.  Ordering against the
		// producer: the legacy default stream synchronises with the engine's (blocking) streams by itself, both ways; any other
		// producer stream is named by the caller and is made to wait, after the frame's kernels, before it may refill d_soa
		slot = h->stage_next; h->stage_next ^= 1;
		if (producer_stream) {
			CU(cudaEventRecord(h->ev_prod, (cudaStream_t)producer_stream));
			CU(cudaStreamWaitEvent(h->s_main, h->ev_prod, 0));
		}
		const int rcf = run_frame(h, slot, comp[0], n_local, n_total, xlo, xhi);
		if (rcf >= 0 && producer_stream) {
			CU(cudaEventRecord(h->ev_prod, h->s_main));
			CU(cudaStreamWaitEvent((cudaStream_t)producer_stream, h->ev_prod, 0));
		}
		return rcf;
	} else {
		const double *src[9];
		for (int a = 0; a < h->ncomp; ++a) src[a] = comp[a];
		if (comp[0] == h->h_stage[0] && h->stage_cap) slot = 0;
		else if (comp[0] == h->h_stage[1] && h->stage_cap) slot = 1;
		else {
			cudaPointerAttributes pa;
			bool pinned = cudaPointerGetAttributes(&pa, comp[0]) == cudaSuccess && pa.type == cudaMemoryTypeHost;
			cudaGetLastError(); // pageable memory makes the query fail on some drivers: not an error here
			if (pinned) { slot = h->stage_next; h->stage_next ^= 1; }
			else { // pageable host memory: one extra host copy into pinned staging
				double *st = sqt_stage_acquire(h, n_local);
				if (!st) return SQT_ERR_CUDA;
				for (int a = 0; a < h->ncomp; ++a) {
					memcpy(st + (size_t)a * n_local, comp[a], sizeof(double) * (size_t)n_local);
					src[a] = st + (size_t)a * n_local;
				}
				slot = (st == h->h_stage[0]) ? 0 : 1;
			}
		}
		// cur[slot] is free once the kernels that read it two frames ago have run
		CU(cudaStreamWaitEvent(h->s_copy, h->ev_ring[slot], 0));
		double *land = h->n_types ? h->d_raw[slot] : h->d_cur[slot]; // (per-type partials: dump order first, run_frame regroups)
		const long long land_pitch = h->n_types ? n_local : h->cur_pitch;
		for (int a = 0; a < h->ncomp; ++a)
			CU(cudaMemcpyAsync(land + a * land_pitch, src[a], sizeof(double) * n_local, cudaMemcpyHostToDevice, h->s_copy));
		CU(cudaEventRecord(h->ev_copy[slot], h->s_copy));
		h->copy_pending[slot] = true;
		CU(cudaStreamWaitEvent(h->s_main, h->ev_copy[slot], 0));
	}
	return h->n_types ? run_frame(h, slot, h->d_raw[slot], n_local, n_total, xlo, xhi) : run_frame(h, slot, h->d_cur[slot], h->cur_pitch, n_total, xlo, xhi);
}

extern "C" int sqt_set_types(sqt_handle *h, const int *types, int n_atoms_local)
{
	if (!h || !types || n_atoms_local < 1) return h ? fail(h, SQT_ERR_ARG, "sqt_set_types: bad arguments") : SQT_ERR_ARG;
	if (!h->n_types) return fail(h, SQT_ERR_STATE, "sqt_set_types: the handle was created without n_types > 1");
	if (h->sized) return fail(h, SQT_ERR_STATE, "sqt_set_types: the species must be given before the first frame");
	for (int i = 0; i < n_atoms_local; ++i)
		if (types[i] < 1 || types[i] > h->n_types) return fail(h, SQT_ERR_ARG, "sqt_set_types: atom %d has type %d outside 1 .. %d", i, types[i], h->n_types);
	h->h_types.assign(types, types + n_atoms_local);
	return SQT_OK;
}

extern "C" int sqt_push_frame(sqt_handle *h, const double *soa, int n_local, long long n_total, double xlo, double xhi)
{
	if (!h || !soa || n_local < 1) return h ? fail(h, SQT_ERR_ARG, "push_frame: bad arguments") : SQT_ERR_ARG;
	const double *comp[9];
	for (int a = 0; a < 9; ++a) comp[a] = a < h->ncomp ? soa + (size_t)a * n_local : nullptr;
	return push_common(h, comp, false, nullptr, n_local, n_total, xlo, xhi);
}

extern "C" int sqt_push_frame_xyz(sqt_handle *h, const double *x, const double *y, const double *z, const double *vx, const double *vy,
                                  const double *vz, int n_local, long long n_total, double xlo, double xhi)
{
	return sqt_push_frame_xyz_vm(h, x, y, z, vx, vy, vz, nullptr, nullptr, nullptr, n_local, n_total, xlo, xhi);
}

extern "C" int sqt_push_frame_xyz_vm(sqt_handle *h, const double *x, const double *y, const double *z, const double *vx, const double *vy,
                                     const double *vz, const double *mux, const double *muy, const double *muz, int n_local, long long n_total,
                                     double xlo, double xhi)
{
	if (!h) return SQT_ERR_ARG;
	if (!y || !z) return fail(h, SQT_ERR_ARG, "push_frame_xyz: null coordinate array");
	const double *comp[9] = { x, y, z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	if (h->comp_v >= 0) { comp[h->comp_v] = vx; comp[h->comp_v + 1] = vy; comp[h->comp_v + 2] = vz; }
	if (h->comp_m >= 0) { comp[h->comp_m] = mux; comp[h->comp_m + 1] = muy; comp[h->comp_m + 2] = muz; }
	return push_common(h, comp, false, nullptr, n_local, n_total, xlo, xhi);
}

extern "C" int sqt_push_frame_device(sqt_handle *h, const double *d_soa, int n_local, long long n_total, double xlo, double xhi)
{
	const double *comp[9] = { d_soa, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	return push_common(h, comp, true, nullptr, n_local, n_total, xlo, xhi);
}

extern "C" int sqt_push_frame_device_ex(sqt_handle *h, const double *d_soa, int n_local, long long n_total, double xlo, double xhi,
                                        void *producer_stream)
{
	const double *comp[9] = { d_soa, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	return push_common(h, comp, true, producer_stream, n_local, n_total, xlo, xhi);
}

cudaStream_t sqt_internal_copy_stream(sqt_handle *h) { return h->s_copy; }

extern "C" int sqt_sync(sqt_handle *h)
{
	if (!h) return SQT_ERR_ARG;
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->s_copy));
	CU(cudaStreamSynchronize(h->s_main));
	if (h->s_coll != h->s_main) CU(cudaStreamSynchronize(h->s_coll));
	return SQT_OK;
}

static int copy_results(sqt_handle *h, const double *d_blk, sqt_results *out, int count, int limit)
{
	if (!out || out->nCSpatial != h->S || out->nCTime != h->T) return fail(h, SQT_ERR_ARG, "results struct not allocated for this deck");
	int rc = sqt_sync(h);
	if (rc) return rc;
	const size_t ST = (size_t)h->S * h->T, T = h->T;
	double *a2[SQT_N2D] = { out->avF_qq2, out->StdDevF_qq2, out->avF_s_qq2, out->StdDevF_s_qq2, out->avF_d_qq2, out->StdDevF_d_qq2, out->avDrTable,
	                        out->avC2_v_rho, out->avC2_rho_v, out->avC2_v_v, out->avC2_v_mu, out->avC2_mu_v, out->avC2_mu_rho, out->avC2_rho_mu, out->avC2_mu_mu };
	double *a1[6] = { out->rrMSDAv, out->rrMQDAv, out->rrMSDCMAv, out->rrCvvAv, out->rrCvcmvcmAv, out->rrCmmAv };
	for (int i = 0; i < SQT_N2D; ++i) if (a2[i]) CU(cudaMemcpy(a2[i], d_blk + i * ST, sizeof(double) * ST, cudaMemcpyDeviceToHost));
	for (int i = 0; i < 6; ++i) if (a1[i]) CU(cudaMemcpy(a1[i], d_blk + SQT_N2D * ST + i * T, sizeof(double) * T, cudaMemcpyDeviceToHost));
	if (h->flag_s && out->rrMSR1_R_Av && out->rrMSR2_VR_Av) { // device [component][T] -> the reference's VecR3[T] / Rank2R3[T]
		std::vector<double> tmp(12 * T);
		CU(cudaMemcpy(tmp.data(), d_blk + SQT_N2D * ST + 6 * T, sizeof(double) * 12 * T, cudaMemcpyDeviceToHost));
		for (size_t t = 0; t < T; ++t) {
			for (int j = 0; j < 3; ++j) out->rrMSR1_R_Av[3 * t + j] = tmp[j * T + t];
			for (int j = 0; j < 9; ++j) out->rrMSR2_VR_Av[9 * t + j] = tmp[(3 + j) * T + t];
		}
	}
	if (h->n_types && out->avF_ab) { // the species' sums, allocated by sqt_results_alloc_types (out->n_types = what was allocated)
		if (out->n_types != h->n_types)
			return fail(h, SQT_ERR_ARG, "results struct allocated for %d species, the handle has %d", out->n_types, h->n_types);
		const size_t nt = h->n_types;
		const double *tb = d_blk + h->tav_off;
		CU(cudaMemcpy(out->avF_ab, tb, sizeof(double) * nt * nt * ST, cudaMemcpyDeviceToHost));
		CU(cudaMemcpy(out->avF_s_a, tb + nt * nt * ST, sizeof(double) * nt * ST, cudaMemcpyDeviceToHost));
		CU(cudaMemcpy(out->avDrTable_a, tb + nt * nt * ST + nt * ST, sizeof(double) * nt * ST, cudaMemcpyDeviceToHost));
		CU(cudaMemcpy(out->rrMSD_a, tb + nt * nt * ST + 2 * nt * ST, sizeof(double) * nt * T, cudaMemcpyDeviceToHost));
		CU(cudaMemcpy(out->rrMQD_a, tb + nt * nt * ST + 2 * nt * ST + nt * T, sizeof(double) * nt * T, cudaMemcpyDeviceToHost));
		for (size_t t = 0; t < nt; ++t) out->nPtls_a[t] = h->type_total[t];
	}
	out->countCorrAv = count;
	out->limitCorrAv = limit;
	out->nPtls = h->nPtls;
	out->g_Vol = h->vol;
	out->kVal = h->kVal;
	out->first_Vol = h->first_vol;
	out->first_limitCorrAv = h->first_limit;
	out->number_call_print = h->n_print;
	return SQT_OK;
}

extern "C" int sqt_fetch_results(sqt_handle *h, sqt_results *out)
{
	if (!h) return SQT_ERR_ARG;
	if (h->latched) {
		int rc = copy_results(h, h->d_latblk, out, h->latched_count, h->latched_limit);
		if (rc == SQT_OK) h->latched = false;
		return rc;
	}
	return copy_results(h, h->d_avblk, out, h->countCorrAv, h->limitCorrAv);
}

// main tail, lmp2sqt.c:327-333.
extern "C" int sqt_finish(sqt_handle *h, sqt_results *out)
{
	if (!h) return SQT_ERR_ARG;
	if (h->n_print != 0) { int rc = sqt_sync(h); return rc ? rc : 0; }
	h->limitCorrAv = h->countCorrAv;
	h->n_print++;
	int rc = copy_results(h, h->d_avblk, out, h->countCorrAv, h->limitCorrAv);
	if (rc) return rc;
	CU(cudaMemset(h->d_avblk, 0, sizeof(double) * h->avblk_n)); // ZeroAvSpacetimeCorr after the print
	h->countCorrAv = 0;
	return 1;
}

extern "C" int sqt_count_corr_av(const sqt_handle *h) { return h ? h->countCorrAv : SQT_ERR_ARG; }
extern "C" int sqt_buffer_count(const sqt_handle *h, int nb) { return (h && nb >= 0 && nb < h->B) ? h->count[nb] : -(1 << 30); }
extern "C" int sqt_number_call_print(const sqt_handle *h) { return h ? h->n_print : SQT_ERR_ARG; }
extern "C" long long sqt_kernel_launches(const sqt_handle *h) { return h ? h->launches : 0; }
extern "C" const char *sqt_kernel_config(const sqt_handle *h) { return h ? h->cfg : ""; }

// d_rho_cur (+ offset) of the last processed frame to the host, waiting only for that frame's rho stage: the copy goes
// through a non-blocking stream of its own (a plain cudaMemcpy would synchronise with the engine's blocking streams and
// hence wait for the self-term kernel)
static int read_rho_block(sqt_handle *h, double *out, int offset)
{
	CU(cudaSetDevice(h->device));
	h->rho_reader = true;
	if (h->rhoc_pending) CU(cudaStreamWaitEvent(h->s_aux, h->ev_rhoc, 0));
	CU(cudaMemcpyAsync(out, h->d_rho_cur + offset, sizeof(double) * 6 * h->S, cudaMemcpyDeviceToHost, h->s_aux));
	CU(cudaStreamSynchronize(h->s_aux));
	return SQT_OK;
}

extern "C" int sqt_get_rho(sqt_handle *h, double *out)
{
	if (!h || !out) return SQT_ERR_ARG;
	return read_rho_block(h, out, 0);
}

extern "C" int sqt_get_kvel(sqt_handle *h, double *out)
{
	if (!h || !out) return SQT_ERR_ARG;
	if (!h->flag_V) return fail(h, SQT_ERR_STATE, "sqt_get_kvel: the handle was created without flag_V");
	return read_rho_block(h, out, h->blk_v * h->rho_n);
}

extern "C" int sqt_get_kmu(sqt_handle *h, double *out)
{
	if (!h || !out) return SQT_ERR_ARG;
	if (!h->flag_M) return fail(h, SQT_ERR_STATE, "sqt_get_kmu: the handle was created without flag_M");
	return read_rho_block(h, out, h->blk_m * h->rho_n);
}

extern "C" int sqt_frame_components(const sqt_handle *h) { return h ? h->ncomp : SQT_ERR_ARG; }

extern "C" int sqt_profile_enable(sqt_handle *h, int enable)
{
	if (!h) return SQT_ERR_ARG;
	h->prof = enable != 0;
	if (h->prof) { // four events per frame; prepared here so that frames pushed while profiling create nothing
		CU(cudaSetDevice(h->device));
		while (h->prof_pool.size() < 4 * 4096) {
			cudaEvent_t e = nullptr;
			CU(cudaEventCreate(&e));
			h->prof_pool.push_back(e);
		}
	}
	return SQT_OK;
}

extern "C" int sqt_profile_read(sqt_handle *h, long long *n_launches, double *ms_total)
{
	if (!h) return SQT_ERR_ARG;
	int rc = sqt_sync(h);
	if (rc) return rc;
	double tot = 0;
	for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
		float ms = 0;
		CU(cudaEventElapsedTime(&ms, h->prof_pool[i], h->prof_pool[i + 1]));
		tot += ms;
	}
	// one pair around the rho(k) kernels (when they are separate launches) and one around the self-term kernel per frame
	const size_t per_frame = h->rmc.on ? 4 : 2;
	if (n_launches) *n_launches = (long long)(h->prof_used / per_frame);
	if (ms_total) *ms_total = tot;
	h->prof_used = 0;
	return SQT_OK;
}

// Device-side wall clock of a region of frames: start is recorded on the main stream, stop after
// the collective stage of the last frame pushed.
extern "C" int sqt_timer_start(sqt_handle *h)
{
	if (!h) return SQT_ERR_ARG;
	CU(cudaSetDevice(h->device));
	CU(cudaEventRecord(h->ev_t0, h->s_main));
	return SQT_OK;
}

extern "C" int sqt_timer_stop(sqt_handle *h, double *ms)
{
	if (!h) return SQT_ERR_ARG;
	CU(cudaSetDevice(h->device));
	if (h->s_coll != h->s_main && h->coll_pending) CU(cudaStreamWaitEvent(h->s_main, h->ev_coll, 0));
	CU(cudaEventRecord(h->ev_t1, h->s_main));
	CU(cudaEventSynchronize(h->ev_t1));
	float f = 0;
	CU(cudaEventElapsedTime(&f, h->ev_t0, h->ev_t1));
	if (ms) *ms = f;
	return SQT_OK;
}

extern "C" int sqt_comm_unique_id(void *id_out)
{
	if (!id_out) return SQT_ERR_ARG;
	return sqt_comm_get_unique_id(id_out, g_create_err, sizeof(g_create_err)) ? SQT_ERR_NCCL : SQT_OK;
}

extern "C" int sqt_comm_init(sqt_handle *h, const void *id, int rank, int nranks)
{
	if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return SQT_ERR_ARG;
	if (h->comm) return fail(h, SQT_ERR_STATE, "communicator already attached");
	if (h->sized && nranks > 1) return fail(h, SQT_ERR_STATE, "attach the communicator before the first frame (the kernel configuration depends on it)");
	if (nranks == 1) return SQT_OK;
	CU(cudaSetDevice(h->device));
	if (sqt_comm_create(&h->comm, id, rank, nranks, h->err, sizeof(h->err))) return SQT_ERR_NCCL;
	h->rank = rank; h->nranks = nranks;
	CU(cudaStreamCreate(&h->s_coll));
	return SQT_OK;
}
