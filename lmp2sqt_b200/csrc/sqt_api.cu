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
	char cfg[512];

	// ---- ring bookkeeping (host mirror of TBuf.count etc.) ----
	int limitCorrAv, countCorrAv, n_print;
	int nSkip, p_file;
	std::vector<int> count;
	std::vector<double> h_invL0;
	long long nPtls;
	int n_local;
	bool sized, first_done;
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
	long long *d_thr;
	double *d_avblk, *d_latblk; // [7*S*T + 3*T]
	size_t avblk_n;
	bool latched;
	int latched_count, latched_limit;
	KernelChoice kc;
	MmaChoice mc;             // blocked-Chebyshev / DMMA path for the real slots (large nCSpatial)
	double *d_mma_partial;
	RhoMmaChoice rmc;
	VecChoice vc;             // particle-per-lane vector kernel (nCSpatial <= 32)
	double *d_rho_partial;
	double *d_mma_extra, *d_rho_extra; // per-warp sums kept outside the DMMA product (MmaChoice.SPOUT)
	int *d_mma_ctr;
	cudaStream_t s_main, s_copy, s_coll;
	cudaEvent_t ev_red, ev_coll, ev_t0, ev_t1;
	bool coll_pending;
	SqtComm *comm;
	int rank, nranks;
	long long launches;
	bool prof;
	std::vector<cudaEvent_t> prof_ev; // pairs
};

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
	double **a2[7] = { &r->avF_qq2, &r->StdDevF_qq2, &r->avF_s_qq2, &r->StdDevF_s_qq2, &r->avF_d_qq2, &r->StdDevF_d_qq2, &r->avDrTable };
	double **a1[3] = { &r->rrMSDAv, &r->rrMQDAv, &r->rrMSDCMAv };
	for (int i = 0; i < 7; ++i) if (!(*a2[i] = (double *)calloc((size_t)S * T, sizeof(double)))) return SQT_ERR_NOMEM;
	for (int i = 0; i < 3; ++i) if (!(*a1[i] = (double *)calloc(T, sizeof(double)))) return SQT_ERR_NOMEM;
	return SQT_OK;
}

extern "C" void sqt_results_free(sqt_results *r)
{
	if (!r) return;
	free(r->avF_qq2); free(r->StdDevF_qq2); free(r->avF_s_qq2); free(r->StdDevF_s_qq2);
	free(r->avF_d_qq2); free(r->StdDevF_d_qq2); free(r->avDrTable);
	free(r->rrMSDAv); free(r->rrMQDAv); free(r->rrMSDCMAv);
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
	if (!(p->rVal > 0)) return fail(h, SQT_ERR_ARG, "sqt_create: rVal must be > 0");
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
	h->limitCorrAv = p->limitCorrAv;
	h->countCorrAv = 0; h->n_print = 0; h->nSkip = 0; h->p_file = 0;
	h->count.assign(h->B, 0);
	h->h_invL0.assign(h->B, 0.0);
	h->nPtls = 0; h->n_local = 0; h->sized = false; h->first_done = false;
	h->L = h->vol = h->first_vol = 0; h->kVal = p->kVal; h->first_limit = 0;
	h->d_cur[0] = h->d_cur[1] = nullptr; h->h_stage[0] = h->h_stage[1] = nullptr; h->stage_cap = 0;
	h->copy_pending[0] = h->copy_pending[1] = false; h->stage_next = 0;
	h->d_ring = nullptr; h->d_partial = nullptr; h->d_tile_ctr = nullptr; h->d_mma_partial = nullptr; h->d_mma_ctr = nullptr; h->mc.on = 0; h->mc.geo = 0; h->vc.on = 0; h->rmc.on = 0; h->d_rho_partial = nullptr; h->d_mma_extra = h->d_rho_extra = nullptr; h->d_hist_cur = h->d_bufHist = nullptr;
	h->d_bufFs = h->d_bufF = h->d_bufMSD = h->d_bufMQD = h->d_bufMSDCM = nullptr;
	h->d_rho0 = h->d_sumr0 = h->d_invL0 = h->d_rho_cur = h->d_cblk = nullptr; h->d_thr = nullptr;
	h->d_avblk = h->d_latblk = nullptr;
	h->latched = false; h->latched_count = 0; h->latched_limit = 0;
	h->comm = nullptr; h->rank = 0; h->nranks = 1; h->launches = 0; h->prof = false; h->coll_pending = false;
	h->s_main = h->s_copy = h->s_coll = nullptr;

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
	h->s_coll = h->s_main; // a separate collective stream is created when a communicator is attached
	for (int i = 0; i < 2; ++i) {
		CUC(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&h->ev_ring[i], cudaEventDisableTiming));
	}
	CUC(cudaEventCreateWithFlags(&h->ev_red, cudaEventDisableTiming));
	CUC(cudaEventCreateWithFlags(&h->ev_coll, cudaEventDisableTiming));
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
	CUC(cudaMalloc(&h->d_rho_cur, sizeof(double) * (6 * S + 3)));
	CUC(cudaMemset(h->d_rho_cur, 0, sizeof(double) * (6 * S + 3)));
	CUC(cudaMalloc(&h->d_cblk, sizeof(double) * (2 * T * S + 2 * T)));
	h->avblk_n = 7 * S * T + 3 * T;
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

extern "C" void sqt_destroy(sqt_handle *h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	cudaDeviceSynchronize();
	if (h->comm) sqt_comm_destroy(h->comm);
	for (int i = 0; i < 2; ++i) {
		cudaFree(h->d_cur[i]);
		if (h->h_stage[i]) cudaFreeHost(h->h_stage[i]);
		if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
		if (h->ev_ring[i]) cudaEventDestroy(h->ev_ring[i]);
	}
	cudaFree(h->d_ring); cudaFree(h->d_partial); cudaFree(h->d_tile_ctr); cudaFree(h->d_mma_partial); cudaFree(h->d_mma_ctr); cudaFree(h->d_rho_partial); cudaFree(h->d_mma_extra); cudaFree(h->d_rho_extra); cudaFree(h->d_hist_cur); cudaFree(h->d_bufHist);
	cudaFree(h->d_bufFs); cudaFree(h->d_bufF); cudaFree(h->d_bufMSD); cudaFree(h->d_bufMQD); cudaFree(h->d_bufMSDCM);
	cudaFree(h->d_rho0); cudaFree(h->d_sumr0); cudaFree(h->d_invL0); cudaFree(h->d_rho_cur); cudaFree(h->d_cblk);
	cudaFree(h->d_thr); cudaFree(h->d_avblk); cudaFree(h->d_latblk);
	for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
	if (h->ev_red) cudaEventDestroy(h->ev_red);
	if (h->ev_coll) cudaEventDestroy(h->ev_coll);
	if (h->ev_t0) cudaEventDestroy(h->ev_t0);
	if (h->ev_t1) cudaEventDestroy(h->ev_t1);
	if (h->s_coll && h->s_coll != h->s_main) cudaStreamDestroy(h->s_coll);
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
		CU(cudaMallocHost(&h->h_stage[i], sizeof(double) * 3 * (size_t)n));
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
// size depends on the atom count.
static int size_for_atoms(sqt_handle *h, int n_local, long long n_total)
{
	CU(cudaSetDevice(h->device));
	h->n_local = n_local;
	h->nPtls = n_total;
	h->cur_pitch = ((long long)n_local + 31) & ~31LL;
	for (int i = 0; i < 2; ++i) CU(cudaMalloc(&h->d_cur[i], sizeof(double) * 3 * h->cur_pitch));
	int rc = sqt_pick_mma(h->S, h->B, n_local, h->n_sm, &h->mc);
	if (rc != 0) return fail(h, SQT_ERR_CUDA, "mma kernel selection failed (rc=%d)", rc);
	if (!h->mc.on) {
		rc = sqt_pick_vec(h->S, h->B, n_local, h->n_sm, &h->vc);
		if (rc != 0) return fail(h, SQT_ERR_CUDA, "vector kernel selection failed (rc=%d)", rc);
		if (h->vc.on && !h->mc.geo) h->vc.on = 0; // rho(k) comes from the tensor-core rho kernel on this path
	}
	const bool lane_family = h->mc.on || h->vc.on; // a particle per lane: particle-minor ring, rho(k) from rho_mma_kernel
	// with a particle-per-lane self kernel on, the classic kernel serves at most the six virtual rho(k) slots (B = 0 for it)
	rc = sqt_pick_kernel(h->S, lane_family ? 0 : h->B, n_local, h->n_sm, &h->kc);
	if (rc != 0) return fail(h, SQT_ERR_CUDA, "no ring-kernel variant for nCSpatial=%d (rc=%d)", h->S, rc);
	h->RP = (h->B + 15) & ~15; // 128-byte rows: a warp's 32 origins of one particle/axis are contiguous
	// the tensor-core kernel walks particles along the warp, so its ring is [slot][axis][particle] (pitch: a multiple of 32
	// particles); the classic kernel walks slots along the warp: [particle][axis][slot]
	h->ring_np = ((long long)n_local + 31) & ~31LL;
	const size_t ring_bytes = lane_family ? sizeof(double) * 3 * (size_t)h->B * (size_t)h->ring_np : sizeof(double) * 3 * (size_t)h->RP * (size_t)n_local;
	cudaError_t e = cudaMalloc(&h->d_ring, ring_bytes);
	if (e != cudaSuccess) return fail(h, SQT_ERR_NOMEM, "origin ring of %.2f GB does not fit on the device: %s", ring_bytes / 1e9, cudaGetErrorString(e));
	CU(cudaMemsetAsync(h->d_ring, 0, ring_bytes, h->s_main));
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
		CU(cudaMalloc(&h->d_mma_ctr, sizeof(int) * nwsg));
		CU(cudaMemsetAsync(h->d_mma_ctr, 0, sizeof(int) * nwsg, h->s_main));
		const size_t xb = h->mc.on ? sizeof(double) * (size_t)h->mc.NWSG * h->mc.KW * h->mc.SPW * 2 : 16;
		CU(cudaMalloc(&h->d_mma_extra, xb));
		CU(cudaMemsetAsync(h->d_mma_extra, 0, xb, h->s_main));
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
	int off = 0;
	if (h->vc.on)
		off = snprintf(h->cfg, sizeof(h->cfg),
		               "self_ring_vec_kernel<C=%d,SPW=%d> (a particle per lane, sums in registers) persistent grid=%d block=%d regs=%d smem=%zuB occ=%d blocks/SM; "
		               "%d slot groups x %d warps, %s tiles of %d particles; + rho(k): ",
		               h->vc.C, h->vc.SPW, h->vc.blocks, h->vc.WPB * 32, h->vc.regs, h->vc.smem, h->vc.occ, h->vc.NWSG, h->vc.KW,
		               h->vc.dynamic ? "dynamic" : "static", h->vc.tile);
	if (h->mc.on)
		off = snprintf(h->cfg, sizeof(h->cfg),
		               "self_ring_mma_kernel<NT=%d,SPW=%d> (blocked Chebyshev S=%d on DMMA%s) persistent grid=%d block=%d regs=%d smem=%zuB occ=%d blocks/SM; "
		               "%d slot groups x %d warps, %s tiles of %d particles; + rho(k): ",
		               h->mc.NT, h->mc.SPW, h->mc.S, h->mc.SPOUT ? ", MSD/MQD outside" : "", h->mc.blocks, h->mc.WPB * 32, h->mc.regs, h->mc.smem, h->mc.occ, h->mc.NWSG, h->mc.KW,
		               h->mc.dynamic ? "dynamic" : "static", h->mc.tile);
	if (h->rmc.on)
		off += snprintf(h->cfg + off, sizeof(h->cfg) - off, "rho_mma_kernel<NT=%d> grid=%d block=%d smem=%zuB occ=%d blocks/SM, static tiles of 32 particles; unused: ",
		                h->mc.NT, h->rmc.blocks, h->rmc.WPB * 32, h->rmc.smem, h->rmc.occ);
	snprintf(h->cfg + off, sizeof(h->cfg) - off,
	         "self_ring_kernel<C=%d,G=%d> persistent grid=%d block=%d regs=%d smem=%zuB occ=%d blocks/SM; %d slot groups x %d warps, "
	         "%s tiles of %d particles (%d tiles); slots=%d(+%d virtual) ring=%.3fGB partials=%.1fMB",
	         h->kc.C, h->kc.G, h->kc.blocks, h->kc.WPB * 32, h->kc.regs, h->kc.smem, h->kc.occ, h->kc.NWSG, h->kc.KW,
	         h->kc.dynamic ? "dynamic" : "static", h->kc.tile, h->kc.ntiles, h->B, SQT_NVIRT, ring_bytes / 1e9, (part_bytes + mma_bytes) / 1e6);
	h->sized = true;
	return SQT_OK;
}

static int run_frame(sqt_handle *h, int slot, long long n_total, double xlo, double xhi)
{
	const int S = h->S, T = h->T, B = h->B;
	h->L = xhi - xlo;                      // lmp2sqt.c:1620
	h->vol = h->L * h->L * h->L;           // :1621
	h->nPtls = n_total;                    // :1622
	if (!h->first_done) { h->first_vol = h->vol; h->first_limit = h->limitCorrAv; h->first_done = true; } // factorDr inputs, :1991-2004
	h->kVal = 2. * M_PI / h->L;            // :1632
	const double invL = 1.0 / h->L;
	const int p = h->p_file;

	int varbox = 0;
	for (int nb = 0; nb < B; ++nb)
		if (h->count[nb] > 0 && h->h_invL0[nb] != invL) { varbox = 1; break; }

	const int Bc = (h->mc.on || h->vc.on) ? 0 : B; // slots the classic kernel treats as real
	RingArgs ra;
	ra.cur = h->d_cur[slot]; ra.ring = h->d_ring; ra.partial = h->d_partial; ra.hist_cur = h->d_hist_cur;
	ra.invL0 = h->d_invL0; ra.thr = h->d_thr; ra.cur_pitch = h->cur_pitch; ra.n_local = h->n_local;
	ra.tile = h->kc.tile; ra.ntiles = h->kc.ntiles; ra.NWSG = h->kc.NWSG; ra.KW = h->kc.KW; ra.total_warps = h->kc.total_warps;
	ra.dynamic = h->kc.dynamic; ra.tile_ctr = h->d_tile_ctr; ra.p = p; ra.T = T; ra.B = Bc; ra.nCS = S; ra.RP = h->RP; ra.V0 = h->kc.V0;
	ra.SPAD = h->kc.SPAD; ra.NV = h->kc.NV; ra.CT = h->kc.CT; ra.invL = invL; ra.inv_rVal = (float)(1.0 / h->p.rVal);
	ra.varbox = varbox;
	cudaEvent_t pe0 = nullptr, pe1 = nullptr;
	if (h->prof) {
		CU(cudaEventCreate(&pe0)); CU(cudaEventCreate(&pe1));
		CU(cudaEventRecord(pe0, h->s_main));
	}
	int rc;
	if (h->vc.on) {
		MmaArgs ma;
		ma.cur = h->d_cur[slot]; ma.ring = h->d_ring; ma.partial = h->d_mma_partial; ma.extra = h->d_mma_extra; ma.tile_ctr = h->d_mma_ctr; ma.hist_cur = h->d_hist_cur;
		ma.invL0 = h->d_invL0; ma.thr = h->d_thr; ma.cur_pitch = h->cur_pitch; ma.n_local = h->n_local; ma.tile = h->vc.tile;
		ma.ntiles = h->vc.ntiles; ma.NWSG = h->vc.NWSG; ma.KW = h->vc.KW; ma.total_warps = h->vc.total_warps; ma.dynamic = h->vc.dynamic;
		ma.p = p; ma.T = T; ma.B = B; ma.nCS = S; ma.RP = h->ring_np; ma.S = 0; ma.NT = 0; ma.invL = invL;
		ma.inv_rVal = (float)(1.0 / h->p.rVal); ma.varbox = varbox;
		rc = sqt_launch_vec(h->vc, ma, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "self_ring_vec_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		h->launches++;
	}
	if (h->mc.on) {
		MmaArgs ma;
		ma.cur = h->d_cur[slot]; ma.ring = h->d_ring; ma.partial = h->d_mma_partial; ma.extra = h->d_mma_extra; ma.tile_ctr = h->d_mma_ctr; ma.hist_cur = h->d_hist_cur;
		ma.invL0 = h->d_invL0; ma.thr = h->d_thr; ma.cur_pitch = h->cur_pitch; ma.n_local = h->n_local; ma.tile = h->mc.tile;
		ma.ntiles = h->mc.ntiles; ma.NWSG = h->mc.NWSG; ma.KW = h->mc.KW; ma.total_warps = h->mc.total_warps; ma.dynamic = h->mc.dynamic;
		ma.p = p; ma.T = T; ma.B = B; ma.nCS = S; ma.RP = h->ring_np; ma.S = h->mc.S; ma.NT = h->mc.NT; ma.invL = invL;
		ma.inv_rVal = (float)(1.0 / h->p.rVal); ma.varbox = varbox;
		rc = sqt_launch_mma(h->mc, ma, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "self_ring_mma_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		h->launches++;
	}
	const bool split = h->s_coll != h->s_main;
	if (h->rmc.on) { // rho(k) and sum_n r on the tensor-core path too: the classic kernel is not launched at all
		// its reduce rewrites d_rho_cur, which the previous frame's allreduce + collective stage (other stream) still reads
		if (split && h->coll_pending) CU(cudaStreamWaitEvent(h->s_main, h->ev_coll, 0));
		RhoMmaArgs qa;
		qa.cur = h->d_cur[slot]; qa.partial = h->d_rho_partial; qa.extra = h->d_rho_extra; qa.spout = h->mc.SPOUT; qa.rho_cur = h->d_rho_cur; qa.cur_pitch = h->cur_pitch;
		qa.n_local = h->n_local; qa.total_warps = h->rmc.total_warps; qa.n_sets = h->rmc.blocks; qa.nCS = S; qa.S = h->mc.S; qa.NT = h->mc.NT; qa.invL = invL;
		rc = sqt_launch_rho_mma(h->mc, h->rmc, qa, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "rho_mma_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		h->launches++;
	} else {
		rc = sqt_launch_ring(h->kc, ra, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "self_ring_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
	}
	if (h->prof) { CU(cudaEventRecord(pe1, h->s_main)); h->prof_ev.push_back(pe0); h->prof_ev.push_back(pe1); }
	CU(cudaEventRecord(h->ev_ring[slot], h->s_main));
	h->launches++;

	if (split && h->coll_pending) CU(cudaStreamWaitEvent(h->s_main, h->ev_coll, 0)); // previous frame's collective stage
	ReduceArgs rr;
	rr.partial = h->d_partial; rr.KW = h->kc.KW; rr.NWSG = h->kc.NWSG; rr.tile_ctr = h->d_tile_ctr; rr.NV = h->kc.NV; rr.SPAD = h->kc.SPAD; rr.CT = h->kc.CT; rr.V0 = h->kc.V0; rr.G = h->kc.G;
	rr.p = p; rr.T = T; rr.B = Bc; rr.nCS = S; rr.hist_cur = h->d_hist_cur; rr.bufHist = h->d_bufHist; rr.bufFs = h->d_bufFs;
	rr.bufMSD = h->d_bufMSD; rr.bufMQD = h->d_bufMQD; rr.rho_cur = h->d_rho_cur; rr.invL0 = h->d_invL0; rr.invL = invL;
	if (h->vc.on) {
		MmaReduceArgs mr;
		mr.partial = h->d_mma_partial; mr.extra = h->d_mma_extra; mr.spout = 0; mr.n_local = h->n_local; mr.KW = h->vc.KW; mr.NWSG = h->vc.NWSG;
		mr.S = 0; mr.NT = h->vc.C + 2; mr.SPW = h->vc.SPW;
		mr.p = p; mr.T = T; mr.B = B; mr.nCS = S; mr.tile_ctr = h->d_mma_ctr; mr.hist_cur = h->d_hist_cur; mr.bufHist = h->d_bufHist;
		mr.bufFs = h->d_bufFs; mr.bufMSD = h->d_bufMSD; mr.bufMQD = h->d_bufMQD; mr.invL0 = h->d_invL0; mr.invL = invL;
		rc = sqt_launch_vec_reduce(mr, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "vec_reduce_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		h->launches++;
	}
	if (h->mc.on) {
		MmaReduceArgs mr;
		mr.partial = h->d_mma_partial; mr.extra = h->d_mma_extra; mr.spout = h->mc.SPOUT; mr.n_local = h->n_local; mr.KW = h->mc.KW; mr.NWSG = h->mc.NWSG; mr.S = h->mc.S; mr.NT = h->mc.NT; mr.SPW = h->mc.SPW;
		mr.p = p; mr.T = T; mr.B = B; mr.nCS = S; mr.tile_ctr = h->d_mma_ctr; mr.hist_cur = h->d_hist_cur; mr.bufHist = h->d_bufHist;
		mr.bufFs = h->d_bufFs; mr.bufMSD = h->d_bufMSD; mr.bufMQD = h->d_bufMQD; mr.invL0 = h->d_invL0; mr.invL = invL;
		rc = sqt_launch_mma_reduce(mr, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "mma_reduce_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		h->launches++;
	}
	if (!h->rmc.on) {
		rc = sqt_launch_reduce(rr, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "frame_reduce_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		h->launches++;
	}
	if (split) {
		CU(cudaEventRecord(h->ev_red, h->s_main));
		CU(cudaStreamWaitEvent(h->s_coll, h->ev_red, 0));
	}
	if (h->comm && sqt_comm_allreduce_f64(h->comm, h->d_rho_cur, 6 * (size_t)S + 3, h->s_coll, h->err, sizeof(h->err)))
		return SQT_ERR_NCCL;
	CollectiveArgs ca;
	ca.rho_cur = h->d_rho_cur; ca.rho0 = h->d_rho0; ca.sumr0 = h->d_sumr0; ca.bufF = h->d_bufF; ca.bufMSDCM = h->d_bufMSDCM;
	ca.p = p; ca.T = T; ca.B = B; ca.nCS = S;
	ca.inv_n2 = 1.0 / ((double)n_total * (double)n_total);
	rc = sqt_launch_collective(ca, h->s_coll);
	if (rc) return fail(h, SQT_ERR_CUDA, "collective_ring_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
	h->launches++;

	// ---- ring state machine, lmp2sqt.c:1638-1647 -------------------------------------------
	for (int nb = 0; nb < B; ++nb) {
		if (h->count[nb] == 0) h->h_invL0[nb] = invL;
		++h->count[nb];
	}
	// ---- AccumSpacetimeCorr, lmp2sqt.c:376-443 ---------------------------------------------
	int status = SQT_OK;
	for (int nb = 0; nb < B; ++nb) {
		if (h->count[nb] != T) continue;
		AccumArgs aa;
		aa.nb = nb; aa.T = T; aa.nCS = S;
		aa.bufF = h->d_bufF; aa.bufFs = h->d_bufFs; aa.bufMSD = h->d_bufMSD; aa.bufMQD = h->d_bufMQD; aa.bufMSDCM = h->d_bufMSDCM;
		aa.bufHist = h->d_bufHist; aa.cblk = h->d_cblk;
		for (int i = 0; i < 7; ++i) aa.av[i] = h->d_avblk + (size_t)i * S * T;
		for (int i = 0; i < 3; ++i) aa.av1[i] = h->d_avblk + (size_t)7 * S * T + (size_t)i * T;
		rc = sqt_launch_pack(aa, h->s_coll);
		if (rc) return fail(h, SQT_ERR_CUDA, "buffer_pack_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		if (h->comm && sqt_comm_allreduce_f64(h->comm, h->d_cblk, 2 * (size_t)T * S + 2 * (size_t)T, h->s_coll, h->err, sizeof(h->err)))
			return SQT_ERR_NCCL;
		rc = sqt_launch_accumulate(aa, h->s_coll);
		if (rc) return fail(h, SQT_ERR_CUDA, "buffer_accumulate_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		h->launches += 2;
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
			status = SQT_FLUSH;
		}
	}
	if (split) { CU(cudaEventRecord(h->ev_coll, h->s_coll)); h->coll_pending = true; }
	h->nSkip = 0;   // lmp2sqt.c:1649
	h->p_file++;
	return status;
}

static int push_common(sqt_handle *h, const double *src, bool on_device, int n_local, long long n_total, double xlo, double xhi)
{
	if (!h || !src || n_local < 1 || n_total < n_local) return h ? fail(h, SQT_ERR_ARG, "push_frame: bad arguments") : SQT_ERR_ARG;
	if (!(xhi > xlo)) return fail(h, SQT_ERR_ARG, "push_frame: box xhi must exceed xlo");
	if (h->nSkip < h->p.nCSkip) { h->nSkip++; return SQT_OK; } // lmp2sqt.c:1612-1615
	CU(cudaSetDevice(h->device));
	if (!h->sized) { int rc = size_for_atoms(h, n_local, n_total); if (rc) return rc; }
	if (n_local != h->n_local || n_total != h->nPtls)
		return fail(h, SQT_ERR_STATE, "atom count changed between frames (%d/%lld -> %d/%lld): the reference sizes its arrays once (lmp2sqt.c:1623-1630)",
		            h->n_local, h->nPtls, n_local, n_total);
	int slot;
	if (on_device) {
		// engine-owned copy so the producer may overwrite its buffer as soon as later work on
		// the legacy default stream is ordered after ours (both engine streams are blocking streams)
		slot = h->stage_next; h->stage_next ^= 1;
		int rc = sqt_launch_soa_copy(h->d_cur[slot], h->cur_pitch, src, n_local, n_local, h->s_main);
		if (rc) return fail(h, SQT_ERR_CUDA, "soa_copy launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		h->launches++;
	} else {
		const double *stage = src;
		if (src == h->h_stage[0] && h->stage_cap) slot = 0;
		else if (src == h->h_stage[1] && h->stage_cap) slot = 1;
		else {
			cudaPointerAttributes pa;
			bool pinned = cudaPointerGetAttributes(&pa, src) == cudaSuccess && pa.type == cudaMemoryTypeHost;
			cudaGetLastError(); // pageable memory makes the query fail on some drivers: not an error here
			if (pinned) { slot = h->stage_next; h->stage_next ^= 1; }
			else { // pageable host memory: one extra host copy into pinned staging
				double *st = sqt_stage_acquire(h, n_local);
				if (!st) return SQT_ERR_CUDA;
				memcpy(st, src, sizeof(double) * 3 * (size_t)n_local);
				stage = st;
				slot = (st == h->h_stage[0]) ? 0 : 1;
			}
		}
		// cur[slot] is free once the ring kernel that read it two frames ago has run
		CU(cudaStreamWaitEvent(h->s_copy, h->ev_ring[slot], 0));
		for (int a = 0; a < 3; ++a)
			CU(cudaMemcpyAsync(h->d_cur[slot] + a * h->cur_pitch, stage + (size_t)a * n_local, sizeof(double) * n_local,
			                   cudaMemcpyHostToDevice, h->s_copy));
		CU(cudaEventRecord(h->ev_copy[slot], h->s_copy));
		h->copy_pending[slot] = true;
		CU(cudaStreamWaitEvent(h->s_main, h->ev_copy[slot], 0));
	}
	return run_frame(h, slot, n_total, xlo, xhi);
}

extern "C" int sqt_push_frame(sqt_handle *h, const double *soa, int n_local, long long n_total, double xlo, double xhi)
{
	return push_common(h, soa, false, n_local, n_total, xlo, xhi);
}

extern "C" int sqt_push_frame_device(sqt_handle *h, const double *d_soa, int n_local, long long n_total, double xlo, double xhi)
{
	return push_common(h, d_soa, true, n_local, n_total, xlo, xhi);
}

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
	double *a2[7] = { out->avF_qq2, out->StdDevF_qq2, out->avF_s_qq2, out->StdDevF_s_qq2, out->avF_d_qq2, out->StdDevF_d_qq2, out->avDrTable };
	double *a1[3] = { out->rrMSDAv, out->rrMQDAv, out->rrMSDCMAv };
	for (int i = 0; i < 7; ++i) CU(cudaMemcpy(a2[i], d_blk + i * ST, sizeof(double) * ST, cudaMemcpyDeviceToHost));
	for (int i = 0; i < 3; ++i) CU(cudaMemcpy(a1[i], d_blk + 7 * ST + i * T, sizeof(double) * T, cudaMemcpyDeviceToHost));
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

extern "C" int sqt_get_rho(sqt_handle *h, double *out)
{
	if (!h || !out) return SQT_ERR_ARG;
	int rc = sqt_sync(h);
	if (rc) return rc;
	CU(cudaMemcpy(out, h->d_rho_cur, sizeof(double) * 6 * h->S, cudaMemcpyDeviceToHost));
	return SQT_OK;
}

extern "C" int sqt_profile_enable(sqt_handle *h, int enable)
{
	if (!h) return SQT_ERR_ARG;
	h->prof = enable != 0;
	return SQT_OK;
}

extern "C" int sqt_profile_read(sqt_handle *h, long long *n_launches, double *ms_total)
{
	if (!h) return SQT_ERR_ARG;
	int rc = sqt_sync(h);
	if (rc) return rc;
	double tot = 0;
	for (size_t i = 0; i + 1 < h->prof_ev.size(); i += 2) {
		float ms = 0;
		CU(cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]));
		tot += ms;
	}
	if (n_launches) *n_launches = (long long)(h->prof_ev.size() / 2);
	if (ms_total) *ms_total = tot;
	for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
	h->prof_ev.clear();
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
	if (nranks == 1) return SQT_OK;
	CU(cudaSetDevice(h->device));
	if (sqt_comm_create(&h->comm, id, rank, nranks, h->err, sizeof(h->err))) return SQT_ERR_NCCL;
	h->rank = rank; h->nranks = nranks;
	CU(cudaStreamCreate(&h->s_coll));
	return SQT_OK;
}
