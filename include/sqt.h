/*
 * sqt.h -- C-ABI of the B200 space-time-correlation engine (libsqt_b200.so).
 *
 * Drop-in boundary for the hot path of infinitekh/lmp2sqt.  The reference has no plugin or
 * FFI interface: it is one translation unit whose `main` makes four calls into the hot path
 * (SURVEY.md section 8b).  Each entry point below replaces one of them; plain pointers and
 * sizes only, no CUDA or torch types.
 *
 *   reference call (file:line)                              replaced by
 *   ------------------------------------------------------  ---------------------------------
 *   AllocArray / Alloc_more      lmp2sqt.c:1821, :1926       sqt_create (+ first sqt_push_frame)
 *   limitCorrAv += floor(...)    lmp2sqt.c:244               sqt_set_limit
 *   InitSpacetimeCorr(cl)        lmp2sqt.c:446, main :282    sqt_begin_file
 *   cl->snap = ReadDump(fp)      lmp2sqt.c:295-298           sqt_stage_acquire (caller fills SoA)
 *   EvalSpacetimeCorr(cl)        lmp2sqt.c:1599, main :300   sqt_push_frame / sqt_push_frame_device
 *   PrintProcess(cl) mid-run     lmp2sqt.c:438-440           sqt_push_frame returns SQT_FLUSH,
 *                                                            then sqt_fetch_results
 *   PrintProcess(cl) end of run  lmp2sqt.c:327-333           sqt_finish
 *
 * Everything numerical between those calls (EvalOneTimeKspace :1144, SetWaitedTimeCorr :1260,
 * EvalTwoTimeEach :1368, EvalTwoTimeKSpace :1477, AccumSpacetimeCorr :368) runs in hand-written
 * sm_100a kernels on device-resident state.  Normalisation (prePrintProcess :576) and the file
 * writers stay on the host (include/sqt_host.h) and consume `sqt_results`.
 *
 * Threading: one caller thread per handle (the reference is single threaded).  Calls are
 * asynchronous with respect to the GPU; sqt_fetch_results / sqt_finish / sqt_sync synchronise.
 * Errors: every function returns SQT_OK (0) or a positive status on success and a negative
 * code on failure; sqt_last_error() gives the message.  There is NO CPU fallback: without a
 * usable CUDA device sqt_create fails with SQT_ERR_CUDA.
 */
#ifndef SQT_H
#define SQT_H

#ifdef __cplusplus
extern "C" {
#endif

#define SQT_OK 0
#define SQT_FLUSH 1          /* countCorrAv reached limitCorrAv in this frame (lmp2sqt.c:438) */
#define SQT_ERR_ARG (-1)
#define SQT_ERR_CUDA (-2)
#define SQT_ERR_NCCL (-3)
#define SQT_ERR_STATE (-4)
#define SQT_ERR_NOMEM (-5)

typedef struct sqt_handle sqt_handle;

/* The deck (lmp2sqt.c:82-92 DATALIST) and the getopt flags (lmp2sqt.c:152-191). */
typedef struct {
	double kVal, deltaT, mass, rVal;
	int limitCorrAv, nCBuffer, nCSpatial, nCTime, nCSkip;
	int flag_f, flag_F, flag_t;   /* -f: F(q,t) and rho(k); -F: also F_s / F_d; -t: MSD, MQD, MSDCM and (with -f) the van Hove table.
	                               * The engine only runs what the flags ask for (lmp2sqt.c:1246-1249, 1370, 1479) */
	int device;          /* CUDA device ordinal this handle owns (sqt_create); first device of sqt_multi_create */
	int n_gpus;          /* sqt_multi_create: GPUs device .. device + n_gpus - 1, particle-sharded in this process; 0 = 1 */
	int flag_V;          /* -V (flag_velocity, lmp2sqt.c:172): current correlators; frames then carry vx vy vz */
	int flag_M;          /* -M (flag_magnet, :177): dipole-weighted density kmu_q1 and its correlators; frames then carry mux muy muz */
	int flag_s;          /* -s (flag_s, :156-160; switches flag_t on like getopt does): per-axis MSD and the stress-like
	                      * squared change of sum_n v (x) r; frames then carry vx vy vz */
	int n_types;         /* per-type partials (not in the reference): > 1 = atoms carry a species 1 .. n_types
	                      * (the dump's `type` column, sqt_set_types) and F_ab(q,t), F_s,a(q,t), MSD_a, MQD_a and the van Hove table of
	                      * every species are accumulated beside the totals.  0 / 1 = off.  Not with -V / -M / -s */
	int reserved[2];     /* must be zero */
} sqt_params;

/* Largest nCSpatial the kernels cover (the reference accepts any; a deck beyond this is refused by sqt_create).  With
 * flag_V or flag_M the limit is 256 (the weighted densities come from the tensor-core rho(k) kernel). */
#define SQT_MAX_NCSPATIAL 1024
#define SQT_MAX_NCSPATIAL_V 256

/*
 * Raw running sums, exactly the arrays prePrintProcess (lmp2sqt.c:576-675) consumes, before
 * normalisation.  2-D arrays are [nCSpatial][nCTime] row-major like the reference's
 * avF_qq2[nk][nt].  The caller owns the arrays (sqt_results_alloc / sqt_results_free).
 */
typedef struct {
	int nCSpatial, nCTime;
	int countCorrAv;         /* completed buffers accumulated into the sums            */
	int limitCorrAv;         /* value in force when the sums were latched              */
	long long nPtls;         /* atoms per frame (all shards)                            */
	double g_Vol;            /* L^3 of the last processed frame  (lmp2sqt.c:1621)       */
	double kVal;             /* 2 pi / L of the last processed frame (lmp2sqt.c:1632)   */
	double first_Vol;        /* L^3 of the FIRST processed frame (factorDr, :1991-2004) */
	int first_limitCorrAv;   /* limitCorrAv at the first processed frame (:2004)        */
	int number_call_print;   /* Number_call_Print after this print (lmp2sqt.c:674)      */
	double *avF_qq2, *StdDevF_qq2, *avF_s_qq2, *StdDevF_s_qq2, *avF_d_qq2, *StdDevF_d_qq2;
	double *avDrTable;
	double *rrMSDAv, *rrMQDAv, *rrMSDCMAv;   /* [nCTime] */
	/* -V only (zero otherwise): avC2_v_rho, avC2_rho_v, avC2_v_v [nCSpatial][nCTime] (lmp2sqt.c:394-398), rrCvvAv,
	 * rrCvcmvcmAv [nCTime] (:423-426) */
	double *avC2_v_rho, *avC2_rho_v, *avC2_v_v;
	double *rrCvvAv, *rrCvcmvcmAv;
	/* -M only: avC2_mu_rho, avC2_rho_mu, avC2_mu_mu, and with -V as well avC2_v_mu, avC2_mu_v (lmp2sqt.c:399-407);
	 * rrCmmAv [nCTime] stays zero: the reference allocates, zeroes and prints it but never adds to it (:420-432) */
	double *avC2_v_mu, *avC2_mu_v, *avC2_mu_rho, *avC2_rho_mu, *avC2_mu_mu;
	double *rrCmmAv;
	/* -s only: rrMSR1_R_Av [nCTime][3] (x y z) and rrMSR2_VR_Av [nCTime][9] (xx xy xz yx yy yz zx zy zz), the reference's
	 * VecR3 / Rank2R3 arrays (:427-432) */
	double *rrMSR1_R_Av, *rrMSR2_VR_Av;
	/* per-type partials (sqt_results_alloc_types; NULL / 0 otherwise), raw sums like the arrays above:
	 *   avF_ab [a][b][nCSpatial][nCTime]   sum over buffers and axes of Re[rho_a(q,t) rho_b*(q,t0)]       (sum over a, b = avF_qq2)
	 *   avF_s_a [a][nCSpatial][nCTime]     the self part restricted to species a                          (sum over a = avF_s_qq2)
	 *   avDrTable_a [a][nCSpatial][nCTime] van Hove counts of species a                                   (sum over a = avDrTable)
	 *   rrMSD_a, rrMQD_a [a][nCTime]       sum over the atoms of species a                                (sum over a = rrMSDAv, rrMQDAv)
	 *   nPtls_a [a]                        atoms of species a over all shards */
	int n_types;             /* species the arrays below are allocated for (sqt_results_alloc_types); must equal the handle's */
	long long nPtls_a[8];
	double *avF_ab, *avF_s_a, *avDrTable_a, *rrMSD_a, *rrMQD_a;
} sqt_results;

int  sqt_results_alloc(sqt_results *r, int nCSpatial, int nCTime);
int  sqt_results_alloc_types(sqt_results *r, int n_types);   /* after sqt_results_alloc: the per-type arrays */
void sqt_results_free(sqt_results *r);

/* ---- lifecycle --------------------------------------------------------------------------- */
int  sqt_create(const sqt_params *p, sqt_handle **out);
void sqt_destroy(sqt_handle *h);
const char *sqt_last_error(const sqt_handle *h);   /* h may be NULL: error of the last failed sqt_create */

/* main's `limitCorrAv += floor((n_snap - nCTime)/(nCTime/nCBuffer))` (lmp2sqt.c:244) happens
 * between GetNameList and the first frame; the host tells the engine the final value here. */
int  sqt_set_limit(sqt_handle *h, int limitCorrAv);

/* InitSpacetimeCorr (lmp2sqt.c:446-475): count[nb] = -nb*nCTime/nCBuffer, nSkip = 0.  Once per
 * input file; the averages persist across files, partially filled buffers are discarded. */
int  sqt_begin_file(sqt_handle *h);

/* Per-type partials (sqt_params.n_types > 1): the species (1 .. n_types) of this handle's atoms in dump order, once, before
 * the first frame.  Atoms keep their line order from frame to frame (snapshot.c:357-368), so the species never change. */
int  sqt_set_types(sqt_handle *h, const int *types, int n_atoms_local);

/* ---- frames ------------------------------------------------------------------------------ */
/* Pinned host staging for the next frame, SoA: x[0..n), y[0..n), z[0..n) as doubles
 * (3*n_atoms_local contiguous; then vx vy vz with flag_V or flag_s, then mux muy muz with flag_M: sqt_frame_components arrays).  Double buffered: the pointer stays valid
 * until the second following sqt_stage_acquire.  Blocks only if the slot's previous H2D copy is still in flight. */
double *sqt_stage_acquire(sqt_handle *h, int n_atoms_local);

/* EvalSpacetimeCorr (lmp2sqt.c:1599-1650) for one frame.  `soa` is either a buffer from
 * sqt_stage_acquire (async H2D, no extra copy), or page-locked host memory of the caller
 * (cudaHostAlloc / cudaHostRegister: async H2D straight from it -- it must stay unmodified until
 * two further pushes or sqt_sync), or pageable host memory (copied synchronously into staging).  n_atoms_local = atoms this handle owns (its particle shard, dump order);
 * n_atoms_total = atoms in the frame over all shards.  xlo/xhi = x box bounds (L, kVal and
 * g_Vol derive from x only, lmp2sqt.c:1620-1632).
 * Returns SQT_OK, SQT_FLUSH, or <0. */
int  sqt_push_frame(sqt_handle *h, const double *soa, int n_atoms_local, long long n_atoms_total,
                    double xlo, double xhi);
/* Same with one pointer per component (this handle's slice of a caller-owned SoA frame; vx/vy/vz may be NULL unless
 * flag_V).  Pinned memory is copied asynchronously and must stay unmodified until two further pushes or sqt_sync. */
int  sqt_push_frame_xyz(sqt_handle *h, const double *x, const double *y, const double *z, const double *vx, const double *vy,
                        const double *vz, int n_atoms_local, long long n_atoms_total, double xlo, double xhi);
/* ... and with the dipole components (unit vectors as the reference's reader leaves them, snapshot.c:440-459; needed with flag_M) */
int  sqt_push_frame_xyz_vm(sqt_handle *h, const double *x, const double *y, const double *z, const double *vx, const double *vy,
                           const double *vz, const double *mux, const double *muy, const double *muz, int n_atoms_local,
                           long long n_atoms_total, double xlo, double xhi);
/* Arrays per frame the handle expects in SoA frames: 3, 6 or 9 (x y z [vx vy vz] [mux muy muz]). */
int  sqt_frame_components(const sqt_handle *h);
/* Same, for a frame already resident in device memory (SoA, 3*n_atoms_local doubles; 6*n with flag_V) on
 * the handle's device -- the binary / on-device frame source path.
 * Stream ordering: the engine copies d_soa on its own (blocking) stream.  A producer that fills d_soa on the legacy
 * default stream is ordered automatically; a producer on any other stream (cudaStreamNonBlocking, per-thread default
 * stream, a torch side stream) must either synchronise before the call or use the _ex variant. */
int  sqt_push_frame_device(sqt_handle *h, const double *d_soa, int n_atoms_local,
                           long long n_atoms_total, double xlo, double xhi);
/* producer_stream: the cudaStream_t (passed as void *) that wrote d_soa.  The engine's copy waits for the work enqueued on
 * it so far, and work enqueued on it after the call waits for the engine's copy, so the producer may refill d_soa. */
int  sqt_push_frame_device_ex(sqt_handle *h, const double *d_soa, int n_atoms_local,
                              long long n_atoms_total, double xlo, double xhi, void *producer_stream);

/* ---- results ----------------------------------------------------------------------------- */
/* After SQT_FLUSH: the sums latched at the flush (the averages were zeroed on the device
 * exactly like ZeroAvSpacetimeCorr, lmp2sqt.c:366).  Otherwise: the current running sums.
 * Synchronises.  On several shards every rank receives identical, fully reduced arrays. */
int  sqt_fetch_results(sqt_handle *h, sqt_results *out);
/* End of run (lmp2sqt.c:327-333): if nothing was printed yet, limitCorrAv = countCorrAv and the
 * sums are returned (returns 1 = "print now"); returns 0 if a mid-run print already happened. */
int  sqt_finish(sqt_handle *h, sqt_results *out);
int  sqt_sync(sqt_handle *h);

/* ---- introspection (bookkeeping parity tests, bench) ------------------------------------- */
int  sqt_count_corr_av(const sqt_handle *h);
int  sqt_buffer_count(const sqt_handle *h, int nb);       /* TBuf.count, lmp2sqt.h:69 */
int  sqt_number_call_print(const sqt_handle *h);
/* rho(k) of the last processed frame, [3][nCSpatial][2] (marker layout lmp2sqt.c:1173-1174), reduced over all shards.
 * Waits only for that frame's rho(k) kernel, its allreduce and the F(q,t) update -- not for the frame's self-term kernel,
 * which keeps running while the caller prepares the next frame.  The first call also tells a sharded engine that its caller
 * reads per-frame results: from then on every frame's self-term kernel is enqueued behind the frame's allreduce (which could
 * otherwise not get an SM beside the persistent kernel and would return only when that kernel ends). */
int  sqt_get_rho(sqt_handle *h, double *out);
/* With flag_V: the current density kvel_q1 of the last processed frame, same layout (lmp2sqt.c:1203-1208). */
int  sqt_get_kvel(sqt_handle *h, double *out);
/* With flag_M: kmu_q1 (lmp2sqt.c:1209-1214). */
int  sqt_get_kmu(sqt_handle *h, double *out);
long long sqt_kernel_launches(const sqt_handle *h);       /* launches of OUR kernels so far */
/* CUDA-event timing of a frame's dominant kernels on their own stream: the self-term kernel of the family in
 * use (self_ring_mma_kernel / self_ring_vec_kernel / self_ring_kernel) plus the rho(k) kernel, one event pair around
 * both per frame.  enable=1 starts recording; sqt_profile_read synchronises and returns the number of frames and
 * the summed milliseconds since enable, then clears. */
int  sqt_profile_enable(sqt_handle *h, int enable);
int  sqt_profile_read(sqt_handle *h, long long *n_launches, double *ms_total);
/* Device-side clock over a region of frames: start records an event on the engine's main
 * stream; stop records one after the last pushed frame's collective stage, waits for it and
 * returns the elapsed milliseconds (CUDA events, not host wall clock). */
int  sqt_timer_start(sqt_handle *h);
int  sqt_timer_stop(sqt_handle *h, double *ms);
/* Human-readable description of the kernel configuration picked for this deck. */
const char *sqt_kernel_config(const sqt_handle *h);

/* ---- multi-GPU: one process (handle) per GPU, particle-sharded --------------------------- */
/* Rank 0 obtains an opaque id (ncclUniqueId, 128 bytes) and ships it to the other ranks by
 * whatever plumbing the host has (torch.distributed broadcast, MPI, a file...). */
#define SQT_COMM_ID_BYTES 128
int  sqt_comm_unique_id(void *id_out);
int  sqt_comm_init(sqt_handle *h, const void *id, int rank, int nranks);

/* ---- multi-GPU inside one process: one reader, frame fan-out to G GPUs ------------------
 * The reference's frame loop (lmp2sqt.c:292-320) stays single threaded in the caller: it fills ONE pinned SoA frame
 * (sqt_multi_stage_acquire) and pushes it; the library owns G engine handles on devices p->device .. p->device +
 * p->n_gpus - 1, one worker thread and one NCCL rank per GPU, splits the frame into G contiguous particle chunks
 * (dump order) and issues one asynchronous H2D copy per component and GPU.  Results are identical on every rank;
 * they are fetched from rank 0.  Same call sequence and return codes as the single-handle API. */
typedef struct sqt_multi sqt_multi;
int  sqt_multi_create(const sqt_params *p, sqt_multi **out);
void sqt_multi_destroy(sqt_multi *m);
const char *sqt_multi_last_error(const sqt_multi *m);
int  sqt_multi_n_gpus(const sqt_multi *m);
sqt_handle *sqt_multi_handle(sqt_multi *m, int rank);      /* introspection only */
int  sqt_multi_set_limit(sqt_multi *m, int limitCorrAv);
int  sqt_multi_set_types(sqt_multi *m, const int *types, int n_atoms);   /* per-type partials: species of ALL atoms, dump order */
int  sqt_multi_begin_file(sqt_multi *m);
/* pinned (portable) SoA for a whole frame: x[n] y[n] z[n] (+ vx[n] vy[n] vz[n] with flag_V / flag_s, + mux[n] muy[n] muz[n] with
 * flag_M); double buffered */
double *sqt_multi_stage_acquire(sqt_multi *m, int n_atoms);
int  sqt_multi_push_frame(sqt_multi *m, const double *soa, int n_atoms, double xlo, double xhi);
int  sqt_multi_fetch_results(sqt_multi *m, sqt_results *out);
int  sqt_multi_finish(sqt_multi *m, sqt_results *out);
int  sqt_multi_sync(sqt_multi *m);
long long sqt_multi_kernel_launches(const sqt_multi *m);

/* ---- on-device synthetic frame source (bench workloads) ------------ */
typedef struct sqt_synth sqt_synth;
/* fcc lattice of 4*n_cells^3 sites at reduced density 0.8442, jitter 0.10, random walk
 * 0.077/axis/frame, counter-based RNG keyed (seed, atom, frame): shard independent.
 * The source owns atoms [first_atom, first_atom + n_atoms_local). */
int  sqt_synth_create(int device, int n_cells, long long first_atom, int n_atoms_local,
                      unsigned long long seed, sqt_synth **out);
/* Advances to the next frame (frame 0 on the first call); returns the device SoA pointer. */
const double *sqt_synth_next(sqt_synth *s);
double sqt_synth_box_length(const sqt_synth *s);
/* Copies the current frame into host memory (SoA, 3*n doubles). */
int  sqt_synth_copy_to_host(sqt_synth *s, double *dst);
/* Copies the current frame into another device buffer (SoA, 3*n doubles) on the same device. */
int  sqt_synth_copy_to_device(sqt_synth *s, double *d_dst);
void sqt_synth_destroy(sqt_synth *s);

/* ---- roofline denominator: measured non-tensor FP64 FMA throughput ----------------------- */
/* Runs a register-resident DFMA loop on every SM; returns TFLOP/s (FMA = 2 flops). */
int  sqt_measure_dfma_peak(int device, double *tflops, double *sm_clock_mhz_est);

#ifdef __cplusplus
}
#endif
#endif
