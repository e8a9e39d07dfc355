/*
 * sqt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the lmp2sqt space-time-correlation hot path
 * (reference lmp2sqt.c:1599-1650 EvalSpacetimeCorr and its callees), used as the
 * parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this library; the product path never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement
 * against the unmodified reference (oracle/_ref/liblmp2sqt_ref.so built from
 * /root/reference by oracle/Makefile) and against the committed golden files in
 * tests/golden/ that the reference binary itself produced.  The -V (velocity) correlators are restated
 * and pinned the same way ahead of the product code (tests/golden/velocity).
 */
#ifndef SQT_ORACLE_H
#define SQT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sqt_oracle sqt_oracle;

/* Mirrors the deck (lmp2sqt.c:82-92) and the getopt flags (lmp2sqt.c:152-191). */
typedef struct {
	double kVal, deltaT, mass, rVal;
	int limitCorrAv, nCBuffer, nCSpatial, nCTime, nCSkip;
	int flag_f, flag_F, flag_t;
	int flag_V; /* -V (flag_velocity, lmp2sqt.c:172): current correlators C2_v_v / C2_rho_v / C2_v_rho, Cvv, Cvcmvcm */
	int flag_M; /* -M (flag_magnet, :177): dipole-weighted density kmu_q1 and C2_mu_mu / C2_rho_mu / C2_mu_rho (+ C2_v_mu / C2_mu_v with -V) */
	int flag_s; /* -s (flag_s, :156; the CLI also sets flag_t): per-axis MSD rrMSR1_R and the stress-like rrMSR2_VR */
} sqt_oracle_params;

sqt_oracle *sqt_oracle_create(const sqt_oracle_params *p);
void sqt_oracle_destroy(sqt_oracle *o);

/* InitSpacetimeCorr (lmp2sqt.c:446-475): ring phase reset, once per input file. */
void sqt_oracle_begin_file(sqt_oracle *o);

/* EvalSpacetimeCorr (lmp2sqt.c:1599-1650) on one frame.  xyz is AoS [n][3].
 * Returns 1 when AccumSpacetimeCorr hit countCorrAv == limitCorrAv in this frame
 * (the reference calls PrintProcess there, lmp2sqt.c:438-440): the normalised
 * results were stored in the "printed" arrays and the averages were zeroed;
 * 0 otherwise; <0 on error. */
int sqt_oracle_push_frame(sqt_oracle *o, const double *xyz, int n, double xlo, double xhi);
/* The same with velocities, AoS [n][3] (required when flag_V is set). */
int sqt_oracle_push_frame_v(sqt_oracle *o, const double *xyz, const double *vel, int n, double xlo, double xhi);

/* The same with the dipole components mu, AoS [n][3] (required when flag_M is set; vel is required for flag_V / flag_s). */
int sqt_oracle_push_frame_vm(sqt_oracle *o, const double *xyz, const double *vel, const double *mu, int n, double xlo, double xhi);

/* End-of-run print (lmp2sqt.c:327-333): only if nothing was printed before,
 * limitCorrAv = countCorrAv, then PrintProcess.  Returns 1 if it printed now. */
int sqt_oracle_finish(sqt_oracle *o);

void sqt_oracle_set_limit(sqt_oracle *o, int limitCorrAv);

/* Scalars. */
int sqt_oracle_countCorrAv(const sqt_oracle *o);       /* live counter (reset by a print) */
int sqt_oracle_printed_count(const sqt_oracle *o);     /* countCorrAv at the last print */
int sqt_oracle_number_call_print(const sqt_oracle *o);
int sqt_oracle_limitCorrAv(const sqt_oracle *o);
double sqt_oracle_kVal(const sqt_oracle *o);           /* last 2*pi/L (lmp2sqt.c:1632) */
int sqt_oracle_nPtls(const sqt_oracle *o);
int sqt_oracle_buffer_count(const sqt_oracle *o, int nb);

/* Arrays.  "raw" = running sums before normalisation; "printed" = what
 * prePrintProcess (lmp2sqt.c:576-675) left in the arrays at the last print.
 * 2-D arrays are [nCSpatial][nCTime] row-major (the reference's avF_qq2[nk][nt]).
 * which: 0 avF, 1 StdDevF, 2 avF_s, 3 StdDevF_s, 4 avF_d, 5 StdDevF_d, 6 avDrTable,
 *        7 avC2_v_rho, 8 avC2_rho_v, 9 avC2_v_v (flag_V), 10 avC2_v_mu, 11 avC2_mu_v (flag_V && flag_M),
 *        12 avC2_mu_rho, 13 avC2_rho_mu, 14 avC2_mu_mu (flag_M) */
const double *sqt_oracle_raw2d(const sqt_oracle *o, int which);
const double *sqt_oracle_printed2d(const sqt_oracle *o, int which);
/* which: 0 rrMSDAv, 1 rrMQDAv, 2 rrMSDCMAv, 3 rrCvvAv, 4 rrCvcmvcmAv (flag_V), 5 rrCmmAv (flag_M; the reference never adds to it),
 *        6..8 rrMSR1_R_Av x y z, 9..17 rrMSR2_VR_Av xx xy xz yx yy yz zx zy zz, 18 rrMSR2_VR_Av_dig, 19 _offdig (flag_s; the last two
 *        exist only in the printed set)  (length nCTime) */
const double *sqt_oracle_raw1d(const sqt_oracle *o, int which);
const double *sqt_oracle_printed1d(const sqt_oracle *o, int which);
/* the raw sums as they stood when the last print normalised them (input of prePrintProcess) */
const double *sqt_oracle_printed_raw2d(const sqt_oracle *o, int which);
const double *sqt_oracle_printed_raw1d(const sqt_oracle *o, int which);
double sqt_oracle_first_vol(const sqt_oracle *o);   /* g_Vol at the first processed frame (factorDr input) */
int sqt_oracle_first_limit(const sqt_oracle *o);    /* limitCorrAv at the first processed frame */
double sqt_oracle_vol(const sqt_oracle *o);
const double *sqt_oracle_factorDr(const sqt_oracle *o);   /* lmp2sqt.c:1987-2009 */
/* current-frame rho(k): [3][nCSpatial][2]  (lmp2sqt.c:1173-1174 marker layout) */
const double *sqt_oracle_rho(const sqt_oracle *o);

#ifdef __cplusplus
}
#endif
#endif
