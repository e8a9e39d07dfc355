/*
 * sqt_oracle.c -- TEST INFRASTRUCTURE ONLY (see sqt_oracle.h).
 *
 * CPU restatement of the reference hot path in the reference's own *algorithmic form*
 * (per-particle cos/sin tables for the current frame and for every live time origin),
 * written so that every floating-point operation happens in the same order as in
 * /root/reference/lmp2sqt.c.  Built with -ffp-contract=off so gcc emits the same
 * un-fused x86-64 arithmetic as the reference build (Makefile:10,24) and F, F_s, F_d,
 * MSD, MQD and the van Hove counts come out bit-identical to the reference run with
 * OMP_NUM_THREADS=1.  (The reference's OpenMP loop at lmp2sqt.c:1377 races on the MSD
 * sums; this restatement is the serial semantics.)
 *
 * It is O((nCBuffer+1) * N * 6*nCSpatial) in memory like the reference, so it is only
 * used at sizes that finish in seconds.
 */
#include "sqt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define NAX 3 /* N_AXIS, lmp2sqt.h:103 */
#define N2D 15
#define N1D 20

typedef struct {
	/* per-origin copies, SetWaitedTimeCorr lmp2sqt.c:1260-1314 */
	double *org_pos;   /* [N][3]                 orgR            */
	double *org_rho;   /* [6 nCS]                org_rho_q1      */
	double *org_self;  /* [N][6 nCS]             org_rho_s_q1    */
	double *org_dist;  /* [N][6 nCS]             org_rho_d_q1    */
	/* per-buffer correlators, TBuf lmp2sqt.h:49-71 */
	double *F, *Fs, *Fd; /* [nCS][T] */
	int *hist;           /* [nCS][T]  DrTable */
	double *msd, *mqd, *msdcm; /* [T] */
	/* -V (flag_velocity): origin velocities and current density, their correlators (lmp2sqt.h:49-71) */
	double *org_vel;   /* [N][3]      orgV          */
	double *org_kvel;  /* [6 nCS]     org_kvel_q1   */
	double *Cvrho, *Crhov, *Cvv2; /* [nCS][T]  C2_v_rho, C2_rho_v, C2_v_v */
	double *cvv, *cvcm;           /* [T]       rrCvv, rrCvcmvcm */
	/* -M (flag_magnet): dipole-weighted density and its correlators */
	double *org_kmu;   /* [6 nCS]     org_kmu_q1    */
	double *Cvmu, *Cmuv, *Cmurho, *Crhomu, *Cmumu; /* [nCS][T]  C2_v_mu, C2_mu_v, C2_mu_rho, C2_rho_mu, C2_mu_mu */
	double *cmm;       /* [T] rrCmm (computed per buffer, never accumulated by AccumSpacetimeCorr: lmp2sqt.c:420-432) */
	double *org_mu;    /* [N][3]      orgMu         */
	/* -s (flag_s): per-axis MSD and the squared change of sum_n v (x) r */
	double *msr1;      /* [T][3]  rrMSR1_R  */
	double *msr2;      /* [T][9]  rrMSR2_VR */
	double org_sumVR[9];
	int count;
} ring_buf;

struct sqt_oracle {
	sqt_oracle_params p;
	int N;            /* nPtls, fixed at the first processed frame */
	int alloc_more;   /* flag_alloc_more */
	int nSkip;
	int countCorrAv, limitCorrAv, n_print, printed_count;
	double kVal, L, vol;
	ring_buf *buf;
	/* current frame (tBuf[0] scratch in the reference, lmp2sqt.c:1098-1102) */
	double *rho, *tmp_row, *self, *dist;
	double *kvel;      /* [6 nCS] kvel_q1: sum_n (q v_a) (-sin, cos)(q r_a), lmp2sqt.c:1203-1208 */
	double *kmu;       /* [6 nCS] kmu_q1, the same with the dipole components, :1209-1214 */
	double sumVR[9];   /* sumVR_ct: sum_n v (x) r, EvalOneTimeSumVR :1122-1142; zeroed only by ZeroOneTimeCorr, i.e. under -f (:1120) */
	const double *cur_xyz, *cur_vel, *cur_mu;
	/* running sums: 0 avF 1 StdDevF 2 avF_s 3 StdDevF_s 4 avF_d 5 StdDevF_d 6 avDrTable 7 avC2_v_rho 8 avC2_rho_v 9 avC2_v_v
	 * 10 avC2_v_mu 11 avC2_mu_v 12 avC2_mu_rho 13 avC2_rho_mu 14 avC2_mu_mu */
	double *sum2d[N2D], *out2d[N2D], *rawp2d[N2D]; /* running sums, normalised at last print, raw at last print */
	double *sum1d[N1D], *out1d[N1D], *rawp1d[N1D]; /* 0 MSD 1 MQD 2 MSDCM 3 Cvv 4 Cvcmvcm 5 Cmm 6-8 MSR1_R x y z 9-17 MSR2_VR xx..zz 18 dig 19 offdig */
	double first_vol; int first_limit;
	double *factorDr;
};

static double *zalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

sqt_oracle *sqt_oracle_create(const sqt_oracle_params *p)
{
	sqt_oracle *o = (sqt_oracle *)calloc(1, sizeof(*o));
	o->p = *p;
	/* GetNameList tail, lmp2sqt.c:2105, and main :240-242 */
	if (o->p.nCBuffer > o->p.nCTime) o->p.nCBuffer = o->p.nCTime;
	if (o->p.nCBuffer < 1) o->p.nCBuffer = 1;
	o->limitCorrAv = p->limitCorrAv;
	o->kVal = p->kVal;
	const int S = o->p.nCSpatial, T = o->p.nCTime, B = o->p.nCBuffer;
	o->buf = (ring_buf *)calloc(B, sizeof(ring_buf));
	for (int nb = 0; nb < B; nb++) {
		ring_buf *b = &o->buf[nb];
		b->F = zalloc((size_t)S * T); b->Fs = zalloc((size_t)S * T); b->Fd = zalloc((size_t)S * T);
		b->hist = (int *)calloc((size_t)S * T, sizeof(int));
		b->msd = zalloc(T); b->mqd = zalloc(T); b->msdcm = zalloc(T);
		b->org_rho = zalloc(2 * NAX * S);
		b->org_kvel = zalloc(2 * NAX * S);
		b->Cvrho = zalloc((size_t)S * T); b->Crhov = zalloc((size_t)S * T); b->Cvv2 = zalloc((size_t)S * T);
		b->cvv = zalloc(T); b->cvcm = zalloc(T);
		b->org_kmu = zalloc(2 * NAX * S);
		b->Cvmu = zalloc((size_t)S * T); b->Cmuv = zalloc((size_t)S * T); b->Cmurho = zalloc((size_t)S * T);
		b->Crhomu = zalloc((size_t)S * T); b->Cmumu = zalloc((size_t)S * T);
		b->cmm = zalloc(T); b->msr1 = zalloc((size_t)T * 3); b->msr2 = zalloc((size_t)T * 9);
	}
	for (int i = 0; i < N2D; i++) { o->sum2d[i] = zalloc((size_t)S * T); o->out2d[i] = zalloc((size_t)S * T); o->rawp2d[i] = zalloc((size_t)S * T); }
	for (int i = 0; i < N1D; i++) { o->sum1d[i] = zalloc(T); o->out1d[i] = zalloc(T); o->rawp1d[i] = zalloc(T); }
	o->rho = zalloc(2 * NAX * S);
	o->kvel = zalloc(2 * NAX * S);
	o->kmu = zalloc(2 * NAX * S);
	o->tmp_row = zalloc(2 * NAX * S);
	o->factorDr = zalloc(S);
	return o;
}

void sqt_oracle_destroy(sqt_oracle *o)
{
	if (!o) return;
	for (int nb = 0; nb < o->p.nCBuffer; nb++) {
		ring_buf *b = &o->buf[nb];
		free(b->F); free(b->Fs); free(b->Fd); free(b->hist); free(b->msd); free(b->mqd);
		free(b->msdcm); free(b->org_rho); free(b->org_pos); free(b->org_self); free(b->org_dist);
		free(b->org_vel); free(b->org_kvel); free(b->Cvrho); free(b->Crhov); free(b->Cvv2); free(b->cvv); free(b->cvcm);
		free(b->org_kmu); free(b->Cvmu); free(b->Cmuv); free(b->Cmurho); free(b->Crhomu); free(b->Cmumu); free(b->cmm); free(b->org_mu);
		free(b->msr1); free(b->msr2);
	}
	free(o->buf);
	for (int i = 0; i < N2D; i++) { free(o->sum2d[i]); free(o->out2d[i]); free(o->rawp2d[i]); }
	for (int i = 0; i < N1D; i++) { free(o->sum1d[i]); free(o->out1d[i]); free(o->rawp1d[i]); }
	free(o->rho); free(o->kvel); free(o->kmu); free(o->tmp_row); free(o->self); free(o->dist); free(o->factorDr);
	free(o);
}

/* InitSpacetimeCorr, lmp2sqt.c:446-475.  C integer arithmetic: (-nb * T) / B truncates
 * toward zero. */
void sqt_oracle_begin_file(sqt_oracle *o)
{
	o->nSkip = 0;
	for (int nb = 0; nb < o->p.nCBuffer; nb++)
		o->buf[nb].count = -nb * o->p.nCTime / o->p.nCBuffer;
}

void sqt_oracle_set_limit(sqt_oracle *o, int limitCorrAv) { o->limitCorrAv = limitCorrAv; }

/* Alloc_more, lmp2sqt.c:1926-2012: per-particle storage plus the van Hove
 * normalisation, which bakes in the limitCorrAv current at the first processed frame. */
static void alloc_more(sqt_oracle *o)
{
	const int S = o->p.nCSpatial, B = o->p.nCBuffer, N = o->N;
	const size_t row = 2 * NAX * (size_t)S;
	o->self = zalloc(N * row);
	o->dist = zalloc(N * row);
	for (int nb = 0; nb < B; nb++) {
		o->buf[nb].org_pos = zalloc((size_t)N * 3);
		o->buf[nb].org_vel = zalloc((size_t)N * 3);
		o->buf[nb].org_mu = zalloc((size_t)N * 3);
		o->buf[nb].org_self = zalloc(N * row);
		o->buf[nb].org_dist = zalloc(N * row);
	}
	const double rVal = o->p.rVal;
	double rho0 = o->N / o->vol, shell_Vol;
	for (int nr = 0; nr < S; nr++) {
		if (nr == 0) shell_Vol = 4 * M_PI / 3. * pow(rVal, 3);
		else shell_Vol = (4. / 3.) * M_PI * (pow((nr + 1) * rVal, 3) - pow(nr * rVal, 3));
		o->factorDr[nr] = 1. / (pow(rho0, 2) * o->vol * shell_Vol * o->limitCorrAv);
	}
}

/* ZeroOneTimeCorr + EvalOneTimeKspace, lmp2sqt.c:1095-1121, 1144-1235. */
static void one_time_kspace(sqt_oracle *o)
{
	const int S = o->p.nCSpatial, N = o->N;
	const int row = 2 * NAX * S;
	const double kVal = o->kVal;
	for (int j = 0; j < row; j++) { o->rho[j] = 0.; if (o->p.flag_V) o->kvel[j] = 0.; if (o->p.flag_M) o->kmu[j] = 0.; }
	for (int j = 0; j < 9; j++) o->sumVR[j] = 0.; /* :1120 */
	if (o->p.flag_F) {
		memset(o->self, 0, sizeof(double) * (size_t)N * row);
		memset(o->dist, 0, sizeof(double) * (size_t)N * row);
	}
	for (int n = 0; n < N; n++) {
		const double *r = o->cur_xyz + 3 * (size_t)n;
		for (int k = 0; k < NAX; k++) {
			double b, c = 0, s = 0, c0 = 0, c1 = 0, s1 = 0, c2, s2;
			for (int m = 0; m < S; m++) {
				const int mR = (S * 2) * k + 2 * m, mI = mR + 1;
				if (m == 0) {
					b = kVal * r[k];
					c = cos(b);
					s = sin(b);
					c0 = c;
				} else if (m == 1) {
					c1 = c; s1 = s;
					c = 2. * c0 * c1 - 1.;
					s = 2. * c0 * s1;
				} else {
					c2 = c1; s2 = s1;
					c1 = c; s1 = s;
					c = 2. * c0 * c1 - c2;
					s = 2. * c0 * s1 - s2;
				}
				if (o->p.flag_V) { /* :1203-1208 */
					const int qn = m + 1;
					const double kv = qn * kVal * o->cur_vel[3 * (size_t)n + k];
					o->kvel[mR] += -kv * s;
					o->kvel[mI] += +kv * c;
				}
				if (o->p.flag_M) { /* :1209-1214 */
					const int qn = m + 1;
					const double km = qn * kVal * o->cur_mu[3 * (size_t)n + k];
					o->kmu[mR] += -km * s;
					o->kmu[mI] += +km * c;
				}
				o->tmp_row[mR] = c;
				o->tmp_row[mI] = s;
			}
		}
		for (int j = 0; j < row; j++) o->rho[j] += o->tmp_row[j];
		if (o->p.flag_F) memcpy(o->self + (size_t)n * row, o->tmp_row, sizeof(double) * row);
	}
	if (o->p.flag_F)
		for (int j = 0; j < row; j++)
			for (int n = 0; n < N; n++)
				o->dist[(size_t)n * row + j] = o->rho[j] - o->self[(size_t)n * row + j];
}

/* SetWaitedTimeCorr, lmp2sqt.c:1260-1314. */
static void capture_origin(sqt_oracle *o, ring_buf *b)
{
	const int S = o->p.nCSpatial, N = o->N;
	const size_t row = 2 * NAX * (size_t)S;
	if (o->p.flag_s) memcpy(b->org_sumVR, o->sumVR, sizeof(o->sumVR)); /* :1273-1275 */
	if (o->p.flag_t) memcpy(b->org_pos, o->cur_xyz, sizeof(double) * 3 * (size_t)N);
	if (o->p.flag_t && o->p.flag_M) memcpy(b->org_mu, o->cur_mu, sizeof(double) * 3 * (size_t)N);
	if (o->p.flag_t && o->p.flag_V) memcpy(b->org_vel, o->cur_vel, sizeof(double) * 3 * (size_t)N);
	if (o->p.flag_f) {
		memcpy(b->org_rho, o->rho, sizeof(double) * row);
		if (o->p.flag_V) memcpy(b->org_kvel, o->kvel, sizeof(double) * row);
		if (o->p.flag_M) memcpy(b->org_kmu, o->kmu, sizeof(double) * row);
		if (o->p.flag_F) {
			memcpy(b->org_self, o->self, sizeof(double) * N * row);
			memcpy(b->org_dist, o->dist, sizeof(double) * N * row);
		}
	}
}

/* InitTwoTimeCorr + EvalTwoTimeEach + EvalTwoTimeKSpace, lmp2sqt.c:1315-1591. */
static void two_time(sqt_oracle *o, ring_buf *b, int lag)
{
	const int S = o->p.nCSpatial, T = o->p.nCTime, N = o->N;
	const int row = 2 * NAX * S;
	/* InitTwoTimeCorr :1315-1367 (rrMSDCM is NOT zeroed: cumulative across buffer reuse) */
	if (o->p.flag_t) {
		b->msd[lag] = 0.; b->mqd[lag] = 0.;
		if (o->p.flag_V) { b->cvv[lag] = 0.; b->cvcm[lag] = 0.; }
		if (o->p.flag_M) b->cmm[lag] = 0.;
		if (o->p.flag_s) { for (int j = 0; j < 3; j++) b->msr1[3 * lag + j] = 0.; for (int j = 0; j < 9; j++) b->msr2[9 * lag + j] = 0.; }
	}
	if (o->p.flag_f) {
		for (int nr = 0; nr < S; nr++) b->hist[nr * T + lag] = 0;
		for (int k = 0; k < S; k++) {
			b->F[k * T + lag] = 0.;
			if (o->p.flag_V) { b->Cvrho[k * T + lag] = 0.; b->Crhov[k * T + lag] = 0.; b->Cvv2[k * T + lag] = 0.; }
			if (o->p.flag_V && o->p.flag_M) { b->Cvmu[k * T + lag] = 0.; b->Cmuv[k * T + lag] = 0.; }
			if (o->p.flag_M) { b->Cmurho[k * T + lag] = 0.; b->Crhomu[k * T + lag] = 0.; b->Cmumu[k * T + lag] = 0.; }
			if (o->p.flag_F) { b->Fs[k * T + lag] = 0.; b->Fd[k * T + lag] = 0.; }
		}
	}
	/* EvalTwoTimeEach :1368-1460 (serial semantics of the racy omp loop) */
	if (o->p.flag_t) {
		double six = 0, siy = 0, siz = 0, sjx = 0, sjy = 0, sjz = 0;
		double svix = 0, sviy = 0, sviz = 0, svjx = 0, svjy = 0, svjz = 0;
		for (int n = 0; n < N; n++) {
			const double *ri = o->cur_xyz + 3 * (size_t)n;
			const double *rj = b->org_pos + 3 * (size_t)n;
			six += ri[0]; siy += ri[1]; siz += ri[2];
			sjx += rj[0]; sjy += rj[1]; sjz += rj[2];
			double Cvv = 0;
			if (o->p.flag_V) { /* :1397-1415 */
				const double *vi = o->cur_vel + 3 * (size_t)n, *vj = b->org_vel + 3 * (size_t)n;
				svix += vi[0]; sviy += vi[1]; sviz += vi[2];
				svjx += vj[0]; svjy += vj[1]; svjz += vj[2];
				Cvv = vi[0] * vj[0];
				Cvv += vi[1] * vj[1];
				Cvv += vi[2] * vj[2];
			}
			double Cmm = 0;
			if (o->p.flag_M) { /* :1417-1421 */
				const double *mi = o->cur_mu + 3 * (size_t)n, *mj = b->org_mu + 3 * (size_t)n;
				Cmm = mi[0] * mj[0];
				Cmm += mi[1] * mj[1];
				Cmm += mi[2] * mj[2];
			}
			double dx = ri[0] - rj[0], dy = ri[1] - rj[1], dz = ri[2] - rj[2];
			double dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
			double dr2 = dx2 + dy2 + dz2;
			if (o->p.flag_f) {
				int i_Dr = floor(sqrt(dr2) / o->p.rVal);
				if (i_Dr < S) b->hist[i_Dr * T + lag]++;
			}
			b->msd[lag] += dr2;
			b->msr1[3 * lag + 0] += dx2; b->msr1[3 * lag + 1] += dy2; b->msr1[3 * lag + 2] += dz2; /* :1434-1436, whatever the flags */
			b->mqd[lag] += dr2 * dr2;
			if (o->p.flag_V) b->cvv[lag] += Cvv;
			if (o->p.flag_M) b->cmm[lag] += Cmm;
		}
		double dcm = (six - sjx) * (six - sjx);
		dcm += (siy - sjy) * (siy - sjy);
		dcm += (siz - sjz) * (siz - sjz);
		b->msdcm[lag] += dcm / (N * N); /* int product, as in the reference :1450 */
		if (o->p.flag_V) { /* :1452-1457 */
			double c = svix * svjx;
			c += sviy * svjy;
			c += sviz * svjz;
			b->cvcm[lag] += c / (N * N);
		}
	}
	/* EvalTwoTimeCollective :1461-1476 */
	if (o->p.flag_s)
		for (int j = 0; j < 9; j++) {
			const double d = o->sumVR[j] - b->org_sumVR[j];
			b->msr2[9 * lag + j] += d * d;
		}
	/* EvalTwoTimeKSpace :1477-1591, w = 1 */
	if (o->p.flag_f) {
		const double w = 1.;
		for (int a = 0; a < NAX; a++) {
			for (int nk = 0; nk < S; nk++) {
				const int mR = (S * 2) * a + 2 * nk, mI = mR + 1;
				if (o->p.flag_F) {
					for (int n = 0; n < N; n++) {
						const double *s = o->self + (size_t)n * row, *d = o->dist + (size_t)n * row;
						const double *s0 = b->org_self + (size_t)n * row, *d0 = b->org_dist + (size_t)n * row;
						b->Fs[nk * T + lag] += w * (s[mR] * s0[mR] + s[mI] * s0[mI]);
						b->Fd[nk * T + lag] += w * (d[mR] * s0[mR] + d[mI] * s0[mI]) +
						                       w * (s[mR] * d0[mR] + s[mI] * d0[mI]);
					}
				}
				b->F[nk * T + lag] += w * (o->rho[mR] * b->org_rho[mR] + o->rho[mI] * b->org_rho[mI]);
				if (o->p.flag_V) { /* :1542-1555 */
					b->Cvv2[nk * T + lag] += w * (o->kvel[mR] * b->org_kvel[mR] + o->kvel[mI] * b->org_kvel[mI]);
					b->Crhov[nk * T + lag] += w * (o->kvel[mR] * b->org_rho[mR] + o->kvel[mI] * b->org_rho[mI]);
					b->Cvrho[nk * T + lag] += w * (o->rho[mR] * b->org_kvel[mR] + o->rho[mI] * b->org_kvel[mI]);
				}
				if (o->p.flag_V && o->p.flag_M) { /* :1556-1564 */
					b->Cmuv[nk * T + lag] += w * (o->kvel[mR] * b->org_kmu[mR] + o->kvel[mI] * b->org_kmu[mI]);
					b->Cvmu[nk * T + lag] += w * (o->kmu[mR] * b->org_kvel[mR] + o->kmu[mI] * b->org_kvel[mI]);
				}
				if (o->p.flag_M) { /* :1565-1576 */
					b->Crhomu[nk * T + lag] += w * (o->kmu[mR] * b->org_rho[mR] + o->kmu[mI] * b->org_rho[mI]);
					b->Cmurho[nk * T + lag] += w * (o->rho[mR] * b->org_kmu[mR] + o->rho[mI] * b->org_kmu[mI]);
					b->Cmumu[nk * T + lag] += w * (o->kmu[mR] * b->org_kmu[mR] + o->kmu[mI] * b->org_kmu[mI]);
				}
			}
		}
	}
}

/* prePrintProcess + ZeroAvSpacetimeCorr, lmp2sqt.c:576-675, 477-540 (file writing is the
 * product's host code, not the oracle's job: the normalised arrays are kept instead). */
static void print_process(sqt_oracle *o)
{
	const int S = o->p.nCSpatial, T = o->p.nCTime, N = o->N;
	double inverseNptls = 1. / N;
	double inverseNptlsSq = 1. / (N * N);
	double average_number = 1. / (3.0 * o->countCorrAv);
	for (int i = 0; i < N2D; i++) { memcpy(o->out2d[i], o->sum2d[i], sizeof(double) * S * T); memcpy(o->rawp2d[i], o->sum2d[i], sizeof(double) * S * T); }
	for (int i = 0; i < N1D; i++) { memcpy(o->out1d[i], o->sum1d[i], sizeof(double) * T); memcpy(o->rawp1d[i], o->sum1d[i], sizeof(double) * T); }
	if (o->p.flag_f) {
		for (int nr = 0; nr < S; nr++)
			for (int nt = 0; nt < T; nt++) {
				const int e = nr * T + nt;
				for (int q = 0; q < (o->p.flag_F ? 3 : 1); q++) {
					double average = (o->out2d[2 * q][e] * inverseNptls) * average_number;
					double averageSq = (o->out2d[2 * q + 1][e] * inverseNptlsSq) * average_number;
					o->out2d[2 * q][e] = average;
					o->out2d[2 * q + 1][e] = sqrt((averageSq - average * average));
				}
				if (o->p.flag_V) { /* :609-613 */
					o->out2d[7][e] *= inverseNptls * average_number;
					o->out2d[8][e] *= inverseNptls * average_number;
					o->out2d[9][e] *= inverseNptls * average_number;
				}
				if (o->p.flag_V && o->p.flag_M) { o->out2d[11][e] *= inverseNptls * average_number; o->out2d[10][e] *= inverseNptls * average_number; } /* :614-617 */
				if (o->p.flag_M) for (int q = 12; q < 15; q++) o->out2d[q][e] *= inverseNptls * average_number; /* :618-622 */
			}
		for (int nt = 1; nt < T; nt++)
			for (int nr = 0; nr < S; nr++) o->out2d[6][nr * T + nt] *= o->factorDr[nr];
	}
	double scale_factor = 1. / (N * o->countCorrAv);
	double factor_msdcm = 1. / (o->countCorrAv);
	double factor_Cvv = 1. / (N * o->countCorrAv * 3.);   /* :640-642 */
	double factor_Cvcmvcm = 1. / (o->countCorrAv * 3.);
	double scale_4stress = .5 * o->p.mass * o->p.mass / (o->countCorrAv * o->vol); /* :637 */
	if (o->p.flag_t)
		for (int nt = 0; nt < T; nt++) {
			o->out1d[0][nt] *= scale_factor;
			o->out1d[2][nt] *= factor_msdcm;
			o->out1d[1][nt] *= scale_factor;
			if (o->p.flag_V) { o->out1d[3][nt] *= factor_Cvv; o->out1d[4][nt] *= factor_Cvcmvcm; }
			if (o->p.flag_s) { /* :660-670 */
				for (int j = 0; j < 3; j++) o->out1d[6 + j][nt] = scale_factor * o->out1d[6 + j][nt];
				for (int j = 0; j < 9; j++) o->out1d[9 + j][nt] = scale_4stress * o->out1d[9 + j][nt];
				double r2[9]; /* xx xy xz yx yy yz zx zy zz; common.c:161-172 */
				for (int j = 0; j < 9; j++) r2[j] = o->out1d[9 + j][nt];
				o->out1d[18][nt] = (r2[0] + r2[4] + r2[8]) / 3.0;
				o->out1d[19][nt] = (r2[1] + r2[2] + r2[3] + r2[5] + r2[6] + r2[7]) / 6.0;
			}
		}
	o->n_print++;
	o->printed_count = o->countCorrAv;
	/* ZeroAvSpacetimeCorr :477-540 */
	o->countCorrAv = 0;
	for (int i = 0; i < N2D; i++) memset(o->sum2d[i], 0, sizeof(double) * S * T);
	for (int i = 0; i < N1D; i++) memset(o->sum1d[i], 0, sizeof(double) * T);
}

/* AccumSpacetimeCorr, lmp2sqt.c:368-444. */
static int accumulate(sqt_oracle *o)
{
	const int S = o->p.nCSpatial, T = o->p.nCTime;
	int printed = 0;
	for (int nb = 0; nb < o->p.nCBuffer; nb++) {
		ring_buf *b = &o->buf[nb];
		if (b->count != T) continue;
		if (o->p.flag_f) {
			for (int k = 0; k < S; k++)
				for (int n = 0; n < T; n++) {
					const int e = k * T + n;
					o->sum2d[0][e] += b->F[e];
					o->sum2d[1][e] += b->F[e] * b->F[e];
					if (o->p.flag_F) {
						o->sum2d[2][e] += b->Fs[e];
						o->sum2d[3][e] += b->Fs[e] * b->Fs[e];
						o->sum2d[4][e] += b->Fd[e];
						o->sum2d[5][e] += b->Fd[e] * b->Fd[e];
					}
					if (o->p.flag_V) { /* :394-398 */
						o->sum2d[7][e] += b->Cvrho[e];
						o->sum2d[8][e] += b->Crhov[e];
						o->sum2d[9][e] += b->Cvv2[e];
					}
					if (o->p.flag_V && o->p.flag_M) { o->sum2d[10][e] += b->Cvmu[e]; o->sum2d[11][e] += b->Cmuv[e]; } /* :399-402 */
					if (o->p.flag_M) { o->sum2d[12][e] += b->Cmurho[e]; o->sum2d[13][e] += b->Crhomu[e]; o->sum2d[14][e] += b->Cmumu[e]; } /* :403-407 */
				}
			for (int nt = 0; nt < T; nt++)
				for (int nr = 0; nr < S; nr++) o->sum2d[6][nr * T + nt] += b->hist[nr * T + nt];
		}
		if (o->p.flag_t)
			for (int nt = 0; nt < T; nt++) {
				o->sum1d[0][nt] += b->msd[nt];
				o->sum1d[2][nt] += b->msdcm[nt];
				o->sum1d[1][nt] += b->mqd[nt];
				if (o->p.flag_V) { o->sum1d[3][nt] += b->cvv[nt]; o->sum1d[4][nt] += b->cvcm[nt]; } /* :423-426 */
				/* rrCmm is NOT accumulated (no rrCmmAv line in :420-432): sum1d[5] stays 0 */
				if (o->p.flag_s) { /* :427-432 */
					for (int j = 0; j < 3; j++) o->sum1d[6 + j][nt] += b->msr1[3 * nt + j];
					for (int j = 0; j < 9; j++) o->sum1d[9 + j][nt] += b->msr2[9 * nt + j];
				}
			}
		b->count = 0;
		++o->countCorrAv;
		if (o->countCorrAv == o->limitCorrAv) { print_process(o); printed = 1; }
	}
	return printed;
}

/* EvalSpacetimeCorr, lmp2sqt.c:1599-1650. */
int sqt_oracle_push_frame(sqt_oracle *o, const double *xyz, int n, double xlo, double xhi)
{
	return sqt_oracle_push_frame_v(o, xyz, NULL, n, xlo, xhi);
}

int sqt_oracle_push_frame_v(sqt_oracle *o, const double *xyz, const double *vel, int n, double xlo, double xhi)
{
	return sqt_oracle_push_frame_vm(o, xyz, vel, NULL, n, xlo, xhi);
}

/* EvalOneTimeSumVR, lmp2sqt.c:1122-1142 */
static void one_time_sum_vr(sqt_oracle *o)
{
	for (int n = 0; n < o->N; n++) {
		const double *r = o->cur_xyz + 3 * (size_t)n, *v = o->cur_vel + 3 * (size_t)n;
		for (int a = 0; a < 3; a++)
			for (int b = 0; b < 3; b++) o->sumVR[3 * a + b] = o->sumVR[3 * a + b] + v[a] * r[b];
	}
}

int sqt_oracle_push_frame_vm(sqt_oracle *o, const double *xyz, const double *vel, const double *mu, int n, double xlo, double xhi)
{
	if ((o->p.flag_V || o->p.flag_s) && !vel) return -2;
	if (o->p.flag_M && !mu) return -2;
	if (o->nSkip < o->p.nCSkip) { o->nSkip++; return 0; }
	o->L = xhi - xlo;
	o->vol = o->L * o->L * o->L;
	if (!o->alloc_more) { o->N = n; o->first_vol = o->vol; o->first_limit = o->limitCorrAv; alloc_more(o); o->alloc_more = 1; }
	if (n != o->N) return -1;
	o->kVal = 2. * M_PI / o->L;
	o->cur_xyz = xyz;
	o->cur_vel = vel;
	o->cur_mu = mu;
	if (o->p.flag_f) one_time_kspace(o);
	if (o->p.flag_s) one_time_sum_vr(o); /* EvalOneTimeCorr :1250-1257 */
	for (int nb = 0; nb < o->p.nCBuffer; nb++) {
		ring_buf *b = &o->buf[nb];
		if (b->count == 0) capture_origin(o, b);
		if (b->count >= 0) two_time(o, b, b->count);
		++b->count;
	}
	int printed = accumulate(o);
	o->nSkip = 0;
	o->cur_xyz = NULL;
	o->cur_vel = NULL;
	o->cur_mu = NULL;
	return printed;
}

/* main tail, lmp2sqt.c:327-333. */
int sqt_oracle_finish(sqt_oracle *o)
{
	if (o->n_print != 0) return 0;
	o->limitCorrAv = o->countCorrAv;
	print_process(o);
	return 1;
}

int sqt_oracle_countCorrAv(const sqt_oracle *o) { return o->countCorrAv; }
int sqt_oracle_printed_count(const sqt_oracle *o) { return o->printed_count; }
int sqt_oracle_number_call_print(const sqt_oracle *o) { return o->n_print; }
int sqt_oracle_limitCorrAv(const sqt_oracle *o) { return o->limitCorrAv; }
double sqt_oracle_kVal(const sqt_oracle *o) { return o->kVal; }
int sqt_oracle_nPtls(const sqt_oracle *o) { return o->N; }
int sqt_oracle_buffer_count(const sqt_oracle *o, int nb) { return o->buf[nb].count; }
const double *sqt_oracle_raw2d(const sqt_oracle *o, int w) { return o->sum2d[w]; }
const double *sqt_oracle_printed2d(const sqt_oracle *o, int w) { return o->out2d[w]; }
const double *sqt_oracle_raw1d(const sqt_oracle *o, int w) { return o->sum1d[w]; }
const double *sqt_oracle_printed1d(const sqt_oracle *o, int w) { return o->out1d[w]; }
const double *sqt_oracle_printed_raw2d(const sqt_oracle *o, int w) { return o->rawp2d[w]; }
const double *sqt_oracle_printed_raw1d(const sqt_oracle *o, int w) { return o->rawp1d[w]; }
double sqt_oracle_first_vol(const sqt_oracle *o) { return o->first_vol; }
int sqt_oracle_first_limit(const sqt_oracle *o) { return o->first_limit; }
double sqt_oracle_vol(const sqt_oracle *o) { return o->vol; }
const double *sqt_oracle_factorDr(const sqt_oracle *o) { return o->factorDr; }
const double *sqt_oracle_rho(const sqt_oracle *o) { return o->rho; }
