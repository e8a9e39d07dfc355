/*
 * sqt_writers.c -- normalisation and result files, format-identical to the reference:
 *   prePrintProcess            lmp2sqt.c:576-675
 *   PrintProcess               lmp2sqt.c:338-367
 *   PrintEtc (SSF, Dt)         lmp2sqt.c:865-1093
 *   PrintSpacetimeCorr         lmp2sqt.c:677-766   output.txt
 *   PrintSpacetimeCorr_binary  lmp2sqt.c:768-864   output.binary
 * Input is the engine's raw sums (sqt_results); everything here is O(nCSpatial*nCTime) host work.
 *
 * Documented deviations (SURVEY.md section 7 "hard parts"):
 *   - Dt columns msdx/msdy/msdz: the reference prints rrMSR1_R_Av, which is uninitialised memory
 *     unless -s is given (lmp2sqt.c:427-431, 532-533); we print 0 then, and the per-axis MSD with -s.
 *   - 1/(N*N) uses a double product; the reference's int product overflows for N > 46340.
 */
#include "../../../include/sqt_host.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define DIM 3

static const char *k_txtCorr = "##space-time corr"; /* common.c:20 */
static const char *k_header[4] = { "full-density", "self-density", "cross-density", "self-vanHove" };

void sqt_host_factor_dr(const sqt_results *r, double rVal, double *factorDr)
{
	const double vol = r->first_Vol;
	const double rho0 = (double)r->nPtls / vol;
	double shell_Vol;
	for (int nr = 0; nr < r->nCSpatial; nr++) {
		if (nr == 0) shell_Vol = 4 * M_PI / 3. * pow(rVal, 3);
		else shell_Vol = (4. / 3.) * M_PI * (pow((nr + 1) * rVal, 3) - pow(nr * rVal, 3));
		factorDr[nr] = 1. / (pow(rho0, 2) * vol * shell_Vol * r->first_limitCorrAv);
	}
}

void sqt_host_normalise(sqt_results *r, const sqt_deck *d, int flag_f, int flag_F, int flag_t)
{
	sqt_host_normalise_v(r, d, flag_f, flag_F, flag_t, 0);
}

void sqt_host_normalise_v(sqt_results *r, const sqt_deck *d, int flag_f, int flag_F, int flag_t, int flag_V)
{
	const sqt_flags fl = { flag_f, flag_F, flag_t, flag_V, 0, 0 };
	sqt_host_normalise_x(r, d, &fl);
}

void sqt_host_normalise_x(sqt_results *r, const sqt_deck *d, const sqt_flags *fl)
{
	const int flag_f = fl->f || fl->F, flag_F = fl->F, flag_t = fl->t || fl->s, flag_V = fl->V, flag_M = fl->M, flag_s = fl->s;
	const int S = r->nCSpatial, T = r->nCTime;
	const double N = (double)r->nPtls;
	const double inverseNptls = 1. / N;
	const double inverseNptlsSq = 1. / (N * N);
	const double average_number = 1. / (3.0 * r->countCorrAv);
	if (flag_f) {
		double *av[3] = { r->avF_qq2, r->avF_s_qq2, r->avF_d_qq2 };
		double *sd[3] = { r->StdDevF_qq2, r->StdDevF_s_qq2, r->StdDevF_d_qq2 };
		for (int q = 0; q < (flag_F ? 3 : 1); q++)
			for (int e = 0; e < S * T; e++) {
				double average = (av[q][e] * inverseNptls) * average_number;
				double averageSq = (sd[q][e] * inverseNptlsSq) * average_number;
				av[q][e] = average;
				sd[q][e] = sqrt((averageSq - average * average));
			}
		if (flag_V) { /* lmp2sqt.c:607-611 */
			double *c2[3] = { r->avC2_v_rho, r->avC2_rho_v, r->avC2_v_v };
			for (int q = 0; q < 3; q++)
				for (int e = 0; e < S * T; e++) c2[q][e] *= inverseNptls * average_number;
		}
		if (flag_V && flag_M) { /* :614-617 */
			double *c2[2] = { r->avC2_mu_v, r->avC2_v_mu };
			for (int q = 0; q < 2; q++)
				for (int e = 0; e < S * T; e++) c2[q][e] *= inverseNptls * average_number;
		}
		if (flag_M) { /* :618-622 */
			double *c2[3] = { r->avC2_mu_rho, r->avC2_rho_mu, r->avC2_mu_mu };
			for (int q = 0; q < 3; q++)
				for (int e = 0; e < S * T; e++) c2[q][e] *= inverseNptls * average_number;
		}
		/* The table is only ever filled under -t (EvalTwoTimeEach as a whole is `if (flag_t)`, lmp2sqt.c:1370): the engine
		 * leaves it zero otherwise, and scaling zeros is what the reference does then (:625-630) */
		double *factorDr = (double *)malloc(sizeof(double) * S);
		sqt_host_factor_dr(r, d->rVal, factorDr);
		for (int nt = 1; nt < T; nt++) /* lag 0 is left unscaled, lmp2sqt.c:626 */
			for (int nr = 0; nr < S; nr++) r->avDrTable[nr * T + nt] *= factorDr[nr];
		free(factorDr);
	}
	if (flag_t) {
		const double scale_factor = 1. / (N * r->countCorrAv);
		const double factor_msdcm = 1. / (r->countCorrAv);
		const double factor_Cvv = 1. / (N * r->countCorrAv * 3.);      /* lmp2sqt.c:642-644 */
		const double factor_Cvcmvcm = 1. / (r->countCorrAv * 3.);
		const double scale_4stress = .5 * d->mass * d->mass / (r->countCorrAv * r->g_Vol); /* :637 */
		for (int nt = 0; nt < T; nt++) {
			r->rrMSDAv[nt] *= scale_factor;
			r->rrMSDCMAv[nt] *= factor_msdcm;
			r->rrMQDAv[nt] *= scale_factor;
			if (flag_V) {
				r->rrCvvAv[nt] *= factor_Cvv;
				r->rrCvcmvcmAv[nt] *= factor_Cvcmvcm;
			}
			if (flag_s) { /* :660-665; the diagonal / off-diagonal averages (:666-669) are formed by the writer */
				for (int j = 0; j < 3; j++) r->rrMSR1_R_Av[3 * nt + j] = scale_factor * r->rrMSR1_R_Av[3 * nt + j];
				for (int j = 0; j < 9; j++) r->rrMSR2_VR_Av[9 * nt + j] = scale_4stress * r->rrMSR2_VR_Av[9 * nt + j];
			}
		}
	}
}

void sqt_host_normalise_types(sqt_results *r, const sqt_deck *d)
{
	const int nt = r->n_types, S = r->nCSpatial, T = r->nCTime;
	if (nt < 2 || !r->avF_ab) return;
	const double N = (double)r->nPtls, cnt = (double)r->countCorrAv;
	const size_t ST = (size_t)S * T;
	for (size_t e = 0; e < (size_t)nt * nt * ST; e++) r->avF_ab[e] = (r->avF_ab[e] * (1. / N)) * (1. / (3.0 * cnt));
	double *factorDr = (double *)malloc(sizeof(double) * S);
	sqt_host_factor_dr(r, d->rVal, factorDr);
	for (int a = 0; a < nt; a++) {
		const double Na = (double)r->nPtls_a[a];
		if (Na <= 0) continue; /* an empty species: all sums are zero already */
		for (size_t e = 0; e < ST; e++) r->avF_s_a[a * ST + e] = (r->avF_s_a[a * ST + e] * (1. / Na)) * (1. / (3.0 * cnt));
		for (int nr = 0; nr < S; nr++)
			for (int lag = 1; lag < T; lag++) r->avDrTable_a[a * ST + (size_t)nr * T + lag] *= factorDr[nr] * (N / Na);
		for (int lag = 0; lag < T; lag++) {
			r->rrMSD_a[a * T + lag] *= 1. / (Na * cnt);
			r->rrMQD_a[a * T + lag] *= 1. / (Na * cnt);
		}
	}
	free(factorDr);
}

static void join(char *out, size_t n, const char *dir, const char *name)
{
	if (dir && *dir) snprintf(out, n, "%s/%s", dir, name);
	else snprintf(out, n, "%s", name);
}

/* PrintSpacetimeCorr, lmp2sqt.c:677-766 */
static void write_text(FILE *fp, const sqt_results *r, const sqt_deck *d, int flag_F)
{
	const int S = r->nCSpatial, T = r->nCTime;
	const double *av[3] = { r->avF_qq2, r->avF_s_qq2, r->avF_d_qq2 };
	const double *sd[3] = { r->StdDevF_qq2, r->StdDevF_s_qq2, r->StdDevF_d_qq2 };
	fprintf(fp, "%s\n", k_txtCorr);
	for (int k2 = 0; k2 < 4; k2++) {
		fprintf(fp, "# %s %7.3f %7.3f %7.3f\n", k_header[k2], r->kVal, (d->nCSkip + 1) * d->deltaT, d->rVal);
		if (k2 < 3) {
			if (k2 == 0 || flag_F)
				for (int nt = 0; nt < T; nt++) {
					double time_d = nt * 1. * d->deltaT;
					fprintf(fp, "%7.3f", time_d);
					for (int nk = 0; nk < S; nk++) fprintf(fp, " %8.4e %8.4e", av[k2][nk * T + nt], sd[k2][nk * T + nt]);
					fprintf(fp, "\n");
				}
		} else {
			for (int nt = 0; nt < T; nt++) {
				for (int nr = 0; nr < S; nr++) fprintf(fp, " %8.4e", r->avDrTable[nr * T + nt]);
				fprintf(fp, "\n");
			}
		}
		fprintf(fp, "\n");
	}
}

/* PrintSpacetimeCorr_binary, lmp2sqt.c:768-864 */
static void write_binary(FILE *fp, const sqt_results *r, const sqt_deck *d, int flag_F)
{
	const int S = r->nCSpatial, T = r->nCTime;
	const double *av[4] = { r->avF_qq2, r->avF_s_qq2, r->avF_d_qq2, r->avDrTable };
	int nTypes = 4;
	int header_txtCorr = (int)strlen(k_txtCorr);
	fwrite(&header_txtCorr, sizeof(int), 1, fp);
	fwrite(k_txtCorr, 1, header_txtCorr, fp);
	fwrite(&nTypes, sizeof(int), 1, fp);
	for (int k2 = 0; k2 < nTypes; k2++) {
		double col2 = r->kVal, col3 = 1.0 * d->deltaT * (d->nCSkip + 1), col4 = d->rVal;
		int yes = 1, no = 0;
		int length = (int)strlen(k_header[k2]);
		fwrite(&length, sizeof(int), 1, fp);
		fwrite(k_header[k2], sizeof(char), length, fp);
		fwrite(&col2, sizeof(double), 1, fp);
		fwrite(&col3, sizeof(double), 1, fp);
		fwrite(&col4, sizeof(double), 1, fp);
		if ((k2 == 1 || k2 == 2) && !flag_F) { fwrite(&no, sizeof(int), 1, fp); continue; }
		fwrite(&yes, sizeof(int), 1, fp);
		for (int nt = 0; nt < T; nt++)
			for (int nk = 0; nk < S; nk++) fwrite(&av[k2][nk * T + nt], sizeof(double), 1, fp);
	}
}

/* PrintEtc, lmp2sqt.c:865-1093: SSF, Dt, the C2_* files of -V / -M and the Stress file of -s */
static int write_etc(const char *dir, const sqt_results *r, const sqt_deck *d, const sqt_flags *fl, const char *date)
{
	const int flag_f = fl->f || fl->F, flag_t = fl->t || fl->s, flag_V = fl->V, flag_M = fl->M, flag_s = fl->s;
	const int S = r->nCSpatial, T = r->nCTime;
	char name[64], fn_Dt[512], fn_SSF[512], fn_stress[512], fn_C2[8][512];
	static const char *c2_name[8] = { "C2_v_rho", "C2_rho_v", "C2_v_v", "C2_v_mu", "C2_mu_v", "C2_mu_rho", "C2_rho_mu", "C2_mu_mu" };
	int nfile = 0;
	do { /* first unused index, probed on the Dt name only (lmp2sqt.c:886-902) */
		snprintf(name, sizeof(name), "Dt%03d.info.%s", nfile, date);
		join(fn_Dt, sizeof(fn_Dt), dir, name);
		snprintf(name, sizeof(name), "SSF%03d.info.%s", nfile, date);
		join(fn_SSF, sizeof(fn_SSF), dir, name);
		snprintf(name, sizeof(name), "Stress%03d.info.%s", nfile, date);
		join(fn_stress, sizeof(fn_stress), dir, name);
		for (int q = 0; q < 8; q++) {
			snprintf(name, sizeof(name), "%s%03d.info.%s", c2_name[q], nfile, date);
			join(fn_C2[q], sizeof(fn_C2[q]), dir, name);
		}
		nfile++;
	} while (0 == access(fn_Dt, F_OK));
	if (flag_f && (flag_V || flag_M)) { /* lmp2sqt.c:933-993: one row per lag, "%7.3f" time then " %8.4e" per nk */
		const double *c2[8] = { r->avC2_v_rho, r->avC2_rho_v, r->avC2_v_v, r->avC2_v_mu, r->avC2_mu_v, r->avC2_mu_rho, r->avC2_rho_mu, r->avC2_mu_mu };
		for (int q = 0; q < 8; q++) {
			if (q < 3 ? !flag_V : (q < 5 ? !(flag_V && flag_M) : !flag_M)) continue;
			FILE *fp = fopen(fn_C2[q], "w");
			if (!fp) return -1;
			for (int nt = 0; nt < T; nt++) {
				fprintf(fp, "%7.3f", nt * 1. * d->deltaT);
				for (int nk = 0; nk < S; nk++) fprintf(fp, " %8.4e", c2[q][nk * T + nt]);
				fprintf(fp, "\n");
			}
			fclose(fp);
		}
	}
	if (flag_f) {
		FILE *fp = fopen(fn_SSF, "w");
		if (!fp) return -1;
		for (int nk = 0; nk < S; nk++) {
			double value1 = r->avF_qq2[nk * T + 0];
			double value2 = sqrt(r->StdDevF_qq2[nk * T + 0] - value1 * value1); /* sic: sigma is already rooted, :996-997 */
			fprintf(fp, "%8.4f" " %8.4g" " %8.4g" "\n", (nk + 1) * r->kVal, value1, value2);
		}
		fclose(fp);
	}
	if (flag_s) { /* lmp2sqt.c:1006-1026 */
		FILE *fp = fopen(fn_stress, "w");
		if (!fp) return -1;
		fprintf(fp, "#time MSVR_dig MSVR_offdig xy yx zy yz xz zx\n");
		fprintf(stderr, "print out stress file\n");
		for (int nt = 0; nt < T; nt++) {
			const double *m = r->rrMSR2_VR_Av + 9 * nt; /* xx xy xz yx yy yz zx zy zz */
			const double dig = (m[0] + m[4] + m[8]) / 3.0;                        /* common.c:164-172 */
			const double offdig = (m[1] + m[2] + m[3] + m[5] + m[6] + m[7]) / 6.0; /* common.c:161-169 */
			double tVal = nt * d->deltaT * (d->nCSkip + 1);
			fprintf(fp, "%8.4g %8.4g %8.4g %8.4g %8.4g %8.4g %8.4g %8.4g %8.4g \n", tVal, dig, offdig, m[1], m[3], m[7], m[5], m[2], m[6]);
		}
		fclose(fp);
	}
	if (flag_t) {
		FILE *fp = fopen(fn_Dt, "w");
		if (!fp) return -1;
		fprintf(fp, "#time MSD msdx msdy msdz D(t) anotherform MQD");
		if (flag_V) fprintf(fp, " Cvv Cvcmvcm");
		if (flag_M) fprintf(fp, " Cmm");
		fprintf(fp, " MSDCM\n");
		fprintf(stderr, "print time correlation \n");
		double *rrDt = (double *)calloc(T > 3 ? T : 3, sizeof(double));
		const double *msd = r->rrMSDAv;
		double fac = 1. / (2. * d->deltaT * (d->nCSkip + 1) * DIM * 2);
		if (T >= 3) {
			rrDt[0] = fac * (-msd[2] + 4. * msd[1] - 3. * msd[0]);
			for (int nr = 1; nr < T - 1; nr++) rrDt[nr] = fac * (msd[nr + 1] - msd[nr - 1]);
			rrDt[T - 1] = fac * (msd[T - 3] - 4. * msd[T - 2] + 3. * msd[T - 1]);
		}
		for (int nt = 0; nt < T; nt++) {
			double tVal = nt * d->deltaT * (d->nCSkip + 1);
			double anotherDt = 0;
			if (nt != 0) anotherDt = msd[nt] / (tVal * DIM * 2);
			const double *m1 = r->rrMSR1_R_Av + 3 * nt; /* zero unless -s */
			fprintf(fp, "%8.4f %8.4g %8.4g %8.4g %8.4g %8.4g %8.4g %8.4g", tVal, msd[nt], flag_s ? m1[0] : 0., flag_s ? m1[1] : 0.,
			        flag_s ? m1[2] : 0., rrDt[nt], anotherDt, r->rrMQDAv[nt]);
			if (flag_V) fprintf(fp, " %8.4f %8.4g", r->rrCvvAv[nt], r->rrCvcmvcmAv[nt]);
			if (flag_M) fprintf(fp, " %8.4f", r->rrCmmAv[nt]); /* always 0: the reference never adds to rrCmmAv (lmp2sqt.c:420-432) */
			fprintf(fp, " %8.4f\n", r->rrMSDCMAv[nt]);
		}
		free(rrDt);
		fclose(fp);
	}
	return 0;
}

int sqt_host_print_types(const char *dir, sqt_results *r, const sqt_deck *d, const sqt_flags *fl, const char *date)
{
	const int nt = r->n_types, S = r->nCSpatial, T = r->nCTime;
	if (nt < 2 || !r->avF_ab) return 0;
	const int flag_f = fl->f || fl->F, flag_F = fl->F, flag_t = fl->t || fl->s;
	char today[32], name[64], path[512];
	if (!date) {
		time_t raw = time(NULL);
		strftime(today, sizeof(today), "%y%m%d", localtime(&raw));
		date = today;
	}
	int nfile = 0;
	do {
		snprintf(name, sizeof(name), "types%03d.info.%s", nfile++, date);
		join(path, sizeof(path), dir, name);
	} while (0 == access(path, F_OK));
	sqt_host_normalise_types(r, d);
	FILE *fp = fopen(path, "w");
	if (!fp) return -1;
	const size_t ST = (size_t)S * T;
	fprintf(fp, "# per-type partials: n_types %d, atoms", nt);
	for (int a = 0; a < nt; a++) fprintf(fp, " %lld", r->nPtls_a[a]);
	fprintf(fp, ", kVal %7.3f deltaT %7.3f rVal %7.3f\n", r->kVal, (d->nCSkip + 1) * d->deltaT, d->rVal);
	if (flag_f)
		for (int a = 0; a < nt; a++)
			for (int b = 0; b < nt; b++) {
				fprintf(fp, "# F_ab %d %d\n", a + 1, b + 1);
				const double *t = r->avF_ab + ((size_t)a * nt + b) * ST;
				for (int lag = 0; lag < T; lag++) {
					fprintf(fp, "%7.3f", lag * 1. * d->deltaT);
					for (int nk = 0; nk < S; nk++) fprintf(fp, " %8.4e", t[(size_t)nk * T + lag]);
					fprintf(fp, "\n");
				}
				fprintf(fp, "\n");
			}
	if (flag_F)
		for (int a = 0; a < nt; a++) {
			fprintf(fp, "# F_s_a %d\n", a + 1);
			const double *t = r->avF_s_a + (size_t)a * ST;
			for (int lag = 0; lag < T; lag++) {
				fprintf(fp, "%7.3f", lag * 1. * d->deltaT);
				for (int nk = 0; nk < S; nk++) fprintf(fp, " %8.4e", t[(size_t)nk * T + lag]);
				fprintf(fp, "\n");
			}
			fprintf(fp, "\n");
		}
	if (flag_f && flag_t)
		for (int a = 0; a < nt; a++) {
			fprintf(fp, "# self-vanHove_a %d\n", a + 1);
			const double *t = r->avDrTable_a + (size_t)a * ST;
			for (int lag = 0; lag < T; lag++) {
				for (int nr = 0; nr < S; nr++) fprintf(fp, " %8.4e", t[(size_t)nr * T + lag]);
				fprintf(fp, "\n");
			}
			fprintf(fp, "\n");
		}
	if (flag_t) {
		fprintf(fp, "# time, then MSD_a MQD_a per species\n");
		for (int lag = 0; lag < T; lag++) {
			fprintf(fp, "%8.4f", lag * d->deltaT * (d->nCSkip + 1));
			for (int a = 0; a < nt; a++) fprintf(fp, " %8.4g %8.4g", r->rrMSD_a[a * T + lag], r->rrMQD_a[a * T + lag]);
			fprintf(fp, "\n");
		}
	}
	fclose(fp);
	return 0;
}

int sqt_host_print_process(const char *dir, sqt_results *r, const sqt_deck *d, int flag_f, int flag_F, int flag_t,
                           const char *date)
{
	return sqt_host_print_process_v(dir, r, d, flag_f, flag_F, flag_t, 0, date);
}

int sqt_host_print_process_v(const char *dir, sqt_results *r, const sqt_deck *d, int flag_f, int flag_F, int flag_t, int flag_V,
                             const char *date)
{
	const sqt_flags fl = { flag_f, flag_F, flag_t, flag_V, 0, 0 };
	return sqt_host_print_process_x(dir, r, d, &fl, date);
}

int sqt_host_print_process_x(const char *dir, sqt_results *r, const sqt_deck *d, const sqt_flags *fl, const char *date)
{
	const int flag_f = fl->f || fl->F, flag_F = fl->F;
	char today[32], path[512];
	if (!date) {
		time_t raw = time(NULL);
		strftime(today, sizeof(today), "%y%m%d", localtime(&raw)); /* lmp2sqt.c:127-132 */
		date = today;
	}
	join(path, sizeof(path), dir, "output.log");
	FILE *flog = fopen(path, "w+");
	if (!flog) return -1;
	fputs("Call PrintProcess : \n", flog);
	sqt_host_normalise_x(r, d, fl);
	if (write_etc(dir, r, d, fl, date)) { fclose(flog); return -1; }
	if (flag_f) {
		const char *fout_filename = "output.binary";
		join(path, sizeof(path), dir, fout_filename);
		FILE *fout = fopen(path, "w+");
		join(path, sizeof(path), dir, "output.txt");
		FILE *ftxt = fopen(path, "w+");
		if (!fout || !ftxt) { fclose(flog); return -1; }
		fprintf(flog, "%s %ld %d \n", fout_filename, ftell(fout), r->limitCorrAv);
		fprintf(stderr, "%s %ld %d \n", fout_filename, ftell(fout), r->limitCorrAv);
		write_text(ftxt, r, d, flag_F);
		write_binary(fout, r, d, flag_F);
		fclose(ftxt);
		fclose(fout);
	}
	fclose(flog);
	return 0;
}
