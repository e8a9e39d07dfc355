/*
 * sqt_analysis.c -- post-processing of the space-time correlators: the consumer of output.binary
 * (SURVEY.md section 8f.2).  Same inputs, same numbers and the same files as the reference's
 * analysis_binary.c; also callable directly on results fetched from the engine (no file in between).
 *
 *   reference (analysis_binary.c)                      here
 *   import_settings :407-447   deck -> deltaT, sizes    sqt_analysis_load / _from_results
 *   load_raw_data   :643-759   sets of output.binary    sqt_analysis_load
 *   scaling_data    :760-771   mean / sd over sets      sqt_analysis_load (tail)
 *   do_not_fourier  :793-897   F', Gamma, D, H(q,t)     sqt_analysis_time
 *   do_fourier      :772-792   cosine transform         sqt_analysis_fourier
 *   print_output    :898-963, Print_R2_data :209-232    sqt_analysis_print
 *
 * Quirks kept because downstream scripts see them (each marked below): the time axis of the tables uses the deck's
 * deltaT while the derivatives use the block header's deltaT*(nCSkip+1); the under-limit test looks at the current
 * sample; the van Hove block goes through the same logarithmic derivatives; the frequency axis is integer arithmetic.
 * Deviation: `-s n` with n > 1 skips the first n-1 sets here; the reference loses its place in the file and exits.
 */
#include "../../../include/sqt_host.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static const char *k_set_tag = "##space-time corr"; /* common.c:20 */
static const char *k_block[SQT_AN_TYPES] = { "full-density", "self-density", "cross-density", "self-vanHove" };
static const int k_more[SQT_AN_TYPES] = { 1, 1, 1, 0 }; /* header_flag_more, analysis_binary.c:81-85 */

static double *table(size_t n, double fill)
{
	double *p = (double *)malloc(sizeof(double) * (n ? n : 1));
	if (p) for (size_t i = 0; i < n; i++) p[i] = fill;
	return p;
}

static int alloc_tables(sqt_analysis *a)
{
	const size_t n = (size_t)a->nCSpatial * (size_t)a->nCTime;
	for (int j = 0; j < SQT_AN_TYPES; j++) { /* init_memory, :357-405: sums 0, derived tables NaN */
		a->mean[j] = table(n, 0.0); a->sd[j] = table(n, 0.0); a->err[j] = table(n, 0.0); a->F[j] = table(n, 0.0);
		a->F1[j] = table(n, NAN); a->Gamma[j] = table(n, NAN); a->D[j] = table(n, NAN); a->Dc[j] = table(n, NAN);
		a->H[j] = table(n, NAN); a->Hc[j] = table(n, NAN);
		if (!a->mean[j] || !a->sd[j] || !a->err[j] || !a->F[j] || !a->F1[j] || !a->Gamma[j] || !a->D[j] || !a->Dc[j] || !a->H[j] || !a->Hc[j])
			return -1;
	}
	return 0;
}

void sqt_analysis_free(sqt_analysis *a)
{
	if (!a) return;
	for (int j = 0; j < SQT_AN_TYPES; j++) {
		free(a->mean[j]); free(a->sd[j]); free(a->err[j]); free(a->F[j]); free(a->F1[j]);
		free(a->Gamma[j]); free(a->D[j]); free(a->Dc[j]); free(a->H[j]); free(a->Hc[j]);
	}
	memset(a, 0, sizeof(*a));
}

/* the four keys analysis_binary.c takes from the deck (:420-427): a line that starts with the key, value by strtod */
static int deck_settings(const char *deck_path, sqt_analysis *a)
{
	FILE *fp = fopen(deck_path, "r");
	if (!fp) return -1;
	char line[4096];
	double nS = 0, nT = 0;
	while (fgets(line, sizeof(line), fp)) {
		if (line[0] == '-') break;
		if (!strncmp(line, "deltaT", 6)) a->deck_deltaT = strtod(line + 6, NULL);
		else if (!strncmp(line, "nCSpatial", 9)) nS = strtod(line + 9, NULL);
		else if (!strncmp(line, "nCTime", 6)) nT = strtod(line + 6, NULL);
		else if (!strncmp(line, "kVal", 4)) a->kVal = strtod(line + 4, NULL);
	}
	fclose(fp);
	a->nCSpatial = (int)nS; a->nCTime = (int)nT;
	a->deltaT = a->deck_deltaT;
	return (a->nCSpatial > 0 && a->nCTime > 0) ? 0 : -1;
}

static void finish_sets(sqt_analysis *a)
{
	if (a->nData <= 1) return; /* scaling_data, :760-771: a single set stays as read */
	const size_t n = (size_t)a->nCSpatial * (size_t)a->nCTime;
	for (int j = 0; j < SQT_AN_TYPES; j++)
		for (size_t i = 0; i < n; i++) {
			a->mean[j][i] /= a->nData;
			a->sd[j][i] = sqrt(a->sd[j][i] / a->nData - a->mean[j][i] * a->mean[j][i]);
			a->err[j][i] = a->sd[j][i] / sqrt((double)a->nData);
		}
}

int sqt_analysis_load(sqt_analysis *a, const char *deck_path, const char *binary_path, int nSetSkip)
{
	memset(a, 0, sizeof(*a));
	if (deck_settings(deck_path, a)) return -1;
	if (alloc_tables(a)) return -2;
	FILE *fp = fopen(binary_path, "rb");
	if (!fp) return -3;
	const int S = a->nCSpatial, T = a->nCTime;
	const int tag_len = (int)strlen(k_set_tag);
	double *block = table((size_t)S * T, 0.0);
	int nSet = 0, rc = 0;
	for (;;) {
		int len = 0;
		char tag[64];
		if (fread(&len, sizeof(int), 1, fp) != 1) break;       /* end of file */
		if (len != tag_len) { rc = -4; break; }                /* "File error Header size different" */
		if (fread(tag, 1, (size_t)len, fp) != (size_t)len) break;
		if (strncmp(tag, k_set_tag, (size_t)len) != 0) continue;
		++nSet;
		const int use = !(nSet < nSetSkip);
		if (use) ++a->nData;
		int nTypes = 0;
		if (fread(&nTypes, sizeof(int), 1, fp) != 1) { rc = -4; break; }
		for (int j = 0; j < nTypes && rc == 0; j++) {
			int nlen = 0, has_data = 0;
			char name[128];
			double col[3];
			if (fread(&nlen, sizeof(int), 1, fp) != 1 || j >= SQT_AN_TYPES || nlen != (int)strlen(k_block[j])) { rc = -5; break; } /* "version check" */
			if (fread(name, 1, (size_t)nlen, fp) != (size_t)nlen || fread(col, sizeof(double), 3, fp) != 3 || fread(&has_data, sizeof(int), 1, fp) != 1) { rc = -4; break; }
			if (use) { a->kVal = col[0]; a->deltaT = col[1]; a->rVal = col[2]; } /* the block header wins over the deck, :706-708 */
			const int same = strncmp(k_block[j], name, (size_t)nlen) == 0;
			if (!has_data) continue;                              /* a block without data carries no array */
			if (fread(block, sizeof(double), (size_t)S * T, fp) != (size_t)S * T) { rc = -4; break; }
			if (!use || !same) continue;
			for (int n = 0; n < T; n++)                           /* file is time-major, tables are [k][t] */
				for (int k = 0; k < S; k++) {
					const double w = block[(size_t)n * S + k];
					a->mean[j][(size_t)k * T + n] += w;
					a->sd[j][(size_t)k * T + n] += w * w;
				}
		}
		if (rc) break;
	}
	free(block);
	fclose(fp);
	if (rc) return rc;
	finish_sets(a);
	return 0;
}

int sqt_analysis_from_results(sqt_analysis *a, const sqt_results *normalised, const sqt_deck *d, int flag_F)
{
	memset(a, 0, sizeof(*a));
	a->nCSpatial = normalised->nCSpatial; a->nCTime = normalised->nCTime;
	a->deck_deltaT = d->deltaT;
	a->kVal = normalised->kVal; a->deltaT = 1.0 * d->deltaT * (d->nCSkip + 1); a->rVal = d->rVal; /* what write_binary puts in the block headers */
	if (alloc_tables(a)) return -2;
	const double *av[SQT_AN_TYPES] = { normalised->avF_qq2, normalised->avF_s_qq2, normalised->avF_d_qq2, normalised->avDrTable };
	const size_t n = (size_t)a->nCSpatial * (size_t)a->nCTime;
	for (int j = 0; j < SQT_AN_TYPES; j++) {
		if ((j == 1 || j == 2) && !flag_F) continue;
		for (size_t i = 0; i < n; i++) { a->mean[j][i] = av[j][i]; a->sd[j][i] = av[j][i] * av[j][i]; }
	}
	a->nData = 1;
	return 0;
}

int sqt_analysis_time(sqt_analysis *a, int nValDiff, double under_limit, const char *outdir)
{
	const int S = a->nCSpatial, T = a->nCTime;
	const double dt = a->deltaT;
	if (T < 3) return -1;
	for (int j = 0; j < SQT_AN_TYPES; j++) {
		char path[4096];
		snprintf(path, sizeof(path), "%s%s%s00.out", outdir ? outdir : "", outdir ? "/" : "", k_block[j]);
		FILE *out = fopen(path, "w");
		if (!out) return -2;
		fprintf(out, "#kVal F0 F1_0 Gamma Dq0\n#deltaT = %9.4f\n", dt);
		double *X = a->mean[j], *F1 = a->F1[j], *G = a->Gamma[j], *D = a->D[j], *Dc = a->Dc[j], *H = a->H[j], *Hc = a->Hc[j];
		for (int k = 0; k < S; k++) {
			const double q = (k + 1) * a->kVal, q2 = q * q;
			const size_t t0 = (size_t)k * T, tN = t0 + T - 1;
			const double F0 = X[t0], lnF0 = log(F0);
			/* lag 0: forward second-order difference of F and of ln F */
			F1[t0] = (-X[t0 + 2] + 4. * X[t0 + 1] - 3. * X[t0]) / (2. * dt);
			G[t0] = (-log(X[t0 + 2]) + 4. * log(X[t0 + 1]) - 3. * log(X[t0])) / (2. * dt);
			D[t0] = -G[t0] / q2;
			Dc[t0] = D[t0];
			H[t0] = D[t0] * F0;
			Hc[t0] = Dc[t0] * F0;
			/* interior lags: central differences, until the correlator has decayed below the limit (the sample tested is the
			 * current one: the reference's index still points there from the previous step, :845) */
			size_t c = t0 + 1;
			int decayed = 0;
			for (; c < tN; c++) {
				if (X[c] < under_limit) { decayed = 1; break; }
				F1[c] = (X[c + 1] - X[c - 1]) / (2. * dt);
				G[c] = (log(X[c + 1]) - log(X[c - 1])) / (2. * dt);
				D[c] = -G[c] / q2;
				Dc[c] = -(log(X[c]) - lnF0) / (q2 * dt * (double)(c - t0));
				H[c] = D[c] * F0;
				Hc[c] = Dc[c] * F0;
			}
			if (!decayed) { /* last lag: backward second-order difference */
				c = tN;
				F1[c] = (X[c - 2] - 4. * X[c - 1] + 3. * X[c]) / (2. * dt);
				G[c] = (log(X[c - 2]) - 4. * log(X[c - 1]) + 3. * log(X[c])) / (2. * dt);
				D[c] = -G[c] / q2;
				Dc[c] = -(log(X[c]) - lnF0) / (q2 * dt * (double)(c - t0));
				H[c] = D[c] * F0;
				Hc[c] = Dc[c] * F0;
			}
			/* the summary row samples lag min(where the loop stopped, nValDiff), :880-885 */
			const size_t pick = c < t0 + (size_t)nValDiff ? c : t0 + (size_t)nValDiff;
			fprintf(out, "%.13lf %.13le %.13le %.13le %.13le \n", q, X[pick], F1[pick], G[pick], D[pick]);
		}
		fclose(out);
		for (int k = 0; k < S; k++) { /* keep F, scale the working copy by its lag-0 value, :889-896 */
			const size_t t0 = (size_t)k * T;
			for (int n = 1; n < T; n++) { a->F[j][t0 + n] = X[t0 + n]; X[t0 + n] /= X[t0]; }
			a->F[j][t0] = X[t0];
			X[t0] = 1.;
		}
	}
	a->fourier = 0;
	const double tmax_data = (T - 1) * a->deck_deltaT;
	a->tMax = 100. < tmax_data ? 100. : tmax_data;
	a->nv = (int)(T * a->tMax / (T - 1) / a->deck_deltaT);
	return 0;
}

/* In-place complex transform of n = 2^p points, decimation in time, twiddles advanced by the rotation recurrence
 * w += w * (cos(theta) - 1, sin(theta)) -- the reference's FftComplex (:253-287); the recurrence (not a table) is kept
 * so that the rounding, hence the printed digits, agree. */
static void fft_pow2(double *re, double *im, int n)
{
	for (int i = 0, rev = 0; i < n; i++) { /* bit-reversal permutation */
		if (i < rev) { double t = re[i]; re[i] = re[rev]; re[rev] = t; t = im[i]; im[i] = im[rev]; im[rev] = t; }
		int half = n / 2;
		while (half >= 1 && rev >= half) { rev -= half; half /= 2; }
		rev += half;
	}
	for (int span = 1; span < n; span *= 2) {
		const double theta = M_PI / span;
		const double dr = cos(theta) - 1., di = sin(theta);
		double wr = 1., wi = 0.;
		for (int k = 0; k < span; k++) {
			for (int i = k; i < n; i += 2 * span) {
				const int j = i + span;
				const double tr = wr * re[j] - wi * im[j], ti = wr * im[j] + wi * re[j];
				re[j] = re[i] - tr; im[j] = im[i] - ti;
				re[i] = re[i] + tr; im[i] = im[i] + ti;
			}
			const double tr = wr * dr - wi * di, ti = wr * di + wi * dr;
			wr += tr; wi += ti;
		}
	}
}

int sqt_analysis_fourier(sqt_analysis *a, int window)
{
	const int S = a->nCSpatial, T = a->nCTime;
	const int n = 2 * (T - 1); /* the even extension; the reference assumes a power of two and so does this */
	if (T < 2 || (n & (n - 1)) != 0) return -1;
	double *re = table((size_t)n, 0.0), *im = table((size_t)n, 0.0);
	if (!re || !im) { free(re); free(im); return -2; }
	for (int j = 0; j < 3; j++)
		for (int k = 0; k < S; k++) {
			double *X = a->mean[j] + (size_t)k * T;
			for (int t = 0; t < T; t++) {
				const double damp = window ? (T - t) / (T + .5) : 1.;
				re[t] = X[t] * damp; im[t] = 0.;
			}
			for (int t = T; t < n; t++) { re[t] = re[n - t]; im[t] = im[n - t]; }
			fft_pow2(re, im, n);
			for (int t = 0; t < T; t++) X[t] = re[t];
		}
	free(re); free(im);
	a->fourier = 1;
	const double nyq = M_PI / a->deck_deltaT;
	a->omegaMax = (int)(10 < nyq ? 10 : nyq);          /* an int in the reference (:127), truncation included */
	a->nv = (int)(T * a->omegaMax / nyq);
	return 0;
}

/* Print_R2_data, :209-232: "#nCS k1 k2 ...", then one row per lag / frequency */
static int print_table(const sqt_analysis *a, const char *outdir, const char *stem, const char *block, const double *v)
{
	char path[4096];
	snprintf(path, sizeof(path), "%s%s%s%s.%s.info", outdir ? outdir : "", outdir ? "/" : "", stem, "", block);
	FILE *fp = fopen(path, "w");
	if (!fp) return -1;
	const int S = a->nCSpatial, T = a->nCTime;
	fprintf(fp, "#%d", S);
	for (int k = 0; k < S; k++) fprintf(fp, " %.13e", (k + 1) * a->kVal);
	fprintf(fp, "\n");
	for (int n = 0; n < T; n++) {
		double x;
		if (a->fourier) x = a->nv ? (double)(n * a->omegaMax / a->nv) : 0.0; /* integer arithmetic in the reference (:219) */
		else x = n * a->deck_deltaT;
		fprintf(fp, "%.13e", x);
		for (int k = 0; k < S; k++) fprintf(fp, " %.13e", v[(size_t)k * T + n]);
		fprintf(fp, "\n");
	}
	fclose(fp);
	return 0;
}

int sqt_analysis_print(const sqt_analysis *a, const char *outdir)
{
	const char *d = a->fourier ? "w" : "t";
	char stem[64];
	int rc = 0;
	for (int j = 0; j < SQT_AN_TYPES; j++) {
		snprintf(stem, sizeof(stem), "Fq%s_scaled", d); rc |= print_table(a, outdir, stem, k_block[j], a->mean[j]);
		snprintf(stem, sizeof(stem), "Fq%s", d);        rc |= print_table(a, outdir, stem, k_block[j], a->F[j]);
		if (!k_more[j]) continue;
		snprintf(stem, sizeof(stem), "GammaQ%s", d);    rc |= print_table(a, outdir, stem, k_block[j], a->Gamma[j]);
		snprintf(stem, sizeof(stem), "Dq%s_c", d);      rc |= print_table(a, outdir, stem, k_block[j], a->Dc[j]);
		snprintf(stem, sizeof(stem), "Dq%s", d);        rc |= print_table(a, outdir, stem, k_block[j], a->D[j]);
		snprintf(stem, sizeof(stem), "D0Hq%s_c", d);    rc |= print_table(a, outdir, stem, k_block[j], a->Hc[j]);
		snprintf(stem, sizeof(stem), "D0Hq%s", d);      rc |= print_table(a, outdir, stem, k_block[j], a->H[j]);
		snprintf(stem, sizeof(stem), "Fq%s1", d);       rc |= print_table(a, outdir, stem, k_block[j], a->F1[j]);
	}
	return rc ? -1 : 0;
}
