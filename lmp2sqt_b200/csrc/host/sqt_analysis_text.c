/*
 * sqt_analysis_text.c -- the TEXT post-processor: consumer of output.txt (second half).
 * Same inputs, same numbers, same files and the same stdout as the reference's analysis.c (= AverageFqt2FqOrFt.c, which
 * differs from it by an M_PI guard and a debug getchar only).
 *
 *   reference (analysis.c)                              here
 *   deck scan            :227-240                        deck_settings
 *   set / block scan     :283-334   (row parse :310-325) sqt_analysis_text_load
 *   mean, sd over sets   :338-346                        sqt_analysis_text_load (tail)
 *   cosine transform     :349-368, FftComplex :558-590   sqt_analysis_text_fourier
 *   F', Gamma, D, H      :370-455                        sqt_analysis_text_time
 *   files                :458-506, Print_R2_data :508-536 sqt_analysis_text_print
 *
 * Kept on purpose, because the files are the contract ("keep byte-compatible input, do not fix"):
 *   - the row parser predates the sigma columns lmp2sqt 1.11 writes: it takes the first nCSpatial numbers of a row, so for
 *     the three density blocks column 0 is the TIME and the rest are F, sigma interleaved; only the van Hove block (no time
 *     column, one value per bin) is read as meant;
 *   - a block whose header matches is read for nCTime rows whether or not it carries rows (`-f` without `-F`);
 *   - in the frequency domain the files are written from tables that were never filled (zeros / NaN): the transform
 *     lands in the working copy, which is not printed;
 *   - the last-lag formulas run where the decay loop stopped, with the neighbours it last used.
 * Not reproduced: the reference dereferences NULL at an unexpected end of file (an error here) and reads one element in
 * front of its table when the very first decay test of wave number 1 fails (0 is used here, see sqt_analysis_text_time).
 */
#include "../../../include/sqt_host.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define LINE_CAP 50000 /* BUFF_LEN, analysis.c:69 */

static const char *k_set_tag = "##space-time corr"; /* common.c:20 */
static const char *k_block[SQT_AN_TYPES] = { "full-density", "self-density", "cross-density", "self-vanHove" };
static const int k_more[SQT_AN_TYPES] = { 1, 1, 1, 0 }; /* header_flag_more, analysis.c:87-91 */

static double *table(size_t n, double fill)
{
	double *p = (double *)malloc(sizeof(double) * (n ? n : 1));
	if (p) for (size_t i = 0; i < n; i++) p[i] = fill;
	return p;
}

/* analysis.c:227-240: lines until one that starts with '-'; a key is a line PREFIX, its value what strtod finds behind it */
static int deck_settings(const char *deck_path, sqt_analysis *a)
{
	FILE *fp = fopen(deck_path, "r");
	if (!fp) return -1;
	char line[LINE_CAP];
	double nS = 0, nT = 0;
	while (fgets(line, sizeof(line), fp)) {
		if (line[0] == '-') break;
		if (!strncmp(line, "deltaT", 6)) a->deck_deltaT = strtod(line + 6, NULL);
		if (!strncmp(line, "nCSpatial", 9)) nS = strtod(line + 9, NULL);
		if (!strncmp(line, "nCTime", 6)) nT = strtod(line + 6, NULL);
		if (!strncmp(line, "kVal", 4)) a->kVal = strtod(line + 4, NULL);
	}
	fclose(fp);
	a->nCSpatial = (int)nS; a->nCTime = (int)nT;
	a->deltaT = a->deck_deltaT; /* deltaTCorr = deltaT, :238 */
	return (a->nCSpatial > 0 && a->nCTime > 0) ? 0 : -1;
}

int sqt_analysis_text_load(sqt_analysis *a, const char *deck_path, const char *txt_path, int nSetSkip)
{
	memset(a, 0, sizeof(*a));
	if (deck_settings(deck_path, a)) return -1;
	const int S = a->nCSpatial, T = a->nCTime;
	const size_t n = (size_t)S * T;
	for (int j = 0; j < SQT_AN_TYPES; j++) { /* :262-279 */
		a->mean[j] = table(n, 0.0); a->sd[j] = table(n, 0.0); a->err[j] = table(n, 0.0); a->F[j] = table(n, 0.0);
		a->F1[j] = table(n, NAN); a->Gamma[j] = table(n, NAN); a->D[j] = table(n, NAN); a->H[j] = table(n, NAN);
		if (!a->mean[j] || !a->sd[j] || !a->err[j] || !a->F[j] || !a->F1[j] || !a->Gamma[j] || !a->D[j] || !a->H[j]) return -2;
	}
	FILE *fp = strcmp(txt_path, "-") ? fopen(txt_path, "r") : stdin;
	if (!fp) return -3;
	char *line = (char *)malloc(LINE_CAP);
	if (!line) return -2;
	int nSet = 0, rc = 0;
	while (rc == 0 && fgets(line, LINE_CAP, fp)) {
		if (strncmp(line, k_set_tag, strlen(k_set_tag))) continue;
		++nSet;
		if (nSet < nSetSkip) continue;
		++a->nData;
		for (int j = 0; j < SQT_AN_TYPES && rc == 0; j++) {
			char want[100] = "# ";
			strcat(want, k_block[j]);
			if (!fgets(line, LINE_CAP, fp)) { rc = -4; break; }
			if (!strncmp(line, want, strlen(want)))
				for (int t = 0; t < T; t++) {
					if (!fgets(line, LINE_CAP, fp)) { rc = -4; break; }
					char *bp = line;
					for (int k = 0; k < S; k++) { /* the first nCSpatial numbers of the row, whatever they are (:318-325) */
						const double w = strtod(bp, &bp);
						a->mean[j][(size_t)k * T + t] += w;
						a->sd[j][(size_t)k * T + t] += w * w;
					}
				}
			if (rc == 0 && !fgets(line, LINE_CAP, fp) && j + 1 < SQT_AN_TYPES) rc = -4; /* the blank line that ends a block */
		}
	}
	free(line);
	if (fp != stdin) fclose(fp);
	if (rc) return rc;
	if (a->nData > 1) /* :338-346 */
		for (int j = 0; j < SQT_AN_TYPES; j++)
			for (size_t i = 0; i < n; i++) {
				a->mean[j][i] /= a->nData;
				a->sd[j][i] = sqrt(a->sd[j][i] / a->nData - a->mean[j][i] * a->mean[j][i]);
				a->err[j][i] = a->sd[j][i] / sqrt((double)a->nData);
			}
	return 0;
}

int sqt_analysis_text_time(sqt_analysis *a, int nValDiff, const char *outdir)
{
	const int S = a->nCSpatial, T = a->nCTime;
	const double dt = a->deck_deltaT;
	if (T < 3) return -1;
	for (int j = 0; j < SQT_AN_TYPES; j++) {
		char path[4096];
		snprintf(path, sizeof(path), "%s%s%s00.out", outdir ? outdir : "", outdir ? "/" : "", k_block[j]);
		FILE *out = fopen(path, "w");
		if (!out) return -2;
		fprintf(out, "#kVal F0 F1_0 Gamma Dq0\n#deltaT = %9.4f\n", dt);
		double *X = a->mean[j], *F1 = a->F1[j], *G = a->Gamma[j], *D = a->D[j], *H = a->H[j];
		for (int k = 0; k < S; k++) {
			const double q = (k + 1) * a->kVal, q2 = q * q;
			const long t0 = (long)k * T, tN = t0 + T - 1;
			const double F0 = X[t0];
			/* lag 0: forward second-order differences of F and ln F (:395-400) */
			F1[t0] = (-X[t0 + 2] + 4. * X[t0 + 1] - 3. * X[t0]) / (2. * dt);
			G[t0] = (-log(X[t0 + 2]) + 4. * log(X[t0 + 1]) - 3. * log(X[t0])) / (2. * dt);
			D[t0] = -G[t0] / q2;
			H[t0] = D[t0] * F0;
			/* interior lags: central differences while the NEXT sample is still >= 0.1 (:403-415) */
			long c = t0 + 1, nx = c + 1, pv = c - 1;
			for (; c < tN; c++) {
				nx = c + 1; pv = c - 1;
				if (X[nx] < 0.1) break;
				F1[c] = (X[nx] - X[pv]) / (2. * dt);
				G[c] = (log(X[nx]) - log(X[pv])) / (2. * dt);
				D[c] = -G[c] / q2;
				H[c] = D[c] * F0;
			}
			/* where the loop stopped: first the central form with the neighbours it last used (:416-419), then -- same
			 * element -- the backward second-order form (:421-425), which is what stays */
			/* c - 2 lies in front of this row when the very first test failed: the previous wave number's last lag, or -- for
			 * wave number 1 -- the word in front of the reference's table (heap bookkeeping read as a double: a denormal);
			 * 0 stands in for it here, so that one summary entry of such a block is not comparable */
			const double Xm2 = c - 2 >= 0 ? X[c - 2] : 0.0;
			F1[c] = (Xm2 - 4. * X[c - 1] + 3. * X[c]) / (2. * dt);
			G[c] = (log(Xm2) - 4. * log(X[c - 1]) + 3. * log(X[c])) / (2. * dt);
			D[c] = -G[c] / q2;
			H[c] = D[c] * F0;
			const long pick = c < t0 + nValDiff ? c : t0 + nValDiff; /* :427-431 */
			fprintf(out, "%lf %le %le %le %le \n", q, X[pick], F1[pick], G[pick], D[pick]);
		}
		fclose(out);
		for (int k = 0; k < S; k++) { /* keep F, scale the working copy by its lag-0 value (:438-445) */
			const size_t t0 = (size_t)k * T;
			for (int t = 1; t < T; t++) { a->F[j][t0 + t] = X[t0 + t]; X[t0 + t] /= X[t0]; }
			a->F[j][t0] = X[t0];
			X[t0] = 1.;
		}
	}
	a->fourier = 0;
	const double tmax_data = (T - 1) * a->deck_deltaT;
	a->tMax = 100. < tmax_data ? 100. : tmax_data;
	a->nv = (int)(T * a->tMax / (T - 1) / a->deck_deltaT);
	printf("nv = %d, tMax = %f, nCTime = %d, deltaTCorr = %f\n", a->nv, a->tMax, T, a->deck_deltaT); /* :449 */
	return 0;
}

/* FftComplex, analysis.c:558-590: radix-2 decimation in time, twiddles by the rotation recurrence (kept for the digits) */
static void fft_pow2(double *re, double *im, int n)
{
	for (int i = 0, rev = 0; i < n; i++) {
		if (i < rev) { double t = re[i]; re[i] = re[rev]; re[rev] = t; t = im[i]; im[i] = im[rev]; im[rev] = t; }
		int half = n / 2;
		while (half >= 1 && rev >= half) { rev -= half; half /= 2; }
		rev += half;
	}
	for (int span = 1; span < n; span *= 2) {
		const double theta = M_PI / span, dr = cos(theta) - 1., di = sin(theta);
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

int sqt_analysis_text_fourier(sqt_analysis *a, int window)
{
	const int S = a->nCSpatial, T = a->nCTime;
	const int n = 2 * (T - 1);
	if (T < 2 || (n & (n - 1)) != 0) return -1; /* the reference assumes a power of two */
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
	a->omegaMax = (int)(10 < nyq ? 10 : nyq); /* an int in the reference (:131, :366) */
	a->nv = (int)(T * a->omegaMax / nyq);
	return 0;
}

/* Print_R2_data, analysis.c:508-536 -- including what it echoes on stdout */
static int print_table(const sqt_analysis *a, const char *outdir, const char *stem, const char *block, const double *v)
{
	char path[4096];
	snprintf(path, sizeof(path), "%s%s%s.%s.info", outdir ? outdir : "", outdir ? "/" : "", stem, block);
	FILE *fp = fopen(path, "w");
	if (!fp) return -1;
	const int S = a->nCSpatial, T = a->nCTime;
	fprintf(fp, "#%d", S);
	for (int k = 0; k < S; k++) {
		const double x = (k + 1) * a->kVal;
		printf("### kVal = %d * %f= %f\n", k + 1, a->kVal, x);
		fprintf(fp, " %9.4e", x);
	}
	fprintf(fp, "\n");
	for (int t = 0; t < a->nv; t++) {
		double x;
		if (a->fourier) x = (double)(t * a->omegaMax / a->nv); /* integer arithmetic in the reference (:522) */
		else x = t * a->deck_deltaT;
		printf("%9.4e", x);
		fprintf(fp, "%9.4e", x);
		for (int k = 0; k < S; k++) {
			printf(" %9.4e", v[(size_t)k * T + t]);
			fprintf(fp, " %9.4e", v[(size_t)k * T + t]);
		}
		fprintf(fp, "\n");
	}
	fclose(fp);
	return 0;
}

int sqt_analysis_text_print(const sqt_analysis *a, const char *outdir)
{
	int rc = 0;
	for (int j = 0; j < SQT_AN_TYPES; j++) { /* :464-505 */
		rc |= print_table(a, outdir, "Fqt", k_block[j], a->F[j]);
		if (!k_more[j]) continue;
		rc |= print_table(a, outdir, "GammaQT", k_block[j], a->Gamma[j]);
		rc |= print_table(a, outdir, "Dqt", k_block[j], a->D[j]);
		rc |= print_table(a, outdir, "D0Hqt", k_block[j], a->H[j]);
		rc |= print_table(a, outdir, "Fqt1", k_block[j], a->F1[j]);
	}
	return rc ? -1 : 0;
}
