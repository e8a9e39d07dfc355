/*
 * sqt_dump.c -- LAMMPS text-dump reader straight into SoA staging (the pinned buffers of
 * sqt_stage_acquire), replacing the reference's AoS Snapshot/Atom path
 * (ReadDumpForOnlyCheck snapshot.c:25-165, ReadDump snapshot.c:166-382).
 *
 * Format authority is the reference reader: nine header lines, the four accepted
 * "ITEM: BOX BOUNDS" variants, the four accepted "ITEM: ATOMS id type xu yu zu ..." column
 * sets (matched by prefix, longest first as in snapshot.c:238-350), fields separated by single
 * spaces (snapshot.c:22), numbers converted with atoi/atof semantics, atoms taken in file-line
 * order with `id` ignored.  Only x y z are kept: the hot path never reads type, velocity or
 * dipole columns.
 */
#include "../../../include/sqt_host.h"

#include <stdlib.h>
#include <string.h>

#define LINE_MAX_LEN 1024 /* MAXLINE, snapshot.h:27 */

static int starts_with(const char *line, const char *prefix) { return strncmp(prefix, line, strlen(prefix)) == 0; }

/* atof on the next space-delimited token; *ok cleared if the token is missing
 * (strtok_null_safe, snapshot.c:391-398 returns "0" and raises error_code). */
static double next_real(char **cursor, int *ok)
{
	char *s = *cursor;
	while (*s == ' ') s++;
	if (*s == '\0' || *s == '\n') { *ok = 0; return 0.0; }
	char *end = s;
	while (*end && *end != ' ') end++;
	/* decimal fast path (Clinger): <= 15 significant digits and |exp10| <= 22 is exactly one
	 * correctly rounded operation, i.e. identical to strtod; everything else goes to strtod. */
	{
		const char *q = s;
		int neg = 0, nd = 0, scale = 0, seen_dot = 0, simple = 1, any = 0;
		unsigned long long mant = 0;
		if (*q == '-') { neg = 1; q++; } else if (*q == '+') q++;
		for (; q < end; q++) {
			if (*q >= '0' && *q <= '9') {
				any = 1;
				if (mant || *q != '0') { if (++nd > 15) { simple = 0; break; } }
				mant = mant * 10 + (unsigned)(*q - '0');
				if (seen_dot) scale++;
			} else if (*q == '.' && !seen_dot) seen_dot = 1;
			else if (*q == '\n' || *q == '\r') break;
			else { simple = 0; break; }
		}
		if (simple && any && scale <= 22) {
			static const double p10[23] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
				                            1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22 };
			double v = (double)mant / p10[scale];
			*cursor = end;
			return neg ? -v : v;
		}
	}
	double v = strtod(s, NULL);
	*cursor = end;
	return v;
}

static int skip_token(char **cursor)
{
	char *s = *cursor;
	while (*s == ' ') s++;
	if (*s == '\0' || *s == '\n') return 0;
	while (*s && *s != ' ') s++;
	*cursor = s;
	return 1;
}

int sqt_dump_read_header(FILE *fp, sqt_frame_header *hd)
{
	char line[LINE_MAX_LEN];
	memset(hd, 0, sizeof(*hd));
	if (!fgets(line, LINE_MAX_LEN, fp) || !starts_with(line, "ITEM: TIMESTEP")) return 0;
	if (!fgets(line, LINE_MAX_LEN, fp)) return 0;
	hd->timestep = atol(line);
	if (!fgets(line, LINE_MAX_LEN, fp) || !starts_with(line, "ITEM: NUMBER OF ATOMS")) return 0;
	if (!fgets(line, LINE_MAX_LEN, fp)) return 0;
	hd->n_atoms = atoi(line);
	if (!fgets(line, LINE_MAX_LEN, fp)) return 0;
	/* the reference compares only the first strlen("ITEM: BOX BOUNDS pp pp pp") characters */
	if (starts_with(line, "ITEM: BOX BOUNDS pp pp pp")) hd->pbcType = 7;
	else if (starts_with(line, "ITEM: BOX BOUNDS pp pp ff")) hd->pbcType = 3;
	else if (starts_with(line, "ITEM: BOX BOUNDS ff ff pp")) hd->pbcType = 4;
	else if (starts_with(line, "ITEM: BOX BOUNDS ff ff ff")) hd->pbcType = 0;
	else return 0;
	double *lo[3] = { &hd->xlo, &hd->ylo, &hd->zlo }, *hi[3] = { &hd->xhi, &hd->yhi, &hd->zhi };
	for (int a = 0; a < 3; a++) {
		if (!fgets(line, LINE_MAX_LEN, fp)) return 0;
		char *c = line;
		int ok = 1;
		*lo[a] = next_real(&c, &ok);
		*hi[a] = next_real(&c, &ok);
	}
	if (!fgets(line, LINE_MAX_LEN, fp)) return 0;
	if (starts_with(line, "ITEM: ATOMS id type xu yu zu mux muy muz vx vy vz")) hd->style = 3;
	else if (starts_with(line, "ITEM: ATOMS id type xu yu zu mux muy muz")) hd->style = 2;
	else if (starts_with(line, "ITEM: ATOMS id type xu yu zu vx vy vz")) hd->style = 1;
	else if (starts_with(line, "ITEM: ATOMS id type xu yu zu")) hd->style = 0;
	else return 0;
	return 1;
}

int sqt_dump_skip_atoms(FILE *fp, const sqt_frame_header *hd)
{
	char line[LINE_MAX_LEN];
	for (int i = 0; i < hd->n_atoms; i++)
		if (!fgets(line, LINE_MAX_LEN, fp)) return 0;
	return 1;
}

int sqt_dump_read_atoms(FILE *fp, const sqt_frame_header *hd, double *x, double *y, double *z)
{
	char line[LINE_MAX_LEN];
	int ok = 1;
	for (int i = 0; i < hd->n_atoms; i++) {
		if (!fgets(line, LINE_MAX_LEN, fp)) return 0;
		/* a short line raises the reference's error_code, which fails the frame when the NEXT atom
		 * line is read (snapshot.c:352-356) -- so a short LAST line is accepted with zeros */
		if (!ok) return 0;
		char *c = line;
		if (!skip_token(&c) || !skip_token(&c)) ok = 0; /* id, type: parsed by the reference, never used */
		x[i] = next_real(&c, &ok);
		y[i] = next_real(&c, &ok);
		z[i] = next_real(&c, &ok);
		/* the wider styles' columns (vx vy vz / mux muy muz) are not used, but the reference parses them: a line that lacks
		 * them is a short line (snapshot.c:268-330).  Three more tokens in every wider style: the 11-column header also
		 * matches the reference's 8-column "mux muy muz" prefix test, which comes first (snapshot.c:262) */
		for (int k = hd->style ? 3 : 0; k > 0; k--)
			if (!skip_token(&c)) ok = 0;
	}
	return 1;
}
