/*
 * sqt_reader.c -- high-rate LAMMPS text-dump reader: the file is memory-mapped, a frame's atom block is
 * located by counting newlines (memchr speed) and its lines are parsed by several threads straight into
 * the SoA staging buffer (the pinned memory of sqt_stage_acquire).
 *
 * Same format authority and the same conversions as sqt_dump.c / the reference reader (snapshot.c:166-382):
 * header line prefixes, single-space separated fields, `id type` skipped, x y z with atof semantics (decimal
 * fast path = correctly rounded, strtod otherwise), atoms in file-line order.  The reference parses about
 * 1 M atom lines per second per process (strtok + atof); at config-2 shape (32 000 atoms) the GPU finishes a
 * frame in 0.2 ms, so the reader, not the kernels, sets the drop-in's wall clock -- hence this file.
 *
 * Deviation: the reference cuts lines at MAXLINE = 1024 characters (fgets); longer lines are taken whole here.
 */
#include "../../../include/sqt_host.h"

#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct sqt_reader {
	const char *base;
	size_t size, pos;
	int fd, n_threads;
	int raw;          /* 1: SQTRAW01 binary frame file (see sqt_host.h), 0: LAMMPS text dump */
};

/* Raw binary frame file (extension, SURVEY.md section 8f.1 "raw f64[3][N] frame file"):
 *   file   := "SQTRAW01" frame*
 *   frame  := int64 timestep, int64 n_atoms, f64 xlo xhi ylo yhi zlo zhi, f64 x[n], f64 y[n], f64 z[n]
 * little endian, no padding.  Atom order is the dump order, exactly what the text path would deliver. */
#define RAW_MAGIC "SQTRAW01"
#define RAW_HDR (8 + 8 + 6 * 8)

static int starts_with(const char *p, const char *end, const char *prefix)
{
	size_t n = strlen(prefix);
	return (size_t)(end - p) >= n && memcmp(p, prefix, n) == 0;
}

/* one line [p, *next) without its newline; returns 0 at end of data */
static int take_line(const sqt_reader *r, size_t *pos, const char **p, const char **e)
{
	if (*pos >= r->size) return 0;
	const char *s = r->base + *pos;
	const char *nl = (const char *)memchr(s, '\n', r->size - *pos);
	const char *end = nl ? nl : r->base + r->size;
	*p = s; *e = end;
	*pos = (size_t)(end - r->base) + (nl ? 1 : 0);
	return 1;
}

/* atof on the token at *c (skipping leading spaces), bounded by line end e */
static double tok_real(const char **c, const char *e, int *ok)
{
	const char *s = *c;
	while (s < e && *s == ' ') s++;
	if (s >= e) { *ok = 0; *c = s; return 0.0; }
	const char *t = s;
	while (t < e && *t != ' ') t++;
	{ /* Clinger fast path: <= 15 significant digits, |exp10| <= 22 => one correctly rounded division */
		const char *q = s;
		int neg = 0, nd = 0, scale = 0, dot = 0, simple = 1, any = 0;
		unsigned long long mant = 0;
		if (*q == '-') { neg = 1; q++; } else if (*q == '+') q++;
		for (; q < t; q++) {
			const unsigned d = (unsigned)(*q - '0');
			if (d <= 9) {
				any = 1;
				if (mant || d) { if (++nd > 15) { simple = 0; break; } }
				mant = mant * 10 + d;
				scale += dot;
			} else if (*q == '.' && !dot) dot = 1;
			else if (*q == '\r') break;
			else { simple = 0; break; }
		}
		if (simple && any && scale <= 22) {
			static const double p10[23] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
				                            1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22 };
			const double v = (double)mant / p10[scale];
			*c = t;
			return neg ? -v : v;
		}
	}
	{ /* rare: exponents, long mantissas, inf/nan -- strtod on a NUL-terminated copy */
		char buf[64];
		size_t n = (size_t)(t - s);
		if (n >= sizeof(buf)) n = sizeof(buf) - 1;
		memcpy(buf, s, n);
		buf[n] = 0;
		*c = t;
		return strtod(buf, NULL);
	}
}

static long parse_long(const char *p, const char *e)
{
	long v = 0;
	int neg = 0;
	while (p < e && (*p == ' ' || *p == '\t')) p++;
	if (p < e && (*p == '-' || *p == '+')) { neg = (*p == '-'); p++; }
	while (p < e && *p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
	return neg ? -v : v;
}

static int tok_skip(const char **c, const char *e)
{
	const char *s = *c;
	while (s < e && *s == ' ') s++;
	if (s >= e) return 0;
	while (s < e && *s != ' ') s++;
	*c = s;
	return 1;
}

sqt_reader *sqt_reader_open(const char *path, int n_threads)
{
	int fd = open(path, O_RDONLY);
	if (fd < 0) return NULL;
	struct stat st;
	if (fstat(fd, &st) != 0) { close(fd); return NULL; }
	sqt_reader *r = (sqt_reader *)calloc(1, sizeof(*r));
	r->fd = fd;
	r->size = (size_t)st.st_size;
	if (r->size) {
		void *m = mmap(NULL, r->size, PROT_READ, MAP_PRIVATE, fd, 0);
		if (m == MAP_FAILED) { close(fd); free(r); return NULL; }
		madvise(m, r->size, MADV_SEQUENTIAL);
		r->base = (const char *)m;
	}
	if (n_threads <= 0) {
		const char *env = getenv("SQT_READER_THREADS");
		n_threads = env && *env ? atoi(env) : 0;
	}
#ifdef _OPENMP
	if (n_threads <= 0) { n_threads = omp_get_num_procs(); if (n_threads > 16) n_threads = 16; }
#else
	n_threads = 1;
#endif
	r->n_threads = n_threads < 1 ? 1 : n_threads;
	if (r->size >= 8 && memcmp(r->base, RAW_MAGIC, 8) == 0) { r->raw = 1; r->pos = 8; }
	return r;
}

void sqt_reader_close(sqt_reader *r)
{
	if (!r) return;
	if (r->base) munmap((void *)r->base, r->size);
	close(r->fd);
	free(r);
}

void sqt_reader_rewind(sqt_reader *r) { if (r) r->pos = r->raw ? 8 : 0; }
int sqt_reader_is_raw(const sqt_reader *r) { return r ? r->raw : 0; }

int sqt_reader_next_header(sqt_reader *r, sqt_frame_header *hd)
{
	const char *p, *e;
	size_t pos = r->pos;
	memset(hd, 0, sizeof(*hd));
	if (r->raw) {
		if (r->size - pos < RAW_HDR) return 0;
		long long ts, n;
		double box[6];
		memcpy(&ts, r->base + pos, 8); memcpy(&n, r->base + pos + 8, 8); memcpy(box, r->base + pos + 16, 48);
		if (n <= 0 || n > 2147483647LL || (r->size - pos - RAW_HDR) / 24 < (size_t)n) return 0;
		hd->timestep = (long)ts; hd->n_atoms = (int)n; hd->pbcType = 7;
		hd->xlo = box[0]; hd->xhi = box[1]; hd->ylo = box[2]; hd->yhi = box[3]; hd->zlo = box[4]; hd->zhi = box[5];
		r->pos = pos + RAW_HDR;
		return 1;
	}
	if (!take_line(r, &pos, &p, &e) || !starts_with(p, e, "ITEM: TIMESTEP")) return 0;
	if (!take_line(r, &pos, &p, &e)) return 0;
	hd->timestep = parse_long(p, e);
	if (!take_line(r, &pos, &p, &e) || !starts_with(p, e, "ITEM: NUMBER OF ATOMS")) return 0;
	if (!take_line(r, &pos, &p, &e)) return 0;
	hd->n_atoms = (int)parse_long(p, e);
	if (!take_line(r, &pos, &p, &e)) return 0;
	if (starts_with(p, e, "ITEM: BOX BOUNDS pp pp pp")) hd->pbcType = 7;
	else if (starts_with(p, e, "ITEM: BOX BOUNDS pp pp ff")) hd->pbcType = 3;
	else if (starts_with(p, e, "ITEM: BOX BOUNDS ff ff pp")) hd->pbcType = 4;
	else if (starts_with(p, e, "ITEM: BOX BOUNDS ff ff ff")) hd->pbcType = 0;
	else return 0;
	double *lo[3] = { &hd->xlo, &hd->ylo, &hd->zlo }, *hi[3] = { &hd->xhi, &hd->yhi, &hd->zhi };
	for (int a = 0; a < 3; a++) {
		if (!take_line(r, &pos, &p, &e)) return 0;
		int ok = 1;
		const char *c = p;
		*lo[a] = tok_real(&c, e, &ok);
		*hi[a] = tok_real(&c, e, &ok);
	}
	if (!take_line(r, &pos, &p, &e)) return 0;
	if (starts_with(p, e, "ITEM: ATOMS id type xu yu zu mux muy muz vx vy vz")) hd->style = 3;
	else if (starts_with(p, e, "ITEM: ATOMS id type xu yu zu mux muy muz")) hd->style = 2;
	else if (starts_with(p, e, "ITEM: ATOMS id type xu yu zu vx vy vz")) hd->style = 1;
	else if (starts_with(p, e, "ITEM: ATOMS id type xu yu zu")) hd->style = 0;
	else return 0;
	r->pos = pos;
	return 1;
}

/* offsets of the starts of lines 0, step, 2*step, ... and of the end of line n-1; 0 if the data ends early */
static int index_lines(const sqt_reader *r, int n, int step, size_t *starts, int n_starts, size_t *end_pos)
{
	size_t pos = r->pos;
	int line = 0, k = 0;
	while (line < n) {
		if (line % step == 0 && k < n_starts) starts[k++] = pos;
		if (pos >= r->size) return 0;
		const char *nl = (const char *)memchr(r->base + pos, '\n', r->size - pos);
		pos = nl ? (size_t)(nl - r->base) + 1 : r->size;
		line++;
	}
	*end_pos = pos;
	return 1;
}

int sqt_reader_skip_atoms(sqt_reader *r, const sqt_frame_header *hd)
{
	size_t endp, dummy;
	if (hd->n_atoms <= 0) return 1;
	if (r->raw) { r->pos += (size_t)hd->n_atoms * 24; return 1; }
	if (!index_lines(r, hd->n_atoms, hd->n_atoms, &dummy, 1, &endp)) { r->pos = r->size; return 0; }
	r->pos = endp;
	return 1;
}

int sqt_reader_read_atoms(sqt_reader *r, const sqt_frame_header *hd, double *x, double *y, double *z)
{
	const int n = hd->n_atoms;
	if (n <= 0) return 1;
	if (r->raw) { /* already SoA: three straight copies into the staging buffer */
		const char *src = r->base + r->pos;
		memcpy(x, src, (size_t)n * 8); memcpy(y, src + (size_t)n * 8, (size_t)n * 8); memcpy(z, src + (size_t)n * 16, (size_t)n * 8);
		r->pos += (size_t)n * 24;
		return 1;
	}
	int T = r->n_threads;
	if (T > n / 2048 + 1) T = n / 2048 + 1; /* threads only pay off on big blocks */
	const int step = (n + T - 1) / T;
	size_t starts[64], endp;
	if (T > 64) T = 64;
	if (!index_lines(r, n, step, starts, T, &endp)) { r->pos = r->size; return 0; }
	int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(T) schedule(static, 1) reduction(| : bad)
#endif
	for (int t = 0; t < T; t++) {
		const int i0 = t * step, i1 = (i0 + step < n) ? i0 + step : n;
		size_t pos = starts[t];
		for (int i = i0; i < i1; i++) {
			const char *p, *e;
			if (!take_line(r, &pos, &p, &e)) { bad |= 2; break; }
			int ok = 1;
			const char *c = p;
			if (!tok_skip(&c, e) || !tok_skip(&c, e)) ok = 0; /* id, type */
			x[i] = tok_real(&c, e, &ok);
			y[i] = tok_real(&c, e, &ok);
			z[i] = tok_real(&c, e, &ok);
			if (!ok && i != n - 1) bad |= 1; /* a short line fails the frame unless it is the last (snapshot.c:352-356) */
		}
	}
	r->pos = endp;
	return bad ? 0 : 1;
}

int sqt_raw_append_frame(const char *path, int create, long long timestep, long long n_atoms, const double box[6],
                         const double *x, const double *y, const double *z)
{
	FILE *fp = fopen(path, create ? "wb" : "ab");
	if (!fp) return -1;
	int ok = 1;
	if (create) ok &= fwrite(RAW_MAGIC, 1, 8, fp) == 8;
	ok &= fwrite(&timestep, 8, 1, fp) == 1;
	ok &= fwrite(&n_atoms, 8, 1, fp) == 1;
	ok &= fwrite(box, 8, 6, fp) == 6;
	ok &= fwrite(x, 8, (size_t)n_atoms, fp) == (size_t)n_atoms;
	ok &= fwrite(y, 8, (size_t)n_atoms, fp) == (size_t)n_atoms;
	ok &= fwrite(z, 8, (size_t)n_atoms, fp) == (size_t)n_atoms;
	fclose(fp);
	return ok ? 0 : -1;
}
