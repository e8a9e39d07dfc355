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
#include <math.h>
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
	int lines_per_thread; /* least number of atom lines a thread must get before the parser forks (SQT_READER_LINES_PER_THREAD) */
	int raw;          /* 0: LAMMPS text dump, 1: SQTRAW01 binary frame file (see sqt_host.h), 2: LAMMPS native binary dump */
	/* LAMMPS binary: where the current frame's chunks start and how its rows are laid out */
	size_t lb_chunks;
	int lb_size_one, lb_nchunk, lb_col[3];
	/* text dumps: what the count pass (sqt_reader_skip_atoms) learned about each atom block, so that the parse pass does
	 * not scan for newlines again: block start / end and the start of every MARK_LINES-th line */
	struct blk_index { size_t atoms_pos, end_pos; int n_atoms; size_t mark0; } *blk;
	size_t n_blk, cap_blk, blk_cursor;
	size_t *marks;
	size_t n_marks, cap_marks;
};

/* Raw binary frame file (extension, SURVEY.md section 8f.1 "raw f64[3][N] frame file"):
 *   file   := "SQTRAW01" frame*
 *   frame  := int64 timestep, int64 n_atoms, f64 xlo xhi ylo yhi zlo zhi, f64 x[n], f64 y[n], f64 z[n]
 * little endian, no padding.  Atom order is the dump order, exactly what the text path would deliver. */
#define RAW_MAGIC "SQTRAW01"
#define RAW_HDR (8 + 8 + 6 * 8)

/* LAMMPS native binary dump (`dump custom ... file.bin`, what LAMMPS' own tools/binary2txt.cpp reads), little endian:
 *   frame := [ int64 -len, char magic[len] ("DUMPCUSTOM"), int32 endian (1), int32 revision ]      (format since 2020; absent in older files)
 *            int64 ntimestep, int64 natoms, int32 triclinic, int32 boundary[3][2], f64 xlo xhi ylo yhi zlo zhi, [f64 xy xz yz]
 *            int32 size_one,
 *            [ int32 len, char units[len], char time_flag, [f64 time], int32 len, char columns[len] ]   (revision > 1)
 *            int32 nchunk, { int32 n, f64 buf[n] } * nchunk               (n / size_one atom rows per chunk, every field a double)
 * Only what the text path accepts is accepted here: orthogonal box, columns `id type xu yu zu ...` (by name when the file
 * carries names, by position otherwise), atoms in stored order (write the dump with `dump_modify sort id`, as the text
 * path needs too: the reference identifies atoms by line order, snapshot.c:357-368). */
typedef struct { const char *p, *e; int ok; } lb_cur;
static void lb_get(lb_cur *c, void *dst, size_t n)
{
	if (!c->ok || (size_t)(c->e - c->p) < n) { c->ok = 0; memset(dst, 0, n); return; }
	memcpy(dst, c->p, n);
	c->p += n;
}
static void lb_skip(lb_cur *c, size_t n)
{
	if (!c->ok || (size_t)(c->e - c->p) < n) { c->ok = 0; return; }
	c->p += n;
}

static int lb_next_header(sqt_reader *r, sqt_frame_header *hd)
{
	lb_cur c = { r->base + r->pos, r->base + r->size, 1 };
	long long ntimestep = 0, natoms = 0;
	int revision = 0, has_magic = 0;
	lb_get(&c, &ntimestep, 8);
	if (!c.ok) return 0;
	if (ntimestep < 0) {
		const long long mlen = -ntimestep;
		if (mlen > 64) return 0;
		char magic[65] = { 0 };
		int endian = 0;
		lb_get(&c, magic, (size_t)mlen);
		lb_get(&c, &endian, 4);
		lb_get(&c, &revision, 4);
		lb_get(&c, &ntimestep, 8);
		if (!c.ok || endian != 1 || strcmp(magic, "DUMPCUSTOM") != 0) return 0; /* `dump atom` files carry scaled coordinates */
		has_magic = 1;
	}
	int triclinic = 0, boundary[6], size_one = 0, nchunk = 0;
	double box[6];
	lb_get(&c, &natoms, 8);
	lb_get(&c, &triclinic, 4);
	lb_get(&c, boundary, 24);
	lb_get(&c, box, 48);
	if (!c.ok || triclinic || natoms <= 0 || natoms > 2147483647LL) return 0;
	lb_get(&c, &size_one, 4);
	int col[3] = { 2, 3, 4 };
	if (has_magic && revision > 1) {
		int len = 0;
		char flag = 0;
		lb_get(&c, &len, 4);
		if (len < 0) return 0;
		lb_skip(&c, (size_t)len);                 /* unit style */
		lb_get(&c, &flag, 1);
		if (flag) lb_skip(&c, 8);                 /* simulation time */
		lb_get(&c, &len, 4);
		if (!c.ok || len < 0 || (size_t)(c.e - c.p) < (size_t)len) return 0;
		/* column names: the reference's header test is the prefix "id type xu yu zu" (snapshot.c:296-331) */
		static const char want[] = "id type xu yu zu";
		if ((size_t)len < sizeof(want) - 1 || memcmp(c.p, want, sizeof(want) - 1) != 0) return 0;
		lb_skip(&c, (size_t)len);
	}
	lb_get(&c, &nchunk, 4);
	if (!c.ok || size_one < 5 || nchunk < 0) return 0;
	memset(hd, 0, sizeof(*hd));
	hd->timestep = (long)ntimestep; hd->n_atoms = (int)natoms;
	hd->xlo = box[0]; hd->xhi = box[1]; hd->ylo = box[2]; hd->yhi = box[3]; hd->zlo = box[4]; hd->zhi = box[5];
	hd->pbcType = (boundary[0] == 0 && boundary[1] == 0 ? 1 : 0) | (boundary[2] == 0 && boundary[3] == 0 ? 2 : 0) | (boundary[4] == 0 && boundary[5] == 0 ? 4 : 0);
	hd->style = size_one >= 11 ? 3 : (size_one >= 8 ? 1 : 0);
	r->lb_chunks = (size_t)(c.p - r->base);
	r->lb_size_one = size_one; r->lb_nchunk = nchunk;
	r->lb_col[0] = col[0]; r->lb_col[1] = col[1]; r->lb_col[2] = col[2];
	r->pos = r->lb_chunks;
	return 1;
}

/* walks the frame's chunks; with x != NULL scatters the coordinate columns into SoA.  1 ok / 0 short or inconsistent file */
static int lb_walk_atoms(sqt_reader *r, const sqt_frame_header *hd, double *x, double *y, double *z, double *vx, double *vy, double *vz)
{
	lb_cur c = { r->base + r->lb_chunks, r->base + r->size, 1 };
	const int so = r->lb_size_one, cx = r->lb_col[0], cy = r->lb_col[1], cz = r->lb_col[2];
	long long row0 = 0;
	for (int k = 0; k < r->lb_nchunk; k++) {
		int n = 0;
		lb_get(&c, &n, 4);
		if (!c.ok || n < 0 || n % so != 0 || (size_t)(c.e - c.p) / 8 < (size_t)n) { r->pos = r->size; return 0; }
		const long long rows = n / so;
		if (row0 + rows > hd->n_atoms) { r->pos = r->size; return 0; }
		if (x) {
			const char *src = c.p;
#ifdef _OPENMP
#pragma omp parallel for num_threads(r->n_threads) schedule(static) if (rows > 65536)
#endif
			for (long long i = 0; i < rows; i++) {
				const char *row = src + (size_t)i * so * 8;
				memcpy(&x[row0 + i], row + 8 * cx, 8);
				memcpy(&y[row0 + i], row + 8 * cy, 8);
				memcpy(&z[row0 + i], row + 8 * cz, 8);
				if (vx) { /* `id type xu yu zu vx vy vz`: the three fields after zu */
					memcpy(&vx[row0 + i], row + 8 * (cz + 1), 8);
					memcpy(&vy[row0 + i], row + 8 * (cz + 2), 8);
					memcpy(&vz[row0 + i], row + 8 * (cz + 3), 8);
				}
			}
		}
		lb_skip(&c, (size_t)n * 8);
		row0 += rows;
	}
	r->pos = (size_t)(c.p - r->base);
	return row0 == hd->n_atoms;
}

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

/* Common-case line "<id> <type> <x> <y> <z>[ ...]\n": one forward scan, no token pre-scan, no bounds checks (the caller
 * guarantees >= FAST_MARGIN readable bytes after p).  Numbers go through the same correctly rounded fast path as
 * tok_real (<= 15 digits, one division).  Returns the start of the next line, or NULL when anything is unusual
 * (exponent, long mantissa, tab, short line, ...): the caller then re-parses that line with the general code, so the
 * semantics are those of the general code in every case. */
#define FAST_MARGIN 512
#ifndef SQT_LINES_PER_THREAD
#define SQT_LINES_PER_THREAD 131072
#endif
#define FAST_SCAN (FAST_MARGIN - 160) /* no loop below runs past line start + FAST_SCAN; the final memchr looks at most 128 further */
static inline const char *fast_real(const char *p, const char *lim, double *out)
{
	static const double p10[19] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18 };
	int neg = 0;
	if (*p == '-') { neg = 1; p++; } else if (*p == '+') p++;
	const char *d0 = p;
	unsigned long long mant = 0;
	unsigned d;
	while ((d = (unsigned)(*p - '0')) <= 9 && p < lim) { mant = mant * 10 + d; p++; }
	int nd = (int)(p - d0), scale = 0;
	if (*p == '.') {
		const char *f0 = ++p;
		while ((d = (unsigned)(*p - '0')) <= 9 && p < lim) { mant = mant * 10 + d; p++; }
		scale = (int)(p - f0);
		nd += scale;
	}
	if (nd == 0 || nd > 15 || (*p != ' ' && *p != '\n' && *p != '\r')) return NULL; /* leading zeros count here: conservative */
	const double v = (double)mant / p10[scale];
	*out = neg ? -v : v;
	return p;
}

static inline const char *fast_line(const char *p, const int extra, double *x, double *y, double *z, double *v3, const int nv)
{
	const char *lim = p + FAST_SCAN;
#define SKIP_SPACES while (*p == ' ' && p < lim) p++
#define SKIP_DIGITS while ((unsigned)(*p - '0') <= 9 && p < lim) p++
	SKIP_SPACES;
	if ((unsigned)(*p - '0') > 9) return NULL;
	SKIP_DIGITS;                                        /* id */
	if (*p != ' ') return NULL;
	SKIP_SPACES;
	if ((unsigned)(*p - '0') > 9) return NULL;
	SKIP_DIGITS;                                        /* type */
	if (*p != ' ') return NULL;
	SKIP_SPACES;
	if (!(p = fast_real(p, lim, x)) || *p != ' ') return NULL;
	SKIP_SPACES;
	if (!(p = fast_real(p, lim, y)) || *p != ' ') return NULL;
	SKIP_SPACES;
	if (!(p = fast_real(p, lim, z))) return NULL;
#undef SKIP_SPACES
#undef SKIP_DIGITS
	if (extra == 0) {
		if (*p == '\n') return p + 1;
		if (p >= lim) return NULL;
		const char *nl = (const char *)memchr(p, '\n', 128); /* trailing blanks or \r */
		return nl ? nl + 1 : NULL;
	}
	if (v3) { /* the nv (3 or 6) columns after zu are wanted: vx vy vz (snapshot.c:285-300), mux muy muz (:263-282), or both */
		const char *q = p;
		for (int k = 0; k < nv; k++) {
			if (*q != ' ') return NULL;
			while (*q == ' ' && q < lim) q++;
			if (!(q = fast_real(q, lim, &v3[k]))) return NULL;
		}
		if (*q == '\n') return q + 1;
		if (q >= lim) return NULL;
		const char *nl = (const char *)memchr(q, '\n', 128);
		return nl ? nl + 1 : NULL;
	}
	/* wider styles (vx vy vz / mux muy muz): the columns are not used, but the reference parses them -- a line that lacks
	 * them is a short line (snapshot.c:268-330), which only the general path judges */
	int tokens = 0;
	const char *end = lim + 128;
	for (;;) {
		while (*p == ' ' && p < end) p++;
		if (*p == '\n' || p >= end) break;
		tokens++;
		while (*p != ' ' && *p != '\n' && p < end) p++;
	}
	if (*p != '\n' || tokens < extra) return NULL;
	return p + 1;
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
	{
		const char *env = getenv("SQT_READER_LINES_PER_THREAD");
		r->lines_per_thread = env && *env ? atoi(env) : SQT_LINES_PER_THREAD;
		if (r->lines_per_thread < 1) r->lines_per_thread = 1;
	}
	if (r->size >= 8 && memcmp(r->base, RAW_MAGIC, 8) == 0) { r->raw = 1; r->pos = 8; }
	else if (r->size >= 8 && memcmp(r->base, "ITEM:", 5) != 0) r->raw = 2; /* neither text nor SQTRAW01: LAMMPS native binary */
	return r;
}

void sqt_reader_close(sqt_reader *r)
{
	if (!r) return;
	free(r->blk);
	free(r->marks);
	if (r->base) munmap((void *)r->base, r->size);
	close(r->fd);
	free(r);
}

void sqt_reader_rewind(sqt_reader *r) { if (r) { r->pos = r->raw == 1 ? 8 : 0; r->blk_cursor = 0; } }
int sqt_reader_is_raw(const sqt_reader *r) { return r ? r->raw : 0; }

int sqt_reader_next_header(sqt_reader *r, sqt_frame_header *hd)
{
	const char *p, *e;
	size_t pos = r->pos;
	memset(hd, 0, sizeof(*hd));
	if (r->raw == 2) return lb_next_header(r, hd);
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

#define MARK_LINES 1024
#define SCAN_BLOCK 4096

#if defined(__SSE2__)
#include <emmintrin.h>
static size_t count_nl(const char *p, size_t n)
{
	size_t c = 0, i = 0;
	const __m128i nl = _mm_set1_epi8('\n'), zero = _mm_setzero_si128();
	while (i + 16 <= n) { /* byte-wise match counters (at most 255 rounds before they could wrap), then one horizontal sum */
		__m128i acc = zero;
		for (int k = 0; k < 255 && i + 16 <= n; k++, i += 16)
			acc = _mm_sub_epi8(acc, _mm_cmpeq_epi8(_mm_loadu_si128((const __m128i *)(p + i)), nl));
		const __m128i sum = _mm_sad_epu8(acc, zero);
		c += (size_t)_mm_cvtsi128_si64(sum) + (size_t)_mm_cvtsi128_si64(_mm_srli_si128(sum, 8));
	}
	for (; i < n; i++) c += (p[i] == '\n');
	return c;
}
#else
static size_t count_nl(const char *p, size_t n)
{
	size_t c = 0;
	for (size_t i = 0; i < n; i++) c += (p[i] == '\n');
	return c;
}
#endif

/* Newline scan of an n-line atom block in 4 KB steps (vector count; memchr only inside the steps that hold a wanted line
 * start).  Records the start of every MARK_LINES-th line (optional) and the end of line n-1.  Same acceptance rule as
 * index_lines: the last line may end with the file instead of a newline.  0 if the data ends early. */
static int scan_block_lines(const sqt_reader *r, int n, size_t *marks, size_t *end_pos)
{
	size_t pos = r->pos, line = 0, next_mark = 0, k = 0;
	while (line < (size_t)n) {
		if (line == next_mark) { if (marks) marks[k] = pos; k++; next_mark += MARK_LINES; }
		if (pos >= r->size) return 0;
		const size_t target = next_mark < (size_t)n ? next_mark : (size_t)n; /* the next line start that is wanted */
		const size_t blk = r->size - pos < SCAN_BLOCK ? r->size - pos : SCAN_BLOCK;
		const size_t c = count_nl(r->base + pos, blk);
		if (line + c < target) {
			line += c; pos += blk;
			if (pos >= r->size) { /* data exhausted: fine only if just the unterminated last line is left */
				if (line == (size_t)n - 1 && r->base[r->size - 1] != '\n') { *end_pos = r->size; return 1; }
				return 0;
			}
			continue;
		}
		size_t q = pos;
		while (line < target) { /* the wanted start lies in this step: walk to it */
			const char *nl = (const char *)memchr(r->base + q, '\n', pos + blk - q);
			q = (size_t)(nl - r->base) + 1;
			line++;
		}
		pos = q;
	}
	*end_pos = pos;
	return 1;
}

int sqt_reader_skip_atoms(sqt_reader *r, const sqt_frame_header *hd)
{
	size_t endp;
	if (hd->n_atoms <= 0) return 1;
	if (r->raw == 2) return lb_walk_atoms(r, hd, NULL, NULL, NULL, NULL, NULL, NULL);
	if (r->raw) { r->pos += (size_t)hd->n_atoms * 24; return 1; }
	const int n = hd->n_atoms;
	const size_t nm = ((size_t)n + MARK_LINES - 1) / MARK_LINES;
	if (r->n_marks + nm > r->cap_marks) {
		size_t cap = r->cap_marks ? r->cap_marks * 2 : 4096;
		while (cap < r->n_marks + nm) cap *= 2;
		size_t *m = (size_t *)realloc(r->marks, cap * sizeof(size_t));
		if (m) { r->marks = m; r->cap_marks = cap; }
	}
	if (r->n_blk == r->cap_blk) {
		size_t cap = r->cap_blk ? r->cap_blk * 2 : 1024;
		struct blk_index *b = (struct blk_index *)realloc(r->blk, cap * sizeof(*b));
		if (b) { r->blk = b; r->cap_blk = cap; }
	}
	const int can_index = r->n_marks + nm <= r->cap_marks && r->n_blk < r->cap_blk;
	if (!scan_block_lines(r, n, can_index ? r->marks + r->n_marks : NULL, &endp)) { r->pos = r->size; return 0; }
	if (can_index) {
		struct blk_index *b = &r->blk[r->n_blk++];
		b->atoms_pos = r->pos; b->end_pos = endp; b->n_atoms = n; b->mark0 = r->n_marks;
		r->n_marks += nm;
	}
	r->pos = endp;
	return 1;
}

/* the `type` column (atoi of the second token, snapshot.c:362) of the atom block at the read position, which stays where it is */
int sqt_reader_peek_types(sqt_reader *r, const sqt_frame_header *hd, int *types)
{
	if (r->raw != 0 || hd->n_atoms <= 0) return 0; /* text dumps only: the binary frame files carry no species */
	size_t pos = r->pos;
	for (int i = 0; i < hd->n_atoms; i++) {
		const char *p, *e;
		if (!take_line(r, &pos, &p, &e)) return 0;
		const char *c = p;
		if (!tok_skip(&c, e)) return 0;
		while (c < e && *c == ' ') c++;
		if (c >= e) return 0;
		types[i] = atoi(c);
	}
	return 1;
}

int sqt_reader_read_atoms(sqt_reader *r, const sqt_frame_header *hd, double *x, double *y, double *z)
{
	return sqt_reader_read_atoms_v(r, hd, x, y, z, NULL, NULL, NULL);
}

int sqt_reader_read_atoms_v(sqt_reader *r, const sqt_frame_header *hd, double *x, double *y, double *z, double *vx, double *vy, double *vz)
{
	return sqt_reader_read_atoms_vm(r, hd, x, y, z, vx, vy, vz, NULL, NULL, NULL);
}

int sqt_reader_read_atoms_vm(sqt_reader *r, const sqt_frame_header *hd, double *x, double *y, double *z, double *vx, double *vy, double *vz,
                             double *mux, double *muy, double *muz)
{
	const int n = hd->n_atoms;
	if (n <= 0) return 1;
	/* velocities: the 8-column `... vx vy vz` style.  The reference's test for the 11-column header is shadowed by its
	 * 8-column dipole prefix test (snapshot.c:262-283), so there it never reads the velocities of such a dump; here the
	 * 11-column style yields them too when BOTH mu and v are asked for (an extension: the reference's -V -M / -s -M runs on
	 * files correlate uninitialised memory).  The raw frame file carries coordinates only. */
	if (vx && !((hd->style == 1 && !mux) || (hd->style == 3 && mux)) ) { r->pos = r->size; return 0; }
	if (vx && r->raw == 1) { r->pos = r->size; return 0; }
	/* dipoles: the `... mux muy muz [vx vy vz]` styles of a text dump */
	if (mux && (r->raw != 0 || (hd->style != 2 && hd->style != 3))) { r->pos = r->size; return 0; }
	const int nv = (vx ? 3 : 0) + (mux ? 3 : 0);
	if (r->raw == 2) return lb_walk_atoms(r, hd, x, y, z, vx, vy, vz);
	if (r->raw) { /* already SoA: straight copies into the staging buffer, in 1 MiB pieces shared by the reader's threads (one
	               * thread moves about 6 GB/s out of the page cache: 4 ms per million atoms, which set the pace at 8 GPUs) */
		const char *src = r->base + r->pos;
		double *dst[3] = { x, y, z };
		const size_t bytes = (size_t)n * 8, piece = (size_t)1 << 20, np = (bytes + piece - 1) / piece;
		int T = r->n_threads;
		if ((size_t)T > 3 * np) T = (int)(3 * np);
		if (T < 1 || n < 262144) T = 1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(T) schedule(static) if (T > 1)
#endif
		for (long long k = 0; k < (long long)(3 * np); k++) {
			const int a = (int)(k / (long long)np);
			const size_t off = (size_t)(k % (long long)np) * piece, len = off + piece <= bytes ? piece : bytes - off;
			memcpy((char *)dst[a] + off, src + (size_t)a * bytes + off, len);
		}
		r->pos += (size_t)n * 24;
		return 1;
	}
	/* Threads only pay off on big blocks: a parallel region costs a wake-up of the workers, which sleep between frames
	 * while the caller drives the GPU -- measured at 2 ms and more per region in CPU-quota'd containers (and far worse when
	 * the OpenMP runtime spins instead), against 1 ms to parse 32 000 lines on one thread.  At least SQT_LINES_PER_THREAD
	 * lines per thread, so a frame below that is parsed by the calling thread alone. */
	int T = r->n_threads;
	if (T > n / r->lines_per_thread) T = n / r->lines_per_thread;
	if (T < 1) T = 1;
	const int step = (n + T - 1) / T;
	size_t starts[64], endp;
	if (T > 64) T = 64;
	const struct blk_index *bi = NULL; /* indexed by the count pass? (blocks are visited in file order) */
	while (r->blk_cursor < r->n_blk && r->blk[r->blk_cursor].atoms_pos < r->pos) r->blk_cursor++;
	if (r->blk_cursor < r->n_blk && r->blk[r->blk_cursor].atoms_pos == r->pos && r->blk[r->blk_cursor].n_atoms == n) bi = &r->blk[r->blk_cursor];
	if (bi) endp = bi->end_pos;
	else if (!index_lines(r, n, step, starts, T, &endp)) { r->pos = r->size; return 0; }
	const int extra = (hd->style == 3 && vx) ? 6 : (hd->style ? 3 : 0); /* 3 also for the 11-column style: the reference's 8-column prefix test matches it first (snapshot.c:262) */
	int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(T) schedule(static, 1) reduction(| : bad) if (T > 1)
#endif
	for (int t = 0; t < T; t++) {
		const int i0 = t * step, i1 = (i0 + step < n) ? i0 + step : n;
		if (i0 >= i1) continue;
		size_t pos;
		if (bi) { /* nearest recorded line start, then forward over at most MARK_LINES - 1 lines */
			pos = r->marks[bi->mark0 + (size_t)i0 / MARK_LINES];
			for (int skip = i0 % MARK_LINES; skip > 0; skip--) {
				const char *nl = (const char *)memchr(r->base + pos, '\n', r->size - pos);
				pos = nl ? (size_t)(nl - r->base) + 1 : r->size;
			}
		} else pos = starts[t];
		for (int i = i0; i < i1; i++) {
			if (r->size - pos > FAST_MARGIN) { /* far enough from the end of the mapping for the unchecked scan */
				double v3[6];
				const char *nx = fast_line(r->base + pos, extra, &x[i], &y[i], &z[i], nv ? v3 : NULL, nv);
				if (nx) {
					if (mux) { mux[i] = v3[0]; muy[i] = v3[1]; muz[i] = v3[2]; }
					if (vx) { const double *q = v3 + (mux ? 3 : 0); vx[i] = q[0]; vy[i] = q[1]; vz[i] = q[2]; }
					pos = (size_t)(nx - r->base);
					continue;
				}
			}
			const char *p, *e;
			if (!take_line(r, &pos, &p, &e)) { bad |= 2; break; }
			int ok = 1;
			const char *c = p;
			if (!tok_skip(&c, e) || !tok_skip(&c, e)) ok = 0; /* id, type */
			x[i] = tok_real(&c, e, &ok);
			y[i] = tok_real(&c, e, &ok);
			z[i] = tok_real(&c, e, &ok);
			if (mux) {
				mux[i] = tok_real(&c, e, &ok);
				muy[i] = tok_real(&c, e, &ok);
				muz[i] = tok_real(&c, e, &ok);
			}
			if (vx) {
				vx[i] = tok_real(&c, e, &ok);
				vy[i] = tok_real(&c, e, &ok);
				vz[i] = tok_real(&c, e, &ok);
			}
			if (!nv)
				for (int k = extra; k > 0; k--)
					if (!tok_skip(&c, e)) ok = 0;          /* unused columns of the wider styles must be there too */
			if (!ok && i != n - 1) bad |= 1; /* a short line fails the frame unless it is the last (snapshot.c:352-356) */
		}
	}
	r->pos = endp;
	if (mux && !bad) /* make_atom_dipole, snapshot.c:440-459: the reference keeps the unit vector (and leaves the field unset below 1e-3: 0 here) */
		for (int i = 0; i < n; i++) {
			const double mu1 = sqrt(mux[i] * mux[i] + muy[i] * muy[i] + muz[i] * muz[i]);
			if (mu1 > 0.001) { mux[i] = mux[i] / mu1; muy[i] = muy[i] / mu1; muz[i] = muz[i] / mu1; }
			else mux[i] = muy[i] = muz[i] = 0.0;
		}
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

/* ---- frame-parallel read-ahead for SMALL text frames -------------------------------------------------------------------
 * The shipped decks' dumps (4 000 / 32 000 atoms per frame) are parsed in about 0.15 / 1 ms per frame by one thread, ten
 * times what the GPU needs for such a frame, and too short for the per-frame fork of sqt_reader_read_atoms to pay (waking
 * sleeping workers costs more than the parse).  Here whole frames are the unit instead: after the count pass has indexed the
 * atom blocks, a few threads parse frames AHEAD of the consumer, each frame by one thread into its own slot of a small ring,
 * and the caller takes them in file order.  No wake-up per frame: the workers only wait when the ring is full.
 * Semantics are the sequential reader's: frames come in file order; the first frame that fails to parse ends the stream
 * (lmp2sqt.c:296-297 stops the file there), as does `max_frames`. */
#include <pthread.h>

struct pf_slot {
	int state;                 /* 0 empty, 1 being filled, 2 ready, 3 end of stream */
	long seq;
	sqt_frame_header hd;
	double *buf;               /* ncomp arrays of cap doubles */
	int cap;
};

struct sqt_prefetch {
	sqt_reader *r;
	int want_v, want_m, ncomp, depth, n_threads;
	long max_frames, next_seq, take_seq;
	int stop;
	struct pf_slot *slot;
	pthread_t *th;
	pthread_mutex_t mu;
	pthread_cond_t cv;
};

long sqt_reader_indexed_frames(const sqt_reader *r) { return (r && r->raw == 0) ? (long)r->n_blk : 0; }

static void *pf_worker(void *arg)
{
	sqt_prefetch *p = (sqt_prefetch *)arg;
	for (;;) {
		pthread_mutex_lock(&p->mu);
		const long seq = p->next_seq++;
		struct pf_slot *s = &p->slot[seq % p->depth];
		while (!p->stop && !(s->state == 0 && s->seq == seq)) pthread_cond_wait(&p->cv, &p->mu); /* the consumer has released seq - depth */
		if (p->stop) { pthread_mutex_unlock(&p->mu); return NULL; }
		s->state = 1;
		pthread_mutex_unlock(&p->mu);

		int ok = seq < p->max_frames && seq < (long)p->r->n_blk;
		if (ok) {
			sqt_reader me = *p->r;             /* private cursor over the shared, read-only mapping and index */
			me.n_threads = 1;
			me.pos = seq == 0 ? 0 : p->r->blk[seq - 1].end_pos;
			me.blk_cursor = (size_t)seq;
			ok = sqt_reader_next_header(&me, &s->hd) && s->hd.n_atoms > 0;
			if (ok) {
				const int n = s->hd.n_atoms;
				if (n > s->cap) {
					free(s->buf);
					s->buf = (double *)malloc(sizeof(double) * (size_t)p->ncomp * (size_t)n);
					s->cap = s->buf ? n : 0;
				}
				ok = s->buf != NULL;
				if (ok) {
					double *c = s->buf, *v = p->want_v ? c + 3 * (size_t)n : NULL, *u = p->want_m ? c + (p->want_v ? 6 : 3) * (size_t)n : NULL;
					ok = sqt_reader_read_atoms_vm(&me, &s->hd, c, c + n, c + 2 * (size_t)n, v, v ? v + n : NULL, v ? v + 2 * (size_t)n : NULL,
					                              u, u ? u + n : NULL, u ? u + 2 * (size_t)n : NULL);
				}
			}
		}
		pthread_mutex_lock(&p->mu);
		s->state = ok ? 2 : 3;
		pthread_cond_broadcast(&p->cv);
		const int done = !ok;
		pthread_mutex_unlock(&p->mu);
		if (done) return NULL;                 /* later frames are never delivered: nothing more for this worker to do */
	}
}

sqt_prefetch *sqt_prefetch_start(sqt_reader *r, int want_v, int want_m, int depth, int n_threads, long max_frames)
{
	if (!r || r->raw != 0 || r->n_blk == 0) return NULL;   /* text dumps whose count pass indexed the atom blocks */
	sqt_prefetch *p = (sqt_prefetch *)calloc(1, sizeof(*p));
	if (!p) return NULL;
	p->r = r; p->want_v = want_v != 0; p->want_m = want_m != 0; p->ncomp = 3 + (want_v ? 3 : 0) + (want_m ? 3 : 0);
	p->n_threads = n_threads < 1 ? 1 : (n_threads > 32 ? 32 : n_threads);
	p->depth = depth < p->n_threads + 1 ? p->n_threads + 1 : depth;
	p->max_frames = max_frames;
	p->slot = (struct pf_slot *)calloc((size_t)p->depth, sizeof(struct pf_slot));
	p->th = (pthread_t *)calloc((size_t)p->n_threads, sizeof(pthread_t));
	if (!p->slot || !p->th) { free(p->slot); free(p->th); free(p); return NULL; }
	for (int i = 0; i < p->depth; i++) p->slot[i].seq = i;
	pthread_mutex_init(&p->mu, NULL);
	pthread_cond_init(&p->cv, NULL);
	for (int i = 0; i < p->n_threads; i++)
		if (pthread_create(&p->th[i], NULL, pf_worker, p) != 0) { p->n_threads = i; break; }
	if (p->n_threads == 0) { sqt_prefetch_stop(p); return NULL; }
	return p;
}

int sqt_prefetch_next(sqt_prefetch *p, sqt_frame_header *hd, const double **soa)
{
	if (!p) return 0;
	pthread_mutex_lock(&p->mu);
	if (p->take_seq > 0) { /* release the slot handed out by the previous call */
		struct pf_slot *prev = &p->slot[(p->take_seq - 1) % p->depth];
		prev->state = 0;
		prev->seq = p->take_seq - 1 + p->depth;
		pthread_cond_broadcast(&p->cv);
	}
	struct pf_slot *s = &p->slot[p->take_seq % p->depth];
	while (!((s->state == 2 || s->state == 3) && s->seq == p->take_seq)) pthread_cond_wait(&p->cv, &p->mu);
	const int ok = s->state == 2;
	if (ok) {
		*hd = s->hd;
		*soa = s->buf;
		p->take_seq++;
	}
	pthread_mutex_unlock(&p->mu);
	return ok;
}

void sqt_prefetch_stop(sqt_prefetch *p)
{
	if (!p) return;
	pthread_mutex_lock(&p->mu);
	p->stop = 1;
	pthread_cond_broadcast(&p->cv);
	pthread_mutex_unlock(&p->mu);
	for (int i = 0; i < p->n_threads; i++) pthread_join(p->th[i], NULL);
	for (int i = 0; i < p->depth; i++) free(p->slot[i].buf);
	pthread_mutex_destroy(&p->mu);
	pthread_cond_destroy(&p->cv);
	free(p->slot); free(p->th); free(p);
}
