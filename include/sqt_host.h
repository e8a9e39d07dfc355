/*
 * sqt_host.h -- host-side C code that stays on the CPU around the engine (libsqt_host.so):
 * the deck, the LAMMPS text-dump reader, the normalisation and the result writers.  These keep
 * the reference's file formats byte-for-byte so that the new path is a drop-in for
 * `lmp2sqt.out -F -t -i deck dump...`.
 *
 *   reference (file:line)                                    here
 *   -------------------------------------------------------  --------------------------------
 *   GetNameList            lmp2sqt.c:2014-2111                sqt_deck_read
 *   PrintNameList2File     lmp2sqt.c:2126-2169                sqt_deck_print
 *   UpdateNameList         lmp2sqt.c:2112-2125                sqt_deck_write
 *   ReadDumpForOnlyCheck   snapshot.c:25-165                  sqt_dump_read_header + sqt_dump_skip_atoms
 *   ReadDump               snapshot.c:166-382                 sqt_dump_read_header + sqt_dump_read_atoms (SoA)
 *                                                             sqt_reader_* (mmap + threads, same semantics)
 *   Alloc_more tail        lmp2sqt.c:1987-2009 (factorDr)     sqt_host_factor_dr
 *   prePrintProcess        lmp2sqt.c:576-675                  sqt_host_normalise
 *   PrintProcess           lmp2sqt.c:338-367                  sqt_host_print_process
 *     PrintEtc             lmp2sqt.c:865-1093  (SSF, Dt)
 *     PrintSpacetimeCorr   lmp2sqt.c:677-766   (output.txt)
 *     PrintSpacetimeCorr_binary lmp2sqt.c:768-864 (output.binary)
 */
#ifndef SQT_HOST_H
#define SQT_HOST_H

#include <stdio.h>
#include "sqt.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SQT_DECK_NNAMES 9
typedef struct {
	double kVal, deltaT, mass, rVal;
	int limitCorrAv, nCBuffer, nCSpatial, nCTime, nCSkip;
	int status[SQT_DECK_NNAMES]; /* vStatus per name, DATALIST order: 0 no data, 1 ok, 2 missing, 3 extra, 4 multiply defined */
} sqt_deck;

/* Returns 1 if every name was read exactly once ("ok"), 0 otherwise (the reference ignores the
 * value, lmp2sqt.c:199), -1 if the file does not exist (the reference then writes a template deck
 * and exits; sqt_deck_write_template does the former). */
int sqt_deck_read(const char *path, sqt_deck *d);
int sqt_deck_print(FILE *fp, const sqt_deck *d);
int sqt_deck_write(const char *path, const sqt_deck *d);
int sqt_deck_write_template(const char *path);

typedef struct {
	long timestep;
	int n_atoms;
	double xlo, xhi, ylo, yhi, zlo, zhi;
	int pbcType;
	int style; /* 0: id type xu yu zu; 1: +vx vy vz; 2: +mux muy muz; 3: +mux muy muz vx vy vz */
} sqt_frame_header;

/* 1 = header read, 0 = end of file or not a frame (the reference stops the file there). */
int sqt_dump_read_header(FILE *fp, sqt_frame_header *hd);
/* Atom lines into SoA x[n], y[n], z[n] in file-line order (`id` is ignored, snapshot.c:357-368). 1 ok / 0 short file. */
int sqt_dump_read_atoms(FILE *fp, const sqt_frame_header *hd, double *x, double *y, double *z);
int sqt_dump_skip_atoms(FILE *fp, const sqt_frame_header *hd);

/* High-rate variant of the same reader: memory-mapped file, atom block parsed by n_threads threads (0 = auto,
 * SQT_READER_THREADS overrides; a thread is only added per SQT_READER_LINES_PER_THREAD = 131072 atom lines, smaller blocks are
 * parsed by the caller alone) straight into the SoA staging buffer.  Same format rules and conversions.  A count pass
 * (sqt_reader_skip_atoms over the file) indexes the atom blocks; after sqt_reader_rewind the parse pass reuses that index. */
/* The same reader also accepts the raw binary frame file (an extension; the reference reads text only):
 *   "SQTRAW01", then per frame: int64 timestep, int64 n_atoms, f64 xlo xhi ylo yhi zlo zhi, f64 x[n], y[n], z[n]
 * (little endian, atoms in dump order).  It is recognised by its magic; sqt_raw_append_frame writes one frame. */
typedef struct sqt_reader sqt_reader;
sqt_reader *sqt_reader_open(const char *path, int n_threads);
int sqt_reader_is_raw(const sqt_reader *r); /* 0 LAMMPS text, 1 SQTRAW01 frame file, 2 LAMMPS native binary (`dump custom`, orthogonal box, id type xu yu zu ...) */
int sqt_raw_append_frame(const char *path, int create, long long timestep, long long n_atoms, const double box[6],
                         const double *x, const double *y, const double *z);
void sqt_reader_close(sqt_reader *r);
void sqt_reader_rewind(sqt_reader *r);
int sqt_reader_next_header(sqt_reader *r, sqt_frame_header *hd);
int sqt_reader_read_atoms(sqt_reader *r, const sqt_frame_header *hd, double *x, double *y, double *z);
/* -V: the same with the velocity columns of the `id type xu yu zu vx vy vz` style (snapshot.c:285-300).  Any other style
 * fails the frame: the reference would correlate uninitialised memory there (make_atom / make_atom_dipole never set vx). */
int sqt_reader_read_atoms_v(sqt_reader *r, const sqt_frame_header *hd, double *x, double *y, double *z, double *vx, double *vy, double *vz);
/* -M: also the dipole columns of the `id type xu yu zu mux muy muz [vx vy vz]` styles, normalised to unit vectors like
 * make_atom_dipole does (snapshot.c:440-459).  With mux != NULL, velocities (vx != NULL) come from the 11-column style only. */
int sqt_reader_read_atoms_vm(sqt_reader *r, const sqt_frame_header *hd, double *x, double *y, double *z, double *vx, double *vy, double *vz,
                             double *mux, double *muy, double *muz);
int sqt_reader_skip_atoms(sqt_reader *r, const sqt_frame_header *hd);
/* Per-type partials: the `type` column of the atom block at the read position (text dumps; the position is not advanced).
 * 1 ok / 0 not available. */
int sqt_reader_peek_types(sqt_reader *r, const sqt_frame_header *hd, int *types);

/* Frame-parallel read-ahead for small text frames (the shipped decks' dumps: 4 000 / 32 000 atoms): after a count pass over the
 * file (sqt_reader_skip_atoms on every frame, which indexes the atom blocks) and sqt_reader_rewind, n_threads workers parse
 * whole frames ahead of the caller, `depth` frames at most, and sqt_prefetch_next hands them out in file order: header and one
 * SoA block x[n] y[n] z[n] [vx vy vz] [mux muy muz] (valid until the next call).  Same acceptance rules as
 * sqt_reader_read_atoms_vm; the first frame that fails ends the stream, as does max_frames.  NULL if the reader is not an
 * indexed text dump. */
typedef struct sqt_prefetch sqt_prefetch;
long sqt_reader_indexed_frames(const sqt_reader *r);
sqt_prefetch *sqt_prefetch_start(sqt_reader *r, int want_v, int want_m, int depth, int n_threads, long max_frames);
int sqt_prefetch_next(sqt_prefetch *p, sqt_frame_header *hd, const double **soa);
void sqt_prefetch_stop(sqt_prefetch *p);

/* factorDr[nr], lmp2sqt.c:1987-2009, from what the engine recorded at the first processed frame. */
void sqt_host_factor_dr(const sqt_results *r, double rVal, double *factorDr);
/* prePrintProcess: normalises r's arrays IN PLACE (like the reference does with its globals). */
void sqt_host_normalise(sqt_results *r, const sqt_deck *d, int flag_f, int flag_F, int flag_t);
void sqt_host_normalise_v(sqt_results *r, const sqt_deck *d, int flag_f, int flag_F, int flag_t, int flag_V);
/* PrintProcess: normalise, then write SSF%03d.info.<date>, Dt%03d.info.<date>, output.txt,
 * output.binary, output.log into `dir` (NULL = cwd).  `date` is the yymmdd suffix (NULL = today,
 * lmp2sqt.c:127-132).  Returns 0 or -1. */
int sqt_host_print_process(const char *dir, sqt_results *r, const sqt_deck *d, int flag_f, int flag_F, int flag_t,
                           const char *date);
/* All getopt flags of the reference (lmp2sqt.c:152-191) in one struct: -f -F -t -V -M -s.  F implies f, s implies t. */
typedef struct { int f, F, t, V, M, s; } sqt_flags;
/* prePrintProcess / PrintProcess for any flag set: with -M also the C2_mu_* (and with -V C2_v_mu / C2_mu_v) files and the Cmm
 * column of the Dt file, with -s the Stress%03d.info.<date> file and the msdx msdy msdz columns (lmp2sqt.c:933-1076). */
void sqt_host_normalise_x(sqt_results *r, const sqt_deck *d, const sqt_flags *fl);
int sqt_host_print_process_x(const char *dir, sqt_results *r, const sqt_deck *d, const sqt_flags *fl, const char *date);
/* Per-type partials (sqt_results.n_types > 1; an extension, the reference has no per-type code).  Normalisation, in place:
 *   F_ab *= 1 / (N 3 cnt)   (so that sum_ab F_ab = F(q,t) of output.txt),   F_s,a *= 1 / (N_a 3 cnt)   (F_s,a(q,0) = 1),
 *   MSD_a, MQD_a *= 1 / (N_a cnt),   van Hove_a[lag >= 1] *= factorDr N / N_a.
 * sqt_host_print_types writes them as text blocks into `types%03d.info.<date>` beside the Dt / SSF files (same index rule). */
void sqt_host_normalise_types(sqt_results *r, const sqt_deck *d);
int sqt_host_print_types(const char *dir, sqt_results *r, const sqt_deck *d, const sqt_flags *fl, const char *date);
/* The same with -V: also normalises the current correlators and writes C2_v_v / C2_v_rho / C2_rho_v%03d.info.<date> and the
 * Cvv, Cvcmvcm columns of the Dt file (lmp2sqt.c:933-993, 1033-1076). */
int sqt_host_print_process_v(const char *dir, sqt_results *r, const sqt_deck *d, int flag_f, int flag_F, int flag_t, int flag_V,
                             const char *date);

/* ---- post-processing of the correlators (reference: analysis_binary.c) ---------------------
 * Tables are [k][t] (k-major) like the reference's corrSum.  Sequence: load (or from_results) -> time | fourier -> print. */
#define SQT_AN_TYPES 4 /* full-density, self-density, cross-density, self-vanHove */
typedef struct {
	int nCSpatial, nCTime, nData;      /* nData = sets of output.binary that were averaged */
	double deck_deltaT;                /* deltaT of the deck: the time axis of the printed tables (deltaTCorr, analysis_binary.c:430) */
	double deltaT, kVal, rVal;         /* from the block headers of output.binary (deltaT*(nCSkip+1), 2 pi/L, rVal) */
	double *mean[SQT_AN_TYPES], *sd[SQT_AN_TYPES], *err[SQT_AN_TYPES]; /* corrSum / corrSumSq / corrSumErr */
	double *F[SQT_AN_TYPES], *F1[SQT_AN_TYPES], *Gamma[SQT_AN_TYPES];  /* F(q,t), dF/dt, d ln F/dt */
	double *D[SQT_AN_TYPES], *Dc[SQT_AN_TYPES], *H[SQT_AN_TYPES], *Hc[SQT_AN_TYPES]; /* -Gamma/q^2, -ln(F/F0)/(q^2 t), times F0 */
	int fourier, omegaMax, nv;
	double tMax;
} sqt_analysis;
/* import_settings + load_raw_data + scaling_data (analysis_binary.c:407-447, 643-771).  0 ok, <0 error. */
int sqt_analysis_load(sqt_analysis *a, const char *deck_path, const char *binary_path, int nSetSkip);
/* Same tables straight from results fetched from the engine and normalised by sqt_host_normalise (one set). */
int sqt_analysis_from_results(sqt_analysis *a, const sqt_results *normalised, const sqt_deck *d, int flag_F);
/* do_not_fourier (:793-897): derivative tables, "<block>00.out" summaries into outdir (NULL = cwd), scaling by F(q,0). */
int sqt_analysis_time(sqt_analysis *a, int nValDiff, double under_limit, const char *outdir);
/* do_fourier (:772-792): cosine transform of the first three blocks; 2 (nCTime - 1) must be a power of two. */
int sqt_analysis_fourier(sqt_analysis *a, int window);
/* print_output (:898-963): Fq?_scaled, Fq?, GammaQ?, Dq?_c, Dq?, D0Hq?_c, D0Hq?, Fq?1 `.<block>.info` files, %.13e. */
int sqt_analysis_print(const sqt_analysis *a, const char *outdir);
void sqt_analysis_free(sqt_analysis *a);

/* ---- the TEXT post-processor (reference: analysis.c:138-536 = AverageFqt2FqOrFt.c:135-507), on output.txt ---------------
 * Same struct (Dc / Hc stay unused).  The reference's stale row parse is kept: the first nCSpatial numbers of a row are
 * taken, so column 0 of the density blocks is the time.  "-" reads stdin. */
int sqt_analysis_text_load(sqt_analysis *a, const char *deck_path, const char *txt_path, int nSetSkip);
int sqt_analysis_text_time(sqt_analysis *a, int nValDiff, const char *outdir);       /* :370-455; also prints the "nv = ..." line */
int sqt_analysis_text_fourier(sqt_analysis *a, int window);                          /* :349-368 */
/* Fqt, GammaQT, Dqt, D0Hqt, Fqt1 `.<block>.info` at %9.4e, and Print_R2_data's echo on stdout (:458-536) */
int sqt_analysis_text_print(const sqt_analysis *a, const char *outdir);

#ifdef __cplusplus
}
#endif
#endif
