/* INTEGRATION.md section 2, as a translation unit: the calls a maintainer of the reference would put into main().
 * tests/test_host.py compiles and links it against the two libraries (it is not run: no GPU on the CPU test box). */
#include "sqt.h"
#include "sqt_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int run(const char *deck_path, const char *dump_path, int flag_f, int flag_F, int flag_t)
{
	sqt_deck deck;
	if (sqt_deck_read(deck_path, &deck) < 0) return 1;
	sqt_params prm;
	memset(&prm, 0, sizeof(prm));
	prm.kVal = deck.kVal; prm.deltaT = deck.deltaT; prm.mass = deck.mass; prm.rVal = deck.rVal;
	prm.limitCorrAv = deck.limitCorrAv; prm.nCBuffer = deck.nCBuffer; prm.nCSpatial = deck.nCSpatial; prm.nCTime = deck.nCTime;
	prm.nCSkip = deck.nCSkip; prm.flag_f = flag_f; prm.flag_F = flag_F; prm.flag_t = flag_t; prm.device = 0;
	sqt_handle *h;
	sqt_results res;
	if (sqt_create(&prm, &h)) { fputs(sqt_last_error(NULL), stderr); return 1; }   /* AllocArray, lmp2sqt.c:1821 */
	if (sqt_results_alloc(&res, deck.nCSpatial, deck.nCTime)) return 1;

	FILE *fp = fopen(dump_path, "r");
	if (!fp) return 1;
	sqt_begin_file(h);                                       /* InitSpacetimeCorr, lmp2sqt.c:282 */
	for (;;) {                                               /* ReadDump + EvalSpacetimeCorr + FreeSnapshot, :295-302 */
		sqt_frame_header hd;
		if (!sqt_dump_read_header(fp, &hd)) break;
		double *stage = sqt_stage_acquire(h, hd.n_atoms);    /* pinned SoA: x[n] y[n] z[n] */
		if (!stage || !sqt_dump_read_atoms(fp, &hd, stage, stage + hd.n_atoms, stage + 2 * (size_t)hd.n_atoms)) break;
		int rc = sqt_push_frame(h, stage, hd.n_atoms, hd.n_atoms, hd.xlo, hd.xhi);
		if (rc < 0) { fputs(sqt_last_error(h), stderr); return 1; }
		if (rc == SQT_FLUSH) {                               /* PrintProcess from AccumSpacetimeCorr, :438-440 */
			sqt_fetch_results(h, &res);
			sqt_host_print_process(NULL, &res, &deck, flag_f, flag_F, flag_t, NULL);
		}
	}
	fclose(fp);
	if (sqt_finish(h, &res) == 1)                            /* end of run, :327-333 */
		sqt_host_print_process(NULL, &res, &deck, flag_f, flag_F, flag_t, NULL);
	/* post-processing straight from memory (analysis_binary.c) */
	sqt_analysis an;
	if (sqt_analysis_from_results(&an, &res, &deck, flag_F) == 0) {
		sqt_analysis_time(&an, 4, 0.0820849986, NULL);
		sqt_analysis_print(&an, NULL);
		sqt_analysis_free(&an);
	}
	sqt_results_free(&res);
	sqt_destroy(h);
	return 0;
}

/* INTEGRATION.md sections 3 and 3b: the same loop with every flag of the reference, G GPUs behind the one caller thread
 * (one reader, one pinned frame, G chunk copies) and, optionally, the per-type partials of the dump's `type` column. */
int run_all(const char *deck_path, const char *dump_path, const sqt_flags *fl, int n_gpus, int n_types)
{
	sqt_deck deck;
	if (sqt_deck_read(deck_path, &deck) < 0) return 1;
	sqt_params prm;
	memset(&prm, 0, sizeof(prm));
	prm.kVal = deck.kVal; prm.deltaT = deck.deltaT; prm.mass = deck.mass; prm.rVal = deck.rVal;
	prm.limitCorrAv = deck.limitCorrAv; prm.nCBuffer = deck.nCBuffer; prm.nCSpatial = deck.nCSpatial; prm.nCTime = deck.nCTime;
	prm.nCSkip = deck.nCSkip; prm.flag_f = fl->f; prm.flag_F = fl->F; prm.flag_t = fl->t;
	prm.flag_V = fl->V; prm.flag_M = fl->M; prm.flag_s = fl->s; prm.n_gpus = n_gpus; prm.n_types = n_types;
	sqt_multi *m;
	sqt_results res;
	if (sqt_multi_create(&prm, &m)) { fputs(sqt_multi_last_error(NULL), stderr); return 1; }
	if (sqt_results_alloc(&res, deck.nCSpatial, deck.nCTime)) return 1;
	if (n_types > 1 && sqt_results_alloc_types(&res, n_types)) return 1;
	sqt_reader *rd = sqt_reader_open(dump_path, 0);
	if (!rd) return 1;
	const int want_v = fl->V || fl->s, want_m = fl->M;
	int types_given = n_types <= 1;
	sqt_multi_begin_file(m);
	for (;;) {
		sqt_frame_header hd;
		if (!sqt_reader_next_header(rd, &hd)) break;
		const size_t n = (size_t)hd.n_atoms;
		if (!types_given) {                                  /* species once, from the first frame's `type` column */
			int *types = (int *)malloc(sizeof(int) * n);
			if (!types || !sqt_reader_peek_types(rd, &hd, types) || sqt_multi_set_types(m, types, hd.n_atoms)) return 1;
			free(types);
			types_given = 1;
		}
		double *st = sqt_multi_stage_acquire(m, hd.n_atoms);  /* x y z [vx vy vz] [mux muy muz], n doubles each */
		if (!st) return 1;
		double *v = want_v ? st + 3 * n : NULL, *u = want_m ? st + (want_v ? 6 : 3) * n : NULL;
		if (!sqt_reader_read_atoms_vm(rd, &hd, st, st + n, st + 2 * n, v, v ? v + n : NULL, v ? v + 2 * n : NULL,
		                              u, u ? u + n : NULL, u ? u + 2 * n : NULL)) break;
		int rc = sqt_multi_push_frame(m, st, hd.n_atoms, hd.xlo, hd.xhi);
		if (rc < 0) { fputs(sqt_multi_last_error(m), stderr); return 1; }
		if (rc == SQT_FLUSH && sqt_multi_fetch_results(m, &res) == 0) {
			sqt_host_print_process_x(NULL, &res, &deck, fl, NULL);
			sqt_host_print_types(NULL, &res, &deck, fl, NULL);
		}
	}
	sqt_reader_close(rd);
	if (sqt_multi_finish(m, &res) == 1) {
		sqt_host_print_process_x(NULL, &res, &deck, fl, NULL);
		sqt_host_print_types(NULL, &res, &deck, fl, NULL);
	}
	sqt_results_free(&res);
	sqt_multi_destroy(m);
	return 0;
}

int main(int argc, char **argv)
{
	const sqt_flags all = { 1, 1, 1, 1, 1, 1 };
	if (argc > 3) return run_all(argv[1], argv[2], &all, atoi(argv[3]), 0);
	return argc > 2 ? run(argv[1], argv[2], 1, 1, 1) : 2;
}
