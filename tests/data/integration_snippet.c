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

int main(int argc, char **argv)
{
	return argc > 2 ? run(argv[1], argv[2], 1, 1, 1) : 2;
}
