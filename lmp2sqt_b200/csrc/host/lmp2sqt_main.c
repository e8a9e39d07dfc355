/*
 * lmp2sqt_main.c -- command-line driver, drop-in for `lmp2sqt.out` on the hot path named by the
 * north star (`-F -t -i deck dump...`).  Same flags, same deck (and the same deck rewrite), same
 * two-pass file handling and limitCorrAv bookkeeping as the reference's main (lmp2sqt.c:117-336);
 * the per-frame work goes through the C-ABI of include/sqt.h to the GPU engine.
 *
 * Out of scope here: -s (stress), -V (velocity), -M (magnet)
 * are rejected with a message instead of silently computing something else.
 */
#include "../../../include/sqt.h"
#include "../../../include/sqt_host.h"

#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <getopt.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static void die_engine(sqt_handle *h, const char *what, int rc)
{
	fprintf(stderr, "%s failed (%d): %s\n", what, rc, sqt_last_error(h));
	exit(1);
}

int main(int argc, char **argv)
{
	bool flag_f = false, flag_F = false, flag_t = false, flag_d = false;
	char inputFilename[512] = "in.lmp2sqt"; /* lmp2sqt.c:108 */
	int device = 0;
	const char *dev_env = getenv("SQT_DEVICE");
	if (dev_env && *dev_env) device = atoi(dev_env);

	char help_message[2000];
	snprintf(help_message, sizeof(help_message),
	         "%s -f -t -s fn_Dt filename 2 ... \n"
	         "	options: \n"
	         "		-f) collective full calculation required long time and large memory \n"
	         "		-F) collective self calculation required long time and large memory \n"
	         "			space-time, twotime, stress etc...\n"
	         "		-t) only twotime correlation \n"
	         "		-s) only twotime correlation + stress tensor   [not available in the B200 build]\n"
	         "		-V) velocity correlation                        [not available in the B200 build]\n"
	         "		-M) magnet correlation                          [not available in the B200 build]\n"
	         "		-d) Get Amount of dynamic memory allocation \n"
	         "   -i) argment filename for input variable \n",
	         argv[0]);
	if (argc < 2) { puts(help_message); return EXIT_FAILURE; }
	for (;;) {
		int opt = getopt(argc, argv, "tfFsVMi:d");
		if (opt == -1) break;
		switch (opt) {
		case 't': puts("flag_t on : two time correlation"); flag_t = true; break;
		case 'f': puts("flag_f on : collective space time correlation "); flag_f = true; break;
		case 'F':
			puts("flag_F on : self-collective space time correlation ");
			puts("flag_f on : collective space time correlation ");
			flag_f = flag_F = true;
			break;
		case 'd': puts("flag_d on : get amount of dynamic memory allocation"); flag_d = true; break;
		case 'i':
			snprintf(inputFilename, sizeof(inputFilename), "%s", optarg);
			printf("intput filename : %s\n", inputFilename);
			break;
		case 's': case 'V': case 'M':
			fprintf(stderr, "option -%c (stress / velocity / magnet correlators) is outside the B200 hot path; use the reference build for it\n", opt);
			return EXIT_FAILURE;
		default: puts(help_message); return EXIT_FAILURE;
		}
	}
	if (optind == argc) { puts(help_message); return EXIT_FAILURE; }

	sqt_deck deck;
	int dr = sqt_deck_read(inputFilename, &deck);
	if (dr < 0) { /* lmp2sqt.c:2034-2057 */
		sqt_deck_write_template(inputFilename);
		printf("GetDataError\n");
		exit(1);
	}
	int limitCorrAv = deck.limitCorrAv; /* NB: the reference zeroes the global first (:125) and GetNameList then loads the deck value */

	/* ---- count pass, lmp2sqt.c:200-250 ---- */
	bool *files_on = (bool *)calloc(argc + 3, sizeof(bool));
	long *num_data = (long *)calloc(argc + 3, sizeof(long));
	int runable = 0, full_n_snaps = 0;
	for (int a = optind; a < argc; a++) {
		sqt_reader *rd = sqt_reader_open(argv[a], 0);
		if (!rd) { fprintf(stderr, "Can`t open file (%s)!!\n", argv[a]); continue; }
		int n_snap = 0;
		sqt_frame_header hd;
		while (sqt_reader_next_header(rd, &hd) && hd.n_atoms > 0 && sqt_reader_skip_atoms(rd, &hd)) n_snap++;
		sqt_reader_close(rd);
		if (n_snap < 5) {
			fprintf(stderr, "The # of snap is too small(<5) \n, or this file is not valid \nWe would  not use 	this file(%s)!!\n", argv[a]);
		} else if (deck.nCBuffer < 1 || floor((n_snap - deck.nCTime) / (deck.nCTime / deck.nCBuffer)) < 1) {
			/* (the reference divides by nCTime/nCBuffer before it forces nCBuffer >= 1; nCBuffer 0 would be a SIGFPE there) */
			fprintf(stderr, "The # of snap is too small()\nit dont make a infomation\nfloor((n_snap - nCTime) /(nCTime/ nCBuffer)) <1\n"
			                "floor((%d - %d) /(%d/ %d)) <1\nWe would  not use 	this file(%s)!!\n",
			        n_snap, deck.nCTime, deck.nCTime, deck.nCBuffer, argv[a]);
		} else {
			num_data[a] = n_snap;
			limitCorrAv += (int)floor((n_snap - deck.nCTime) / (deck.nCTime / deck.nCBuffer)); /* int division, :244 */
			full_n_snaps += n_snap;
			files_on[a] = true;
			runable++;
		}
	}
	if (runable == 0) { puts("Check your input options or datafile length!!"); exit(1); }
	deck.limitCorrAv = limitCorrAv;
	sqt_deck_print(stderr, &deck);
	sqt_deck_write(inputFilename, &deck); /* UpdateNameList, :257 */

	sqt_params prm;
	memset(&prm, 0, sizeof(prm));
	prm.kVal = deck.kVal; prm.deltaT = deck.deltaT; prm.mass = deck.mass; prm.rVal = deck.rVal;
	prm.limitCorrAv = limitCorrAv; prm.nCBuffer = deck.nCBuffer; prm.nCSpatial = deck.nCSpatial; prm.nCTime = deck.nCTime;
	prm.nCSkip = deck.nCSkip; prm.flag_f = flag_f; prm.flag_F = flag_F; prm.flag_t = flag_t; prm.device = device;
	if (deck.nCBuffer > deck.nCTime) { fputs("Error nCBuffer> nCTime\n", stderr); exit(1); }

	if (flag_d) { /* AmountAllocArray, lmp2sqt.c:1658-1820: report and stop */
		sqt_reader *rd = sqt_reader_open(argv[optind], 1);
		sqt_frame_header hd;
		long long n = 0;
		if (rd && sqt_reader_next_header(rd, &hd)) n = hd.n_atoms;
		sqt_reader_close(rd);
		int B = deck.nCBuffer < 1 ? 1 : deck.nCBuffer;
		long long ring = 3LL * ((B + 15) / 16 * 16) * n * 8;
		long long bufs = (long long)B * deck.nCTime * deck.nCSpatial * (8 + 8 + 4) + 14LL * deck.nCSpatial * deck.nCTime * 8;
		printf("memory Consume : device origin ring %lld MiB, correlator buffers %lld MiB, staging 2 x %lld MiB pinned\n",
		       ring >> 20, bufs >> 20, (3 * n * 8) >> 20);
		puts("End fake Job!!");
		exit(1);
	}

	sqt_handle *h = NULL;
	int rc = sqt_create(&prm, &h);
	if (rc) die_engine(NULL, "sqt_create", rc);
	sqt_results res;
	if (sqt_results_alloc(&res, deck.nCSpatial, deck.nCTime)) { fputs("Reserving memory Error!!!!!!\n", stdout); exit(1); }

	int full_n_snaps_index = 0, progress = -1;
	double kVal = deck.kVal;
	for (int a = optind; a < argc; a++) {
		if (!files_on[a]) continue;
		sqt_reader *rd = sqt_reader_open(argv[a], 0);
		sqt_begin_file(h); /* InitSpacetimeCorr, :282 */
		sqt_frame_header hd;
		if (sqt_reader_next_header(rd, &hd)) { /* Init_reciprocal_space, :2171-2202 (message only: kVal is reset per frame) */
			double L0 = hd.xhi - hd.xlo;
			int n_mul = (int)round(kVal / (2. * M_PI / L0));
			if (n_mul <= 0) n_mul = 1;
			double new_dk = (2. * M_PI / L0) * n_mul;
			fprintf(stderr, "Update for input dk param: %f -> %f \n", kVal, new_dk);
			kVal = new_dk;
		}
		sqt_reader_rewind(rd);
		fprintf(stderr, "FULL_SNAPS = %5d\n", full_n_snaps);
		for (long ns = 0; ns < num_data[a]; ns++) {
			if (!sqt_reader_next_header(rd, &hd)) break;
			double *stage = sqt_stage_acquire(h, hd.n_atoms);
			if (!stage) die_engine(h, "sqt_stage_acquire", -1);
			if (!sqt_reader_read_atoms(rd, &hd, stage, stage + hd.n_atoms, stage + 2 * (size_t)hd.n_atoms)) break;
			rc = sqt_push_frame(h, stage, hd.n_atoms, hd.n_atoms, hd.xlo, hd.xhi); /* EvalSpacetimeCorr, :300 */
			if (rc < 0) die_engine(h, "sqt_push_frame", rc);
			if (rc == SQT_FLUSH) { /* PrintProcess from AccumSpacetimeCorr, :438-440 */
				rc = sqt_fetch_results(h, &res);
				if (rc) die_engine(h, "sqt_fetch_results", rc);
				if (sqt_host_print_process(NULL, &res, &deck, flag_f, flag_F, flag_t, NULL)) { perror("PrintProcess"); exit(1); }
			}
			full_n_snaps_index++;
			int new_progress = (int)(10000.0 * full_n_snaps_index / full_n_snaps);
			if (new_progress != progress) {
				progress = new_progress;
				fprintf(stderr, "\r %4.2f%%(", progress * .01);
				fprintf(stderr, "main:%4.2f%%", progress * .01);
				fflush(stderr);
			}
		}
		sqt_reader_close(rd);
		fprintf(stderr, "\nEnd : file : %s\n", argv[a]);
	}
	rc = sqt_finish(h, &res);
	if (rc < 0) die_engine(h, "sqt_finish", rc);
	if (rc == 1) { /* lmp2sqt.c:327-333 */
		fprintf(stderr, "limit corr  = %d, countCorrAv = %d\n", limitCorrAv, res.countCorrAv);
		if (sqt_host_print_process(NULL, &res, &deck, flag_f, flag_F, flag_t, NULL)) { perror("PrintProcess"); exit(1); }
	}
	fprintf(stderr, "engine: %s; %lld kernel launches\n", sqt_kernel_config(h), sqt_kernel_launches(h));
	sqt_results_free(&res);
	sqt_destroy(h);
	free(files_on);
	free(num_data);
	return 0;
}
