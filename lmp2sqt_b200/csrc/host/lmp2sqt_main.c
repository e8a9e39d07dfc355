/*
 * lmp2sqt_main.c -- command-line driver, drop-in for `lmp2sqt.out` on the hot path named by the
 * north star (`-F -t -i deck dump...`).  Same flags, same deck (and the same deck rewrite), same
 * two-pass file handling and limitCorrAv bookkeeping as the reference's main (lmp2sqt.c:117-336);
 * the per-frame work goes through the C-ABI of include/sqt.h to the GPU engine.
 *
 * -V (velocity / current correlators, lmp2sqt.c:172), -M (dipole-density correlators, :177) and -s (per-axis MSD and the
 * stress-like correlator, :156-160) are served too.  Which dump style carries what follows the
 * reference's reader: velocities come from `id type xu yu zu vx vy vz`, dipoles from `... mux muy muz [vx vy vz]`
 * (snapshot.c:263-331); asking for both reads the 11-column style in full, which the reference's reader never does
 * (its 11-column branch is shadowed by the 8-column dipole test, :262 vs :308).
 *
 * Extension: `-T n` accumulates the per-type partials F_ab(q,t), F_s,a(q,t), MSD_a and the van Hove table of the n species the
 * dump's `type` column names (written to types%03d.info.<date>; everything else is unchanged).
 * Extension: `-g N` (or SQT_GPUS=N) shards every frame by particle over N GPUs of this box (devices SQT_DEVICE ..
 * SQT_DEVICE + N - 1): one reader thread, one pinned frame, N asynchronous chunk copies.
 */
#include "../../../include/sqt.h"
#include "../../../include/sqt_host.h"

#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <getopt.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void die_engine(sqt_multi *m, const char *what, int rc)
{
	fprintf(stderr, "%s failed (%d): %s\n", what, rc, sqt_multi_last_error(m));
	exit(1);
}

int main(int argc, char **argv)
{
	bool flag_f = false, flag_F = false, flag_t = false, flag_d = false, flag_V = false, flag_M = false, flag_s = false;
	char inputFilename[512] = "in.lmp2sqt"; /* lmp2sqt.c:108 */
	int device = 0, n_gpus = 1, n_types = 0;
	const char *dev_env = getenv("SQT_DEVICE");
	if (dev_env && *dev_env) device = atoi(dev_env);
	const char *gpus_env = getenv("SQT_GPUS");
	if (gpus_env && *gpus_env) n_gpus = atoi(gpus_env);

	char help_message[2000];
	snprintf(help_message, sizeof(help_message),
	         "%s -f -t -s fn_Dt filename 2 ... \n"
	         "	options: \n"
	         "		-f) collective full calculation required long time and large memory \n"
	         "		-F) collective self calculation required long time and large memory \n"
	         "			space-time, twotime, stress etc...\n"
	         "		-t) only twotime correlation \n"
	         "		-s) only twotime correlation + stress tensor \n"
	         "		-V) velocity correlation \n"
	         "		-M) magnet correlation \n"
	         "		-d) Get Amount of dynamic memory allocation \n"
	         "   -i) argment filename for input variable \n"
	         "   -g) number of GPUs of this box to shard the particles over   [B200 build only; default 1 or $SQT_GPUS]\n"
	         "   -T) number of species (dump column `type` = 1..n): per-type partials F_ab, F_s,a, MSD_a   [B200 build only]\n",
	         argv[0]);
	if (argc < 2) { puts(help_message); return EXIT_FAILURE; }
	for (;;) {
		int opt = getopt(argc, argv, "tfFsVMi:dg:T:");
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
		case 'V': puts("flag_velocity on : velocity correlation on"); flag_V = true; break; /* lmp2sqt.c:171-173 */
		case 'g': n_gpus = atoi(optarg); break;
		case 'T': n_types = atoi(optarg); break;
		case 's': /* lmp2sqt.c:156-160 */
			puts("flag_s on : stress tensor calculation on");
			puts("require flag_t : also on");
			flag_t = true;
			flag_s = true;
			break;
		case 'M': puts("flag_magnet on : magnet correlation on"); flag_M = true; break; /* :177-179 */
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
	if (n_gpus < 1) n_gpus = 1;
	{ /* before anything is counted or the deck rewritten: what the kernels cannot serve is refused here, with a message */
		const int max_s = (flag_V || flag_M) ? SQT_MAX_NCSPATIAL_V : SQT_MAX_NCSPATIAL;
		if (deck.nCSpatial > max_s) {
			fprintf(stderr, "nCSpatial %d is beyond what the B200 kernels cover (%d%s); use the reference build for this deck\n",
			        deck.nCSpatial, max_s, (flag_V || flag_M) ? " with -V / -M" : "");
			exit(1);
		}
	}

	/* ---- count pass, lmp2sqt.c:200-250 ---- */
	bool *files_on = (bool *)calloc(argc + 3, sizeof(bool));
	long *num_data = (long *)calloc(argc + 3, sizeof(long));
	/* one reader per input file for both passes: the count pass indexes the atom blocks (line starts), the parse pass
	 * reuses that index after a rewind instead of scanning every block for newlines a second time */
	sqt_reader **readers = (sqt_reader **)calloc(argc + 3, sizeof(sqt_reader *));
	int runable = 0, full_n_snaps = 0;
	for (int a = optind; a < argc; a++) {
		sqt_reader *rd = readers[a] = sqt_reader_open(argv[a], 0);
		if (!rd) { fprintf(stderr, "Can`t open file (%s)!!\n", argv[a]); continue; }
		int n_snap = 0;
		sqt_frame_header hd;
		while (sqt_reader_next_header(rd, &hd) && hd.n_atoms > 0 && sqt_reader_skip_atoms(rd, &hd)) n_snap++;
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
	prm.n_gpus = n_gpus; prm.flag_V = flag_V; prm.flag_M = flag_M; prm.flag_s = flag_s; prm.n_types = n_types > 1 ? n_types : 0;
	const sqt_flags fl = { flag_f, flag_F, flag_t, flag_V, flag_M, flag_s };
	const bool want_v = flag_V || flag_s, want_m = flag_M;
	if (deck.nCBuffer > deck.nCTime) { fputs("Error nCBuffer> nCTime\n", stderr); exit(1); }

	if (flag_d) { /* AmountAllocArray, lmp2sqt.c:1658-1820: report and stop */
		sqt_reader *rd = readers[optind];
		sqt_frame_header hd;
		long long n = 0;
		if (rd) { sqt_reader_rewind(rd); if (sqt_reader_next_header(rd, &hd)) n = hd.n_atoms; }
		int B = deck.nCBuffer < 1 ? 1 : deck.nCBuffer;
		long long ring = 3LL * ((B + 15) / 16 * 16) * n * 8;
		long long bufs = (long long)B * deck.nCTime * deck.nCSpatial * (8 + 8 + 4) + 14LL * deck.nCSpatial * deck.nCTime * 8;
		printf("memory Consume : device origin ring %lld MiB, correlator buffers %lld MiB, staging 2 x %lld MiB pinned\n",
		       ring >> 20, bufs >> 20, (3 * n * 8) >> 20);
		puts("End fake Job!!");
		exit(1);
	}

	/* stdout carries exactly the reference's banner lines: whatever NCCL has to say while the communicator is set up (its
	 * version line under NCCL_DEBUG=VERSION ignores NCCL_DEBUG_FILE) goes to stderr */
	int saved_stdout = -1;
	if (n_gpus > 1) {
		setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
		fflush(stdout);
		saved_stdout = dup(STDOUT_FILENO);
		if (saved_stdout >= 0) dup2(STDERR_FILENO, STDOUT_FILENO);
	}
	sqt_multi *h = NULL;
	int rc = sqt_multi_create(&prm, &h);
	if (saved_stdout >= 0) {
		fflush(stdout);
		dup2(saved_stdout, STDOUT_FILENO);
		close(saved_stdout);
	}
	if (rc) die_engine(NULL, "sqt_multi_create", rc);
	sqt_results res;
	if (sqt_results_alloc(&res, deck.nCSpatial, deck.nCTime)) { fputs("Reserving memory Error!!!!!!\n", stdout); exit(1); }
	if (n_types > 1 && sqt_results_alloc_types(&res, n_types)) { fputs("Reserving memory Error!!!!!!\n", stdout); exit(1); }
	bool types_set = n_types <= 1;

	int full_n_snaps_index = 0, progress = -1;
	const double t_loop0 = now_s();
	double kVal = deck.kVal;
	for (int a = optind; a < argc; a++) {
		if (!files_on[a]) continue;
		sqt_reader *rd = readers[a];
		if (!rd) continue;
		sqt_reader_rewind(rd);
		sqt_multi_begin_file(h); /* InitSpacetimeCorr, :282 */
		sqt_frame_header hd;
		if (sqt_reader_next_header(rd, &hd)) { /* Init_reciprocal_space, :2171-2202 (message only: kVal is reset per frame) */
			double L0 = hd.xhi - hd.xlo;
			int n_mul = (int)round(kVal / (2. * M_PI / L0));
			if (n_mul <= 0) n_mul = 1;
			double new_dk = (2. * M_PI / L0) * n_mul;
			fprintf(stderr, "Update for input dk param: %f -> %f \n", kVal, new_dk);
			kVal = new_dk;
			if (want_v && !want_m && hd.style != 1) {
				/* the reference correlates whatever malloc left in Atom.vx for such dumps (snapshot.c:426-437) */
				fprintf(stderr, "-V / -s need dumps written as `id type xu yu zu vx vy vz` (file %s)\n", argv[a]);
				exit(1);
			}
			if (want_m && (want_v ? hd.style != 3 : (hd.style != 2 && hd.style != 3))) {
				fprintf(stderr, "-M needs dumps written as `id type xu yu zu mux muy muz%s` (file %s)\n", want_v ? " vx vy vz" : "", argv[a]);
				exit(1);
			}
			if (!types_set) { /* species from the first frame that is read: line order is fixed for the whole run */
				int *types = (int *)malloc(sizeof(int) * (size_t)(hd.n_atoms > 0 ? hd.n_atoms : 1));
				if (!sqt_reader_peek_types(rd, &hd, types)) { fprintf(stderr, "-T needs a text dump with a `type` column (file %s)\n", argv[a]); exit(1); }
				rc = sqt_multi_set_types(h, types, hd.n_atoms);
				if (rc) die_engine(h, "sqt_set_types", rc);
				free(types);
				types_set = true;
			}
		}
		/* small text frames (the shipped decks: 4 000 / 32 000 atoms) are parsed AHEAD, whole frames by a few threads, and taken
		 * in file order here; large frames keep the in-frame threads of the reader (sqt_host.h) */
		sqt_prefetch *pf = NULL;
		{
			const char *pe = getenv("SQT_PREFETCH");
			if (!(pe && *pe == '0') && sqt_reader_is_raw(rd) == 0 && hd.n_atoms > 0 && hd.n_atoms < 262144 &&
			    sqt_reader_indexed_frames(rd) >= num_data[a] && num_data[a] >= 8) {
				int thr = pe && atoi(pe) > 0 ? atoi(pe) : 6;
				pf = sqt_prefetch_start(rd, want_v, want_m, 2 * thr + 2, thr, num_data[a]);
			}
		}
		sqt_reader_rewind(rd);
		fprintf(stderr, "FULL_SNAPS = %5d\n", full_n_snaps);
		for (long ns = 0; ns < num_data[a]; ns++) {
			const double *ahead = NULL;
			if (pf ? !sqt_prefetch_next(pf, &hd, &ahead) : !sqt_reader_next_header(rd, &hd)) break;
			double *stage = sqt_multi_stage_acquire(h, hd.n_atoms);
			if (!stage) die_engine(h, "sqt_stage_acquire", -1);
			const size_t na = (size_t)hd.n_atoms;
			if (want_v && !want_m && hd.style != 1) {
				/* the reference correlates whatever malloc left in Atom.vx for such dumps (snapshot.c:426-437) */
				fprintf(stderr, "-V / -s need dumps written as `id type xu yu zu vx vy vz` (file %s)\n", argv[a]);
				exit(1);
			}
			if (want_m && (want_v ? hd.style != 3 : (hd.style != 2 && hd.style != 3))) {
				fprintf(stderr, "-M needs dumps written as `id type xu yu zu mux muy muz%s` (file %s)\n", want_v ? " vx vy vz" : "", argv[a]);
				exit(1);
			}
			double *sv = want_v ? stage + 3 * na : NULL, *sm = want_m ? stage + (want_v ? 6 : 3) * na : NULL;
			if (pf) memcpy(stage, ahead, sizeof(double) * na * (size_t)(3 + (want_v ? 3 : 0) + (want_m ? 3 : 0)));
			else if (!sqt_reader_read_atoms_vm(rd, &hd, stage, stage + na, stage + 2 * na, sv, sv ? sv + na : NULL, sv ? sv + 2 * na : NULL,
			                                   sm, sm ? sm + na : NULL, sm ? sm + 2 * na : NULL)) break;
			rc = sqt_multi_push_frame(h, stage, hd.n_atoms, hd.xlo, hd.xhi); /* EvalSpacetimeCorr, :300 */
			if (rc < 0) die_engine(h, "sqt_push_frame", rc);
			if (rc == SQT_FLUSH) { /* PrintProcess from AccumSpacetimeCorr, :438-440 */
				rc = sqt_multi_fetch_results(h, &res);
				if (rc) die_engine(h, "sqt_fetch_results", rc);
				if (sqt_host_print_process_x(NULL, &res, &deck, &fl, NULL) || sqt_host_print_types(NULL, &res, &deck, &fl, NULL)) { perror("PrintProcess"); exit(1); }
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
		sqt_prefetch_stop(pf);
		fprintf(stderr, "\nEnd : file : %s\n", argv[a]);
	}
	rc = sqt_multi_finish(h, &res);
	if (rc < 0) die_engine(h, "sqt_finish", rc);
	const double t_loop = now_s() - t_loop0; /* second pass over the files: read, copy, compute, until the last result is back */
	if (rc == 1) { /* lmp2sqt.c:327-333 */
		fprintf(stderr, "limit corr  = %d, countCorrAv = %d\n", limitCorrAv, res.countCorrAv);
		if (sqt_host_print_process_x(NULL, &res, &deck, &fl, NULL) || sqt_host_print_types(NULL, &res, &deck, &fl, NULL)) { perror("PrintProcess"); exit(1); }
	}
	fprintf(stderr, "engine: %d GPU%s; %s; %lld kernel launches; frame loop %.3f s = %.1f frames/s\n", sqt_multi_n_gpus(h),
	        sqt_multi_n_gpus(h) > 1 ? "s, particle-sharded" : "", sqt_kernel_config(sqt_multi_handle(h, 0)), sqt_multi_kernel_launches(h),
	        t_loop, full_n_snaps_index / (t_loop > 0 ? t_loop : 1));
	sqt_results_free(&res);
	sqt_multi_destroy(h);
	for (int a = optind; a < argc; a++) sqt_reader_close(readers[a]);
	free(readers);
	free(files_on);
	free(num_data);
	return 0;
}
