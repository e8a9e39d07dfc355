/*
 * analysis_main.c -- command-line drop-in for the reference's `analysis_binary.out` (analysis_binary.c:166-207):
 * same options, reads `in.lmp2sqt` from the working directory and the output.binary named on the command line,
 * writes the same files into the working directory.
 */
#include "../../../include/sqt_host.h"

#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void usage(const char *prog)
{
	printf("Usage: %s [-t{time_corr}] [-s{skip}n] [-w{window}] input-file \n"
	       "\t--time -t    : time space(default) \n"
	       "\t--fourier -f : omega space  \n"
	       "\t--window -w  : do windows \n"
	       "\t--nskip -s (with option int): Skip early data\n"
	       "\t--ndiff -d (with option int): \n"
	       "\t--ulimit -u (with option float:default->exp(-2.5)): \n"
	       "\t--help -h    : usage of this function\n", prog);
	exit(0);
}

int main(int argc, char **argv)
{
	int do_fourier = 0, do_window = 0, n_skip = -1, n_diff = 4; /* defaults: analysis_binary.c:410-422 */
	double under_limit = exp(-2.5);
	static struct option long_options[] = {
		{ "time", no_argument, 0, 't' }, { "fourier", no_argument, 0, 'f' }, { "window", no_argument, 0, 'w' },
		{ "nskip", required_argument, 0, 's' }, { "ndiff", required_argument, 0, 'd' }, { "ulimit", required_argument, 0, 'u' },
		{ "help", no_argument, 0, 'h' }, { 0, 0, 0, 0 } };
	if (argc < 2) usage(argv[0]);
	opterr = 0;
	for (;;) {
		int c = getopt_long(argc, argv, "ftws:d:u:h", long_options, NULL);
		if (c == -1) break;
		switch (c) {
		case 't': do_fourier = 0; break;
		case 'f': do_fourier = 1; break;
		case 'w': do_window = 1; break;
		case 's': n_skip = atoi(optarg); break;
		case 'd': n_diff = atoi(optarg); break;
		case 'u': under_limit = atof(optarg); break;
		default: usage(argv[0]);
		}
	}
	if (optind >= argc) usage(argv[0]);
	sqt_analysis a;
	int rc = sqt_analysis_load(&a, "in.lmp2sqt", argv[optind], n_skip);
	if (rc) {
		fprintf(stderr, rc == -1 ? "cannot read in.lmp2sqt\n" : rc == -3 ? "Unable to open '%s'\n" : "File error in '%s' (header or block layout)\n", argv[optind]);
		return EXIT_FAILURE;
	}
	printf("%d\n", a.nData);
	rc = do_fourier ? sqt_analysis_fourier(&a, do_window) : sqt_analysis_time(&a, n_diff, under_limit, NULL);
	if (rc) { fprintf(stderr, do_fourier ? "2 (nCTime - 1) must be a power of two for the transform\n" : "time-domain analysis failed (%d)\n", rc); return EXIT_FAILURE; }
	if (!do_fourier) printf("nv = %d, tMax = %.13f, nCTime = %d, deltaTCorr = %.13f\n", a.nv, a.tMax, a.nCTime, a.deck_deltaT);
	rc = sqt_analysis_print(&a, NULL);
	sqt_analysis_free(&a);
	return rc ? EXIT_FAILURE : 0;
}
