/*
 * analysis_text_main.c -- command-line drop-in for the reference's text post-processor `analysis.out` (analysis.c:138-506;
 * AverageFqt2FqOrFt.c is the same program): same options, reads `in.lmp2sqt` from the working directory and the output.txt
 * named on the command line ("-" = stdin), writes the same files into the working directory and the same lines to stdout.
 */
#include "../../../include/sqt_host.h"

#include <errno.h>
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void usage(const char *prog, int line)
{
	printf("Usage: %s [-t{time_corr}] [-s{skip}n] [-w{window}] input-file \n"
	       "\t--time -t    : time space(default) \n"
	       "\t--fourier -f : omega space  \n"
	       "\t--window -w  : do windows \n"
	       "\t--nskip -s (with option int): Skip early data\n"
	       "\t--ndiff -d (with option int): \n"
	       "\t--help -h    : usage of this function\n"
	       "\t--verbose -v : equaivalent with help \n"
	       " if you want to use stdin, you should used -  \n"
	       "Error line number %d\n", prog, line);
	exit(0);
}

int main(int argc, char **argv)
{
	int do_fourier = 0, do_window = 0, n_skip = -1, n_diff = 4, verbose = 0; /* defaults: analysis.c:166-184 */
	static struct option long_options[] = {
		{ "verbose", no_argument, 0, 'v' }, { "time", no_argument, 0, 't' }, { "fourier", no_argument, 0, 'f' }, { "window", no_argument, 0, 'w' },
		{ "nskip", required_argument, 0, 's' }, { "ndiff", required_argument, 0, 'd' }, { "ulimit", required_argument, 0, 'u' },
		{ "help", no_argument, 0, 'h' }, { 0, 0, 0, 0 } };
	if (argc < 2 || !strcmp(argv[1], "-h")) usage(argv[0], __LINE__);
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
		case 'u': break; /* parsed by the reference, never used by its loops (the decay test is a constant 0.1, :409) */
		case 'v': verbose = 1; break;
		default: usage(argv[0], __LINE__);
		}
	}
	(void)verbose;
	if (optind >= argc) usage(argv[0], __LINE__);
	sqt_analysis a;
	int rc = sqt_analysis_text_load(&a, "in.lmp2sqt", argv[optind], n_skip);
	if (rc) {
		if (rc == -3) fprintf(stderr, "Unable to open '%s': %s\n", argv[optind], strerror(errno));
		else fprintf(stderr, rc == -1 ? "cannot read in.lmp2sqt\n" : "unexpected end of '%s' inside a block\n", argv[optind]);
		return EXIT_FAILURE;
	}
	printf("%d\n", a.nData);
	rc = do_fourier ? sqt_analysis_text_fourier(&a, do_window) : sqt_analysis_text_time(&a, n_diff, NULL);
	if (rc) {
		fprintf(stderr, do_fourier ? "2 (nCTime - 1) must be a power of two for the transform\n" : "time-domain analysis failed (%d)\n", rc);
		return EXIT_FAILURE;
	}
	rc = sqt_analysis_text_print(&a, NULL);
	sqt_analysis_free(&a);
	return rc ? EXIT_FAILURE : 0;
}
