/*
 * sqt_deck.c -- the NameList deck (in.lmp2sqt.*), same syntax and same rewrite side effect as the
 * reference (GetNameList lmp2sqt.c:2014-2111, PrintNameList2File :2126-2169, UpdateNameList
 * :2112-2125).  Re-parsed from the format description, not ported: a table of (name, type,
 * offset) instead of the reference's macro machinery.
 */
#include "../../../include/sqt_host.h"

#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct { const char *name; int is_real; size_t off; } deck_entry;

/* DATALIST order, lmp2sqt.c:82-92 */
static const deck_entry k_entries[SQT_DECK_NNAMES] = {
	{ "kVal", 1, offsetof(sqt_deck, kVal) },
	{ "deltaT", 1, offsetof(sqt_deck, deltaT) },
	{ "mass", 1, offsetof(sqt_deck, mass) },
	{ "rVal", 1, offsetof(sqt_deck, rVal) },
	{ "limitCorrAv", 0, offsetof(sqt_deck, limitCorrAv) },
	{ "nCBuffer", 0, offsetof(sqt_deck, nCBuffer) },
	{ "nCSpatial", 0, offsetof(sqt_deck, nCSpatial) },
	{ "nCTime", 0, offsetof(sqt_deck, nCTime) },
	{ "nCSkip", 0, offsetof(sqt_deck, nCSkip) },
};

int sqt_deck_write_template(const char *path)
{
	FILE *fp = fopen(path, "w");
	if (!fp) return -1;
	for (int k = 0; k < SQT_DECK_NNAMES; k++) {
		fprintf(fp, "%s\t", k_entries[k].name);
		if (strlen(k_entries[k].name) < 8) fprintf(fp, "\t");
		if (k_entries[k].is_real) fprintf(fp, "%#g ", 0.00);
		else fprintf(fp, "%d ", 0);
		fprintf(fp, "\n");
	}
	fprintf(fp, "----\n");
	fclose(fp);
	return 0;
}

int sqt_deck_read(const char *path, sqt_deck *d)
{
	char buff[100];
	FILE *fp = fopen(path, "r");
	memset(d, 0, sizeof(*d));
	d->nCSkip = 0; /* default, lmp2sqt.c:2032 */
	if (!fp) return -1;
	int ok = 1;
	for (;;) {
		/* one record per line of at most 79 characters, name then value(s) */
		if (!fgets(buff, 80, fp)) break;
		if (feof(fp)) break; /* like the reference, a last line without newline is dropped (:2063-2064) */
		char *token = strtok(buff, " \t\n");
		if (!token) break; /* blank line terminates the deck */
		int match = 0;
		for (int k = 0; k < SQT_DECK_NNAMES; k++) {
			if (strcmp(token, k_entries[k].name) != 0) continue;
			match = 1;
			if (d->status[k] != 0) { d->status[k] = 4; ok = 0; break; }
			d->status[k] = 1;
			token = strtok(NULL, ", \t\n");
			if (token) {
				if (k_entries[k].is_real) *(double *)((char *)d + k_entries[k].off) = atof(token);
				else *(int *)((char *)d + k_entries[k].off) = (int)atol(token);
			} else { d->status[k] = 2; ok = 0; }
			token = strtok(NULL, ", \t\n");
			if (token) { d->status[k] = 3; ok = 0; }
			break;
		}
		if (!match) ok = 0;
	}
	fclose(fp);
	if (d->nCBuffer > d->nCTime) d->nCBuffer = d->nCTime; /* :2105 */
	for (int k = 0; k < SQT_DECK_NNAMES; k++)
		if (d->status[k] != 1) ok = 0;
	return ok;
}

int sqt_deck_print(FILE *fp, const sqt_deck *d)
{
	static const char *notes[5] = { "** no data", "", "** missing data", "** extra data", "** multiply defined" };
	fprintf(fp, "NameList -- data\n");
	for (int k = 0; k < SQT_DECK_NNAMES; k++) {
		fprintf(fp, "%s\t", k_entries[k].name);
		if (strlen(k_entries[k].name) < 8) fprintf(fp, "\t");
		if (d->status[k] > 0) {
			if (k_entries[k].is_real) fprintf(fp, "%#g ", *(const double *)((const char *)d + k_entries[k].off));
			else fprintf(fp, "%d ", *(const int *)((const char *)d + k_entries[k].off));
		}
		fputs(notes[d->status[k] >= 0 && d->status[k] <= 4 ? d->status[k] : 0], fp);
		fprintf(fp, "\n");
	}
	fprintf(fp, "----\n");
	return 0;
}

int sqt_deck_write(const char *path, const sqt_deck *d)
{
	FILE *fp = fopen(path, "w");
	if (!fp) return -1;
	sqt_deck_print(fp, d);
	fclose(fp);
	return 0;
}
