/*
 * mc_gen.c -- synthetic genotype generator (bench + tests).
 *
 * Replaces the reference's --simulate path for workload generation
 * (multiclust.c:167-186 is biallelic-only and cannot make BASELINE.json's
 * multi-allelic / missing / polyploid configs; SURVEY.md finding 4).  Writes
 * STRUCTURE text the reference parser accepts (read_file.c:38-300) and/or the
 * MCB1 container (include/mc_format.h).  The MCB output is recoded the way the
 * reference recodes text input: alleles that were never drawn get no slot,
 * loci with a missing copy get the phantom slot (read_file.c:527-530).
 *
 * usage: mc_gen --I n --L n [--P 2] [--K 3] [--jmax 5] [--miss 0] [--seed s]
 *               [--npops 3] [--stru FILE] [--interleaved] [--mcb FILE]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mc_format.h"
#include "mc_synth.h"

/* allele label written to text for nominal allele j at locus l */
static int label_of(int l, int j)
{
	return 101 + 3 * j + (l % 4);
}

int main(int argc, char **argv)
{
	mcs_params g = { 20261018ULL, 3, 5, 0, 2 };
	int I = 0, L = 0, npops = 3, interleaved = 0, a, i, l, c, j;
	const char *stru = NULL, *mcb = NULL;
	uint8_t *raw;
	mcb_data d;

	for (a = 1; a < argc; a++) {
		if (!strcmp(argv[a], "--I") && a + 1 < argc) I = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--L") && a + 1 < argc) L = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--P") && a + 1 < argc) g.ploidy = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--K") && a + 1 < argc) g.K = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--jmax") && a + 1 < argc) g.jmax = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--miss") && a + 1 < argc) g.miss_bp = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--seed") && a + 1 < argc) g.seed = strtoull(argv[++a], NULL, 0);
		else if (!strcmp(argv[a], "--npops") && a + 1 < argc) npops = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--stru") && a + 1 < argc) stru = argv[++a];
		else if (!strcmp(argv[a], "--mcb") && a + 1 < argc) mcb = argv[++a];
		else if (!strcmp(argv[a], "--interleaved")) interleaved = 1;
		else {
			fprintf(stderr, "mc_gen: bad option '%s'\n", argv[a]);
			return 2;
		}
	}
	if (I <= 0 || L <= 0 || g.ploidy <= 0 || g.jmax < 2 || g.jmax > 254
		|| g.K < 1 || npops < 1) {
		fprintf(stderr, "mc_gen: need --I and --L (> 0), 2 <= jmax <= 254\n");
		return 2;
	}

	raw = malloc((size_t)I * L * g.ploidy);
	if (!raw)
		return 3;
	for (i = 0; i < I; i++)
		for (l = 0; l < L; l++)
			for (c = 0; c < g.ploidy; c++)
				raw[((size_t)i * L + l) * g.ploidy + c]
					= mcs_code(&g, i, l, c);

	if (stru) {
		FILE *fp = fopen(stru, "w");
		if (!fp) {
			perror(stru);
			return 3;
		}
		for (l = 0; l < L; l++)
			fprintf(fp, "%sloc%d", l ? " " : "", l + 1);
		fprintf(fp, "\n");
		for (i = 0; i < I; i++) {
			if (interleaved) {
				fprintf(fp, "ind%d pop%d", i, i % npops);
				for (l = 0; l < L; l++)
					for (c = 0; c < g.ploidy; c++) {
						int r = raw[((size_t)i * L + l) * g.ploidy + c];
						fprintf(fp, " %d", r == MC_MISSING_CODE
							? -9 : label_of(l, r));
					}
				fprintf(fp, "\n");
			} else {
				for (c = 0; c < g.ploidy; c++) {
					fprintf(fp, "ind%d pop%d", i, i % npops);
					for (l = 0; l < L; l++) {
						int r = raw[((size_t)i * L + l) * g.ploidy + c];
						fprintf(fp, " %d", r == MC_MISSING_CODE
							? -9 : label_of(l, r));
					}
					fprintf(fp, "\n");
				}
			}
		}
		fclose(fp);
	}

	if (mcb) {
		int64_t nlab = 0;
		int *map = malloc(sizeof(int) * 256);

		memset(&d, 0, sizeof d);
		d.I = I; d.L = L; d.P = g.ploidy; d.npops = npops;
		d.J = calloc((size_t)L, 4);
		d.nreal = calloc((size_t)L, 4);
		d.locale = malloc(4 * (size_t)I);
		d.labels = malloc(4 * (size_t)L * 256);
		d.codes = raw;
		for (i = 0; i < I; i++)
			d.locale[i] = i % npops;
		for (l = 0; l < L; l++) {
			int seen[256] = { 0 }, miss = 0, n = 0;
			for (i = 0; i < I; i++)
				for (c = 0; c < g.ploidy; c++) {
					int r = raw[((size_t)i * L + l) * g.ploidy + c];
					if (r == MC_MISSING_CODE)
						miss = 1;
					else
						seen[r] = 1;
				}
			for (j = 0; j < 255; j++) {
				map[j] = -1;
				if (seen[j]) {
					map[j] = n++;
					d.labels[nlab++] = label_of(l, j);
				}
			}
			d.nreal[l] = n;
			d.J[l] = n ? n + miss : 0;	/* all-missing locus: 0 */
			for (i = 0; i < I; i++)
				for (c = 0; c < g.ploidy; c++) {
					size_t x = ((size_t)i * L + l) * g.ploidy + c;
					if (raw[x] != MC_MISSING_CODE)
						raw[x] = (uint8_t)map[raw[x]];
				}
		}
		if (mcb_write(mcb, &d)) {
			perror(mcb);
			return 3;
		}
		free(map);
		free(d.J); free(d.nreal); free(d.locale); free(d.labels);
	}
	free(raw);
	return 0;
}
