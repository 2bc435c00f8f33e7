/*
 * mc_format.h -- "MCB1" binary genotype container (host C, header only).
 *
 * The reference only reads STRUCTURE text (read_file.c:38-300) and recodes
 * alleles with two bubble sorts per locus (read_file.c:518,577), which is
 * O((I*ploidy)^2 * L) and unusable at bench sizes (SURVEY.md finding 5).
 * This container holds the *result* of that recoding -- the flat layout the
 * CUDA path consumes -- so that large synthetic workloads can be handed to
 * the reference harness, the oracle and the GPU path without the text parser.
 *
 * Layout (little endian):
 *   char     magic[4] = "MCB1"
 *   int32    I, L, P, npops
 *   int32    J[L]        allele slots per locus, INCLUDING the reference's
 *                        phantom slot for loci with missing data
 *                        (read_file.c:527-530; SURVEY.md finding 3)
 *   int32    nreal[L]    real (labelled) alleles per locus (read_file.c:580-585)
 *   int32    labels[sum nreal]   ascending allele labels per locus
 *   int32    locale[I]   sampling locale index of each individual
 *   uint8    codes[I][L][P]      0..nreal-1, 255 = missing
 */
#ifndef MC_FORMAT_H
#define MC_FORMAT_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MCB_MISSING 255	/* allele code of a missing copy */

typedef struct {
	int32_t I, L, P, npops;
	int32_t *J;       /* [L] */
	int32_t *nreal;   /* [L] */
	int32_t *labels;  /* [sum nreal] */
	int64_t *lab_off; /* [L+1] prefix sums of nreal */
	int32_t *locale;  /* [I] */
	uint8_t *codes;   /* [I][L][P] */
} mcb_data;

static inline void mcb_free(mcb_data *d)
{
	free(d->J);
	free(d->nreal);
	free(d->labels);
	free(d->lab_off);
	free(d->locale);
	free(d->codes);
	memset(d, 0, sizeof *d);
}

static inline int mcb_write(const char *path, const mcb_data *d)
{
	FILE *fp = fopen(path, "wb");
	int64_t nlab = 0;
	int l;

	if (!fp)
		return -1;
	for (l = 0; l < d->L; l++)
		nlab += d->nreal[l];
	fwrite("MCB1", 1, 4, fp);
	fwrite(&d->I, 4, 1, fp);
	fwrite(&d->L, 4, 1, fp);
	fwrite(&d->P, 4, 1, fp);
	fwrite(&d->npops, 4, 1, fp);
	fwrite(d->J, 4, (size_t)d->L, fp);
	fwrite(d->nreal, 4, (size_t)d->L, fp);
	fwrite(d->labels, 4, (size_t)nlab, fp);
	fwrite(d->locale, 4, (size_t)d->I, fp);
	fwrite(d->codes, 1, (size_t)d->I * d->L * d->P, fp);
	return fclose(fp);
}

static inline int mcb_read(const char *path, mcb_data *d)
{
	FILE *fp = fopen(path, "rb");
	char magic[4];
	int64_t nlab = 0;
	size_t n;
	int l;

	memset(d, 0, sizeof *d);
	if (!fp)
		return -1;
	if (fread(magic, 1, 4, fp) != 4 || memcmp(magic, "MCB1", 4))
		goto BAD;
	if (fread(&d->I, 4, 1, fp) != 1 || fread(&d->L, 4, 1, fp) != 1
		|| fread(&d->P, 4, 1, fp) != 1 || fread(&d->npops, 4, 1, fp) != 1)
		goto BAD;
	if (d->I <= 0 || d->L <= 0 || d->P <= 0)
		goto BAD;
	d->J = malloc(sizeof(int32_t) * (size_t)d->L);
	d->nreal = malloc(sizeof(int32_t) * (size_t)d->L);
	d->lab_off = malloc(sizeof(int64_t) * ((size_t)d->L + 1));
	d->locale = malloc(sizeof(int32_t) * (size_t)d->I);
	if (!d->J || !d->nreal || !d->lab_off || !d->locale)
		goto BAD;
	if (fread(d->J, 4, (size_t)d->L, fp) != (size_t)d->L
		|| fread(d->nreal, 4, (size_t)d->L, fp) != (size_t)d->L)
		goto BAD;
	d->lab_off[0] = 0;
	for (l = 0; l < d->L; l++) {
		nlab += d->nreal[l];
		d->lab_off[l + 1] = nlab;
	}
	d->labels = malloc(sizeof(int32_t) * (size_t)(nlab ? nlab : 1));
	n = (size_t)d->I * d->L * d->P;
	d->codes = malloc(n);
	if (!d->labels || !d->codes)
		goto BAD;
	if (fread(d->labels, 4, (size_t)nlab, fp) != (size_t)nlab
		|| fread(d->locale, 4, (size_t)d->I, fp) != (size_t)d->I
		|| fread(d->codes, 1, n, fp) != n)
		goto BAD;
	fclose(fp);
	return 0;
BAD:
	fclose(fp);
	mcb_free(d);
	return -2;
}

#endif /* MC_FORMAT_H */
