/*
 * read_data.c -- STRUCTURE reader and allele recoding (reference
 * read_file.c:38-300, 411-429, 443-663), producing 8-bit allele codes.
 *
 * Same observable result as the reference parser, without its two bubble
 * sorts per locus (read_file.c:518, 577 are O(n^2) in the haplotype count):
 *   - stacked vs interleaved rows are told apart by the first two names
 *     (read_file.c:84-95);
 *   - the alleles of a locus are recoded 0..n-1 in ascending label order
 *     (572-588);
 *   - a locus with a missing copy gets one extra, unlabelled allele slot that
 *     no copy is ever counted in (527-530 vs 580-585; SURVEY.md finding 3);
 *     an all-missing locus gets no slot at all (525-526);
 *   - missing copies are counted nowhere (651-657): code 255.
 * A file whose name ends in ".mcb" is read as the binary container of
 * include/mc_format.h instead (already recoded).
 */
#include <ctype.h>
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#include "mc_format.h"
#include "multiclust.h"

/* next white-space delimited word, newly allocated; NULL at end of file */
static char *next_word(FILE *fp)
{
	size_t n = 0, cap = 32;
	char *w;
	int c;

	do {
		c = fgetc(fp);
	} while (c == ' ' || c == '\t' || c == '\n' || c == '\r');
	if (c == EOF)
		return NULL;
	w = malloc(cap);
	while (c != EOF && !isspace(c)) {
		if (n + 2 > cap)
			w = realloc(w, cap *= 2);
		w[n++] = (char)c;
		c = fgetc(fp);
	}
	if (c != EOF)
		ungetc(c, fp);
	w[n] = 0;
	return w;
}

static void skip_line(FILE *fp)
{
	int c;
	do {
		c = fgetc(fp);
	} while (c != EOF && c != '\n');
}

/* columns on the rest of the current line; leaves fp on the next line */
static int count_columns(FILE *fp)
{
	int n = 0, in_word = 0, c;

	while ((c = fgetc(fp)) != EOF && c != '\n') {
		if (c == ' ' || c == '\t' || c == '\r') {
			in_word = 0;
		} else if (!in_word) {
			in_word = 1;
			n++;
		}
	}
	return n;
}

/* non-empty lines from fp to the end of the file */
static int count_lines(FILE *fp)
{
	int n = 0, blank = 1, c;

	while ((c = fgetc(fp)) != EOF) {
		if (c == '\n') {
			n += !blank;
			blank = 1;
		} else if (!isspace(c)) {
			blank = 0;
		}
	}
	return n + !blank;
}

static int locale_index(data *dat, const char *name)
{
	for (int n = 0; n < dat->numpops; n++)
		if (!strcmp(dat->pops[n], name))
			return n;
	dat->pops = realloc(dat->pops, sizeof *dat->pops * ((size_t)dat->numpops + 1));
	dat->pops[dat->numpops] = strdup(name);
	return dat->numpops++;
}

static int cmp_int(const void *a, const void *b)
{
	const int x = *(const int *)a, y = *(const int *)b;
	return (x > y) - (x < y);
}

/* recode one locus: raw[h] for h < nhap (stride `stride`) -> codes */
static int recode_locus(data *dat, int l, const int *raw, size_t stride, int nhap,
	int *scratch, int32_t **labels, int64_t *nlab, int64_t *caplab)
{
	int n = 0, miss = 0, nreal = 0;

	for (int h = 0; h < nhap; h++) {
		const int v = raw[(size_t)h * stride];
		if (v == MISSING)
			miss = 1;
		else
			scratch[n++] = v;
	}
	qsort(scratch, (size_t)n, sizeof *scratch, cmp_int);
	for (int x = 0; x < n; x++)
		if (!x || scratch[x] != scratch[x - 1])
			scratch[nreal++] = scratch[x];
	if (nreal > 254)
		return mmessage(ERROR_MSG, INVALID_USER_SETUP, "locus %d has %d "
			"distinct alleles; 8-bit allele codes allow 254\n", l + 1, nreal);
	if (*nlab + nreal > *caplab) {
		*caplab = 2 * (*caplab + nreal);
		*labels = realloc(*labels, sizeof **labels * (size_t)*caplab);
	}
	memcpy(*labels + *nlab, scratch, sizeof *scratch * (size_t)nreal);
	dat->label_off[l] = *nlab;
	*nlab += nreal;
	dat->nreal[l] = nreal;
	dat->uniquealleles[l] = nreal ? nreal + miss : 0;
	if (miss && nreal)
		dat->missing_data = 1;
	return NO_ERROR;
}

static int code_of(const int32_t *labels, int n, int v)
{
	int lo = 0, hi = n - 1;

	while (lo <= hi) {
		const int mid = (lo + hi) / 2;
		if (labels[mid] == v)
			return mid;
		if (labels[mid] < v)
			lo = mid + 1;
		else
			hi = mid - 1;
	}
	return MC_CODE_MISSING;
}

static void finish_dims(data *dat)
{
	dat->M = 0;
	dat->allele_off[0] = 0;
	for (int l = 0; l < dat->L; l++) {
		dat->allele_off[l + 1] = dat->allele_off[l] + dat->uniquealleles[l];
		if (dat->uniquealleles[l] > dat->M)
			dat->M = dat->uniquealleles[l];
	}
	dat->i_p = calloc((size_t)dat->numpops, sizeof *dat->i_p);
	for (int i = 0; i < dat->I; i++)
		dat->i_p[dat->idv[i].locale]++;
	dat->I_K = calloc((size_t)dat->I, sizeof *dat->I_K);
}

static int read_mcb_file(options *opt, data *dat)
{
	mcb_data d;
	char buf[64];

	if (mcb_read(opt->filename, &d))
		return message(stderr, __FILE__, __func__, __LINE__, ERROR_MSG,
			FILE_OPEN_ERROR, opt->filename);
	dat->I = d.I; dat->L = d.L; dat->ploidy = d.P;
	dat->uniquealleles = d.J; dat->nreal = d.nreal; dat->labels = d.labels;
	dat->label_off = d.lab_off; dat->codes = d.codes;
	dat->allele_off = malloc(sizeof(int32_t) * ((size_t)d.L + 1));
	dat->idv = malloc(sizeof *dat->idv * (size_t)d.I);
	for (int n = 0; n < d.npops; n++) {
		snprintf(buf, sizeof buf, "pop%d", n);
		locale_index(dat, buf);
	}
	for (int i = 0; i < d.I; i++) {
		snprintf(buf, sizeof buf, "ind%d", i);
		dat->idv[i].name = strdup(buf);
		dat->idv[i].locale = d.locale[i];
	}
	for (int l = 0; l < d.L; l++)
		if (d.J[l] > d.nreal[l])
			dat->missing_data = 1;
	free(d.locale);
	finish_dims(dat);
	return NO_ERROR;
}

int read_file(options *opt, data *dat)
{
	const size_t flen = strlen(opt->filename);
	FILE *fp;
	char *name1, *name2, *word;
	int skip_line_two = 0, ncol, nhap, *raw = NULL, *scratch, err = NO_ERROR;
	int32_t *labels = NULL;
	int64_t nlab = 0, caplab = 0;

	if (flen > 4 && !strcmp(opt->filename + flen - 4, ".mcb"))
		return read_mcb_file(opt, dat);

	if (!(fp = fopen(opt->filename, "r")))
		return message(stderr, __FILE__, __func__, __LINE__, ERROR_MSG,
			FILE_OPEN_ERROR, opt->filename);

	/* header: one name per locus (or per column) */
	dat->L = count_columns(fp);
	if (opt->R_format)
		dat->L -= 2;
	if (!(name1 = next_word(fp)))
		return mmessage(ERROR_MSG, END_OF_FILE, opt->filename);
	if (!strcmp(name1, "-1")) {	/* inter-marker distances: ignored */
		skip_line_two = 1;
		skip_line(fp);
		free(name1);
		if (!(name1 = next_word(fp)))
			return mmessage(ERROR_MSG, END_OF_FILE, opt->filename);
	}
	skip_line(fp);
	if (!(name2 = next_word(fp)))
		return mmessage(ERROR_MSG, END_OF_FILE, opt->filename);
	if (strcmp(name1, name2))
		opt->interleaved = 1;
	free(name1);
	free(name2);
	ncol = count_columns(fp) - 1;	/* minus the locale column */

	if (opt->interleaved && ncol != dat->L && ncol != dat->ploidy * dat->L)
		return mmessage(ERROR_MSG, FILE_FORMAT_ERROR, "number of columns (%u) "
			"in '%s' is not a multiple of ploidy (%d)\n", dat->L,
			opt->filename, dat->ploidy);
	if (!opt->interleaved && ncol != dat->L)
		return mmessage(ERROR_MSG, FILE_FORMAT_ERROR, "number of columns (%u) "
			"in '%s' does not match number of alleles (%d) given for "
			"first individual\n", dat->L, opt->filename, ncol);
	if (opt->interleaved && ncol == dat->L)
		dat->L /= dat->ploidy;

	/* same count as the reference, including its one-short count when the
	 * inter-marker distance line is present (read_file.c:121) */
	dat->I = count_lines(fp) + 2 - skip_line_two;
	if (!opt->interleaved && dat->I % dat->ploidy)
		return mmessage(ERROR_MSG, FILE_FORMAT_ERROR, "number of lines (%d) in "
			"'%s' is not a multiple of ploidy (%d)\n", dat->I,
			opt->filename, dat->ploidy);
	if (opt->interleaved) {
		nhap = dat->I * dat->ploidy;
	} else {
		nhap = dat->I;
		dat->I /= dat->ploidy;
	}
	if (dat->ploidy > 16)
		return mmessage(ERROR_MSG, INVALID_USER_SETUP, "ploidy %d exceeds the "
			"16 copies per locus the device layout holds\n", dat->ploidy);

	/* raw alleles, haplotype-major like dat->IL */
	raw = malloc(sizeof *raw * (size_t)nhap * dat->L);
	dat->idv = calloc((size_t)dat->I, sizeof *dat->idv);
	if (!raw || !dat->idv)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "genotype table\n");
	rewind(fp);
	skip_line(fp);
	if (skip_line_two)
		skip_line(fp);
	for (int h = 0, idv = 0; h < nhap; h += opt->interleaved ? dat->ploidy : 1) {
		const int rows = opt->interleaved ? dat->ploidy : 1;
		if (!(word = next_word(fp)))
			break;
		if (opt->interleaved || !(h % dat->ploidy)) {
			dat->idv[idv].name = word;
			if (!(word = next_word(fp)))
				return mmessage(ERROR_MSG, END_OF_FILE, opt->filename);
			dat->idv[idv].locale = locale_index(dat, word);
			free(word);
			idv++;
		} else {	/* repeated name and locale of a stacked row */
			free(word);
			free(next_word(fp));
		}
		for (int l = 0; l < dat->L; l++)
			for (int j = 0; j < rows; j++) {
				int v;
				if (fscanf(fp, "%d", &v) != 1)
					return mmessage(ERROR_MSG, FILE_FORMAT_ERROR,
						"failed to read locus %d of haplotype "
						"%d in file '%s'.  Check option -R.\n",
						l + 1, h + j + 1, opt->filename);
				raw[(size_t)(h + j) * dat->L + l] = v;
			}
	}
	fclose(fp);

	/* --missing: remap to the default marker (read_file.c:411-429) */
	if (opt->missing_value != MISSING)
		for (size_t x = 0; x < (size_t)nhap * dat->L; x++) {
			if (raw[x] == MISSING)
				return mmessage(ERROR_MSG, INVALID_USER_SETUP, "The "
					"default missing value (%d) is observed in the "
					"input file, but the user has defined the "
					"missing value to be %d.\n", MISSING,
					opt->missing_value);
			if (raw[x] == opt->missing_value)
				raw[x] = MISSING;
		}

	/* recode */
	dat->uniquealleles = calloc((size_t)dat->L, sizeof(int32_t));
	dat->nreal = calloc((size_t)dat->L, sizeof(int32_t));
	dat->allele_off = calloc((size_t)dat->L + 1, sizeof(int32_t));
	dat->label_off = calloc((size_t)dat->L + 1, sizeof(int64_t));
	dat->codes = malloc((size_t)dat->I * dat->L * dat->ploidy);
	scratch = malloc(sizeof *scratch * (size_t)nhap);
	for (int l = 0; l < dat->L && !err; l++)
		err = recode_locus(dat, l, raw + l, (size_t)dat->L, nhap, scratch,
			&labels, &nlab, &caplab);
	if (err)
		return err;
	dat->label_off[dat->L] = nlab;
	dat->labels = labels;
	for (int i = 0; i < dat->I; i++)
		for (int l = 0; l < dat->L; l++)
			for (int a = 0; a < dat->ploidy; a++) {
				const int v = raw[(size_t)(i * dat->ploidy + a) * dat->L + l];
				dat->codes[((size_t)i * dat->L + l) * dat->ploidy + a]
					= v == MISSING ? MC_CODE_MISSING
					: (uint8_t)code_of(labels + dat->label_off[l],
						dat->nreal[l], v);
			}
	free(scratch);
	free(raw);
	finish_dims(dat);
	return NO_ERROR;
}

/* CUDA contexts take 1-3 s to come up; they are created by a helper thread
 * while the main thread reads and recodes the data file */
static struct {
	pthread_t thread;
	int started, n, device, rc[64];
	mc_ctx *ctx[64];
	double seconds;
} g_early;

static void *early_main(void *arg)
{
	struct timespec t0, t1;

	(void)arg;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int r = 0; r < g_early.n; r++)
		g_early.rc[r] = mc_create(&g_early.ctx[r], g_early.device + r);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	g_early.seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
	return NULL;
}

void start_device_contexts(options *opt)
{
	if (opt->parse_only || opt->n_gpus > 64 || (opt->shard_fits && opt->n_gpus > 1))
		return;
	g_early.n = opt->n_gpus;
	g_early.device = opt->device;
	g_early.started = !pthread_create(&g_early.thread, NULL, early_main, NULL);
}

/* hand the recoded genotypes to the device(s): with --gpus N device r gets
 * the individuals [r*I/N, (r+1)*I/N) and the allele slots of the whole sample */
int upload_data(options *opt, data *dat, model *mod)
{
	const int n = opt->n_gpus;
	int rc;

	mod->n_gpus = n;
	mod->gpus = calloc((size_t)n, sizeof *mod->gpus);
	mod->row_first = calloc((size_t)n + 1, sizeof *mod->row_first);
	if (!mod->gpus || !mod->row_first)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "device table\n");
	if (dat->I < n)
		return mmessage(ERROR_MSG, INVALID_USER_SETUP, "--gpus %d exceeds the "
			"number of individuals (%d)\n", n, dat->I);
	for (int r = 0; r <= n; r++)
		mod->row_first[r] = (int)((long long)dat->I * r / n);
	if (g_early.started) {
		pthread_join(g_early.thread, NULL);
		g_early.started = 0;
		if (opt->timing)
			fprintf(stderr, "timing (s): %d device context(s) %.3f, overlapped with "
				"reading the data\n", g_early.n, g_early.seconds);
	}
	for (int r = 0; r < n; r++) {
		const int rows = mod->row_first[r + 1] - mod->row_first[r];
		if (r < g_early.n && g_early.ctx[r]) {
			mod->gpus[r] = g_early.ctx[r];
			g_early.ctx[r] = NULL;
		} else if ((rc = mc_create(&mod->gpus[r], opt->device + r))) {
			return mmessage(ERROR_MSG, GPU_ERROR, "%s\n", mc_last_error(NULL));
		}
		if ((rc = mc_set_data(mod->gpus[r], rows, dat->L, dat->ploidy,
			dat->uniquealleles, dat->codes
			+ (size_t)mod->row_first[r] * dat->L * dat->ploidy)))
			return mmessage(ERROR_MSG, GPU_ERROR, "%s\n",
				mc_last_error(mod->gpus[r]));
	}
	mod->gpu = mod->gpus[0];
	mod->T = dat->allele_off[dat->L];
	return NO_ERROR;
}
