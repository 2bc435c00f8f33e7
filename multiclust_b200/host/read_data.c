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

/* the whole text file in memory and a read position: the scanner below does
 * what fgetc / fscanf("%d") did for the reference, a few hundred MB/s faster */
typedef struct {
	const char *p, *end;
} cursor;

static int is_blank(int c)
{
	return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f';
}

/* next white-space delimited word, newly allocated; NULL at end of file */
static char *next_word(cursor *cu)
{
	const char *b;
	char *w;

	while (cu->p < cu->end && is_blank((unsigned char)*cu->p))
		cu->p++;
	if (cu->p >= cu->end)
		return NULL;
	b = cu->p;
	while (cu->p < cu->end && !is_blank((unsigned char)*cu->p))
		cu->p++;
	w = malloc((size_t)(cu->p - b) + 1);
	memcpy(w, b, (size_t)(cu->p - b));
	w[cu->p - b] = 0;
	return w;
}

static void skip_line(cursor *cu)
{
	const char *nl = memchr(cu->p, '\n', (size_t)(cu->end - cu->p));

	cu->p = nl ? nl + 1 : cu->end;
}

/* columns on the rest of the current line; leaves the cursor on the next line */
static int count_columns(cursor *cu)
{
	int n = 0, in_word = 0;

	while (cu->p < cu->end && *cu->p != '\n') {
		const int c = (unsigned char)*cu->p++;
		if (c == ' ' || c == '\t' || c == '\r') {
			in_word = 0;
		} else if (!in_word) {
			in_word = 1;
			n++;
		}
	}
	if (cu->p < cu->end)
		cu->p++;
	return n;
}

/* non-empty lines from the cursor to the end of the file (cursor unchanged) */
static int count_lines(const cursor *cu)
{
	const char *q = cu->p;
	int n = 0;

	while (q < cu->end) {
		const char *nl = memchr(q, '\n', (size_t)(cu->end - q));
		const char *e = nl ? nl : cu->end;
		/* a line counts when it holds anything but white space */
		while (q < e && is_blank((unsigned char)*q))
			q++;
		n += q < e;
		q = e + 1;
	}
	return n;
}

/* fscanf("%d"): skip white space, optional sign, decimal digits */
static int scan_int(cursor *cu, int *out)
{
	const char *q = cu->p;
	long long v = 0;
	int neg = 0, digits = 0;

	while (q < cu->end && is_blank((unsigned char)*q))
		q++;
	if (q < cu->end && (*q == '-' || *q == '+'))
		neg = *q++ == '-';
	while (q < cu->end && *q >= '0' && *q <= '9') {
		v = v * 10 + (*q++ - '0');
		if (v > 4294967296LL)
			v = 4294967296LL;	/* out of range anyway */
		digits++;
	}
	if (!digits)
		return 0;
	cu->p = q;
	*out = (int)(neg ? -v : v);
	return 1;
}

static char *slurp(const char *path, size_t *len)
{
	FILE *fp = fopen(path, "rb");
	char *buf = NULL;
	long long n;

	if (!fp)
		return NULL;
	if (fseeko(fp, 0, SEEK_END) || (n = ftello(fp)) < 0 || fseeko(fp, 0, SEEK_SET)
		|| !(buf = malloc((size_t)n + 1))
		|| fread(buf, 1, (size_t)n, fp) != (size_t)n) {
		free(buf);
		fclose(fp);
		return NULL;
	}
	fclose(fp);
	buf[n] = 0;
	*len = (size_t)n;
	return buf;
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

/* distinct non-missing labels of one locus, in order of first appearance */
typedef struct {
	int *v;
	int n, cap, miss;
} label_set;

/* position of v in the set, appended when new; -1 when memory runs out */
static inline int label_add(label_set *ls, int v)
{
	for (int x = 0; x < ls->n; x++)
		if (ls->v[x] == v)
			return x;
	if (ls->n == ls->cap) {
		ls->cap = ls->cap ? 2 * ls->cap : 8;
		if (!(ls->v = realloc(ls->v, sizeof *ls->v * (size_t)ls->cap)))
			return -1;
	}
	ls->v[ls->n] = v;
	return ls->n++;
}

/* the alleles of every locus recoded 0..n-1 in ascending label order
 * (read_file.c:572-588): one row-major sweep over the raw table collects the
 * distinct labels per locus (a handful each), which are then sorted -- instead
 * of the reference's two bubble sorts over all haplotypes of every locus */
static int recode_loci(data *dat, const int *raw, int nhap, int32_t **labels_out,
	int64_t *nlab_out)
{
	const int L = dat->L, P = dat->ploidy;
	label_set *sets = calloc((size_t)L, sizeof *sets);
	/* first sweep: position of every copy's label in its locus's set */
	const size_t ncell = (size_t)nhap * (size_t)L;
	uint8_t *prov = malloc(ncell ? ncell : 1);
	uint8_t (*rank)[256];
	int32_t *labels;
	int64_t nlab = 0;

	if (!sets || !prov)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "allele labels\n");
	for (int h = 0; h < nhap; h++) {
		const int *row = raw + (size_t)h * L;
		uint8_t *pr = prov + (size_t)h * L;
		for (int l = 0; l < L; l++) {
			if (row[l] == MISSING) {
				sets[l].miss = 1;
				pr[l] = MC_CODE_MISSING;
			} else {
				const int x = sets[l].n < 255 ? label_add(&sets[l], row[l]) : -2;
				if (x == -1)
					return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "allele labels\n");
				if (x == -2 || x > 254)
					return mmessage(ERROR_MSG, INVALID_USER_SETUP, "locus %d has "
						"more than 254 distinct alleles; 8-bit allele codes "
						"allow 254\n", l + 1);
				pr[l] = (uint8_t)x;
			}
		}
	}
	for (int l = 0; l < L; l++)
		nlab += sets[l].n;
	labels = malloc(sizeof *labels * (size_t)(nlab ? nlab : 1));
	rank = malloc(sizeof *rank * ((size_t)L + 1));
	if (!labels || !rank)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "allele labels\n");
	nlab = 0;
	for (int l = 0; l < L; l++) {
		const int nreal = sets[l].n;
		int first[256];

		memcpy(first, sets[l].v, sizeof(int) * (size_t)nreal);
		qsort(sets[l].v, (size_t)nreal, sizeof(int), cmp_int);
		for (int x = 0; x < nreal; x++)		/* first-appearance position -> rank */
			for (int y = 0; y < nreal; y++)
				if (sets[l].v[y] == first[x])
					rank[l][x] = (uint8_t)y;
		rank[l][MC_CODE_MISSING] = MC_CODE_MISSING;
		memcpy(labels + nlab, sets[l].v, sizeof(int) * (size_t)nreal);
		dat->label_off[l] = nlab;
		nlab += nreal;
		dat->nreal[l] = nreal;
		dat->uniquealleles[l] = nreal ? nreal + sets[l].miss : 0;
		if (sets[l].miss && nreal)
			dat->missing_data = 1;
		free(sets[l].v);
	}
	free(sets);
	/* second sweep: codes in individual-major [I][L][P] order */
	for (int i = 0; i < dat->I; i++)
		for (int a = 0; a < P; a++) {
			const uint8_t *pr = prov + (size_t)(i * P + a) * L;
			uint8_t *out = dat->codes + (size_t)i * L * P + a;
			for (int l = 0; l < L; l++)
				out[(size_t)l * P] = rank[l][pr[l]];
		}
	free(prov);
	free(rank);
	*labels_out = labels;
	*nlab_out = nlab;
	return NO_ERROR;
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
	cursor cur, *fp = &cur, data_start;
	char *text, *name1, *name2, *word;
	size_t text_len = 0;
	int skip_line_two = 0, ncol, nhap, *raw = NULL, err = NO_ERROR;
	int32_t *labels = NULL;
	int64_t nlab = 0;

	if (flen > 4 && !strcmp(opt->filename + flen - 4, ".mcb"))
		return read_mcb_file(opt, dat);

	struct timespec tq0, tq1;
	clock_gettime(CLOCK_MONOTONIC, &tq0);
#define LAP(what) do { if (opt->timing) { clock_gettime(CLOCK_MONOTONIC, &tq1); \
	fprintf(stderr, "timing (s): parse: %-22s %.3f\n", what, (double)(tq1.tv_sec - tq0.tv_sec) \
		+ 1e-9 * (double)(tq1.tv_nsec - tq0.tv_nsec)); tq0 = tq1; } } while (0)
	if (!(text = slurp(opt->filename, &text_len)))
		return message(stderr, __FILE__, __func__, __LINE__, ERROR_MSG,
			FILE_OPEN_ERROR, opt->filename);
	cur.p = text;
	cur.end = text + text_len;
	LAP("file into memory");

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

	LAP("header, line count");
	/* raw alleles, haplotype-major like dat->IL */
	raw = malloc(sizeof *raw * (size_t)nhap * dat->L);
	dat->idv = calloc((size_t)dat->I, sizeof *dat->idv);
	if (!raw || !dat->idv) {
		free(raw);
		free(text);
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "genotype table\n");
	}
	data_start.p = text;
	data_start.end = text + text_len;
	*fp = data_start;
	skip_line(fp);
	if (skip_line_two)
		skip_line(fp);
	for (int h = 0, idv = 0; h < nhap; h += opt->interleaved ? dat->ploidy : 1) {
		const int rows = opt->interleaved ? dat->ploidy : 1;
		if (!(word = next_word(fp))) {
			/* fewer rows than the line count promised (a short last
			 * record): the recoder must not see uninitialised rows */
			free(raw);
			free(text);
			return mmessage(ERROR_MSG, END_OF_FILE, opt->filename);
		}
		if (opt->interleaved || !(h % dat->ploidy)) {
			dat->idv[idv].name = word;
			if (!(word = next_word(fp))) {
				free(raw);
				free(text);
				return mmessage(ERROR_MSG, END_OF_FILE, opt->filename);
			}
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
				if (!scan_int(fp, &v)) {
					free(raw);
					free(text);
					return mmessage(ERROR_MSG, FILE_FORMAT_ERROR,
						"failed to read locus %d of haplotype "
						"%d in file '%s'.  Check option -R.\n",
						l + 1, h + j + 1, opt->filename);
				}
				raw[(size_t)(h + j) * dat->L + l] = v;
			}
	}
	free(text);
	LAP("numbers");

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
	if (!dat->uniquealleles || !dat->nreal || !dat->allele_off || !dat->label_off
		|| !dat->codes) {
		free(raw);
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "recoded genotypes\n");
	}
	if ((err = recode_loci(dat, raw, nhap, &labels, &nlab))) {
		free(raw);
		return err;
	}
	LAP("labels and 8-bit codes");
	dat->label_off[dat->L] = nlab;
	dat->labels = labels;
	free(raw);
	finish_dims(dat);
	LAP("finish");
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
