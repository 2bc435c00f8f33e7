/*
 * write_results.c -- result files of a fit (reference write_file.c:203-732).
 *
 * File names and formats are the reference's.  The reference derives its
 * admixture partitions and Q tables from the 4-D posterior array diklm summed
 * over loci and alleles (write_file.c:359-381, 446-459, 525-543); those sums
 * D_ik are what the device keeps from the last E-step, so they are fetched
 * (mc_get_posterior) instead of the array that no longer exists.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "multiclust.h"

#define GPU(call) gpu_check(mod, (call), #call)

/* parameters of `slot` and posterior sums from all devices: p (and a pooled
 * eta) from device 0, the eta / posterior rows of each device's individuals */
int gather_state(options *opt, data *dat, model *mod, int slot, double *eta,
	double *p, double *post)
{
	const int sharded = opt->admixture && !opt->eta_constrained;

	(void)dat;
	for (int r = 0; r < mod->n_gpus; r++) {
		const size_t row = (size_t)mod->row_first[r] * mod->K;
		GPU(mc_get_params(mod->gpus[r], slot,
			sharded ? eta + row : (r == 0 ? eta : NULL), r == 0 ? p : NULL));
		GPU(mc_get_posterior(mod->gpus[r], post + row));
	}
	return NO_ERROR;
}

/* numerators of the popq tables (write_file.c:446-459, 658-666): the
 * posterior rows summed over the individuals of every locale, on the device
 * (mc_locale_sums); the shards of an individual-sharded fit are added in rank
 * order */
static int locale_sums(data *dat, model *mod)
{
	const size_t n = (size_t)dat->numpops * mod->K;
	int *loc = malloc(sizeof *loc * (size_t)dat->I);
	double *part = malloc(sizeof *part * (n ? n : 1));

	free(mod->popq_host);
	mod->popq_host = calloc(n ? n : 1, sizeof(double));
	if (!loc || !part || !mod->popq_host) {
		free(loc);
		free(part);
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "popq table\n");
	}
	for (int i = 0; i < dat->I; i++)
		loc[i] = dat->idv[i].locale;
	for (int r = 0; r < mod->n_gpus; r++) {
		GPU(mc_locale_sums(mod->gpus[r], loc + mod->row_first[r], dat->numpops, part));
		for (size_t x = 0; x < n; x++)
			mod->popq_host[x] += part[x];
	}
	free(loc);
	free(part);
	return NO_ERROR;
}

/* parameters of slot pindex and the posterior sums of the last E-step */
int fetch_results(options *opt, data *dat, model *mod)
{
	(void)opt;
	free(mod->eta_host);
	free(mod->p_host);
	free(mod->post_host);
	mod->eta_host = malloc(sizeof(double) * (size_t)mod->eta_len);
	mod->p_host = malloc(sizeof(double) * (size_t)(mod->T > 0 ? mod->K * mod->T : 1));
	mod->post_host = malloc(sizeof(double) * (size_t)dat->I * mod->K);
	if (!mod->eta_host || !mod->p_host || !mod->post_host)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "result buffers\n");
	{
		int err = gather_state(opt, dat, mod, mod->pindex, mod->eta_host, mod->p_host,
			mod->post_host);
		return err ? err : locale_sums(dat, mod);
	}
}

/* "<path>/<file>" or the -o prefix, followed by a formatted tail */
static FILE *open_result(options *opt, char *name, size_t len, const char *tail)
{
	FILE *fp;

	if (opt->outfile_name) {
		snprintf(name, len, "%s%s", opt->outfile_name, tail);
	} else {
		const size_t n = strlen(opt->path);
		const int sep = n && opt->path[n - 1] != '/' && opt->path[n - 1] != '\\';
		snprintf(name, len, "%s%s%s%s", opt->path, sep ? "/" : "",
			opt->filename_file, tail);
	}
	if (!(fp = fopen(name, "w")))
		message(stderr, __FILE__, __func__, __LINE__, ERROR_MSG,
			FILE_OPEN_ERROR, name);
	return fp;
}

/* argmax over k of the posterior, the first maximum wins
 * (write_file.c:369-375, 590-598); done on the device */
static void partition(data *dat, model *mod)
{
	int *cnt = calloc((size_t)mod->K, sizeof *cnt);

	for (int k = 0; k < mod->K; k++)
		mod->count_K[k] = 0;
	for (int r = 0; r < mod->n_gpus; r++) {
		GPU(mc_partition(mod->gpus[r], dat->I_K + mod->row_first[r], cnt));
		for (int k = 0; k < mod->K; k++)
			mod->count_K[k] += cnt[k];
	}
	free(cnt);
}

void partition_admixture(data *dat, model *mod)
{
	partition(dat, mod);
}

void partition_mixture(data *dat, model *mod)
{
	partition(dat, mod);
}

/* write_file.c:203-337 */
int write_file_detail(options *opt, data *dat, model *mod)
{
	const char *kind = opt->admixture ? "admix" : "mix";
	char name[4096], tail[128];
	FILE *fp;

	snprintf(tail, sizeof tail, ".%s.K=%d.out.txt", kind, mod->K);
	if (!(fp = open_result(opt, name, sizeof name, tail)))
		return FILE_OPEN_ERROR;
	fprintf(fp, "logL = %f (%s)\n", mod->logL,
		mod->converged ? "converged" : "not converged");
	fprintf(fp, "AIC = %f\n", aic(mod));
	fprintf(fp, "BIC = %f\n\n", bic(dat, mod));
	fprintf(fp, "count.K\n");
	for (int k = 0; k < mod->K; k++)
		fprintf(fp, "%d ", mod->count_K[k]);
	fprintf(fp, "\n\n");
	fclose(fp);

	if (!opt->admixture || opt->eta_constrained) {
		snprintf(tail, sizeof tail, ".%s.K=%d.etak.txt", kind, mod->K);
		if (!(fp = open_result(opt, name, sizeof name, tail)))
			return FILE_OPEN_ERROR;
		fprintf(fp, "i\tk\tetak\n");
		for (int k = 0; k < mod->K; k++)
			fprintf(fp, "%d\t%f\n", k, mod->eta_host[k]);
	} else {
		snprintf(tail, sizeof tail, ".%s.K=%d.etaik.txt", kind, mod->K);
		if (!(fp = open_result(opt, name, sizeof name, tail)))
			return FILE_OPEN_ERROR;
		fprintf(fp, "i\tk\tetaik\n");
		for (int i = 0; i < dat->I; i++)
			for (int k = 0; k < mod->K; k++)
				fprintf(fp, "%d\t%d\t%f\n", i, k,
					mod->eta_host[(size_t)i * mod->K + k]);
	}
	fprintf(fp, "\n");
	fclose(fp);

	snprintf(tail, sizeof tail, ".%s.K=%d.pklm.txt", kind, mod->K);
	if (!(fp = open_result(opt, name, sizeof name, tail)))
		return FILE_OPEN_ERROR;
	fprintf(fp, "k\tl\tm\tKLM\n");
	for (int k = 0; k < mod->K; k++)
		for (int l = 0; l < dat->L; l++)
			for (int m = 0; m < dat->uniquealleles[l]; m++)
				fprintf(fp, "%d\t%d\t%d\t%f\n", k, l, m,
					mod->p_host[(size_t)k * mod->T + dat->allele_off[l] + m]);
	fprintf(fp, "\n");
	fclose(fp);
	return NO_ERROR;
}

/* locale means of the posterior, scaled by `scale` */
static int write_popq(options *opt, data *dat, model *mod, const char *tail,
	double scale)
{
	char name[4096];
	double *q = calloc((size_t)dat->numpops * mod->K, sizeof *q);
	FILE *fp;

	if (!q)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "popq table\n");
	if (!(fp = open_result(opt, name, sizeof name, tail))) {
		free(q);
		return FILE_OPEN_ERROR;
	}
	for (int k = 0; k < mod->K; k++)
		for (int n = 0; n < dat->numpops; n++)
			q[(size_t)n * mod->K + k] = mod->popq_host[(size_t)n * mod->K + k]
				/ (scale * dat->i_p[n]);
	for (int n = 0; n < dat->numpops; n++) {
		fprintf(fp, "%s:\t", dat->pops[n]);
		for (int k = 0; k < mod->K; k++)
			fprintf(fp, "%lf\t", q[(size_t)n * mod->K + k]);
		fprintf(fp, "%d\n", dat->i_p[n]);
	}
	fclose(fp);
	free(q);
	return NO_ERROR;
}

static int write_indivq(options *opt, data *dat, model *mod, const char *tail,
	const double *q, double scale)
{
	char name[4096];
	FILE *fp;

	if (!(fp = open_result(opt, name, sizeof name, tail)))
		return FILE_OPEN_ERROR;
	for (int i = 0; i < dat->I; i++) {
		fprintf(fp, "%d\t%s\t(x)\t%s\t:", i, dat->idv[i].name,
			dat->pops[dat->idv[i].locale]);
		for (int k = 0; k < mod->K; k++)
			fprintf(fp, "\t%f", q[(size_t)i * mod->K + k] / scale);
		fprintf(fp, "\n");
	}
	fclose(fp);
	return NO_ERROR;
}

/* write_file.c:397-470: expected fraction of a locale's alleles from k */
int popq_admix(options *opt, data *dat, model *mod)
{
	char tail[64];

	snprintf(tail, sizeof tail, "_admix_popq_%d.popq", mod->K);
	return write_popq(opt, dat, mod, tail, (double)(dat->ploidy * dat->L));
}

/* write_file.c:485-565: posterior fractions when eta is pooled or data are
 * missing, else the estimated eta_ik */
int indivq_admix(options *opt, data *dat, model *mod)
{
	char tail[64];

	snprintf(tail, sizeof tail, "_admix_indivq_%d.indivq", mod->K);
	if (opt->eta_constrained || dat->missing_data)
		return write_indivq(opt, dat, mod, tail, mod->post_host,
			(double)(dat->ploidy * dat->L));
	return write_indivq(opt, dat, mod, tail, mod->eta_host, 1.0);
}

/* write_file.c:615-680 */
int popq_mix(options *opt, data *dat, model *mod)
{
	return write_popq(opt, dat, mod, "_mix_popq.popq", 1.0);
}

/* write_file.c:693-732 */
int indivq_mix(options *opt, data *dat, model *mod)
{
	char tail[64];

	snprintf(tail, sizeof tail, ".mix.K=%d.indivq", mod->K);
	return write_indivq(opt, dat, mod, tail, mod->post_host, 1.0);
}
