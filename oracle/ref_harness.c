/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference
 * objects (compiled in place from /root/reference by oracle/Makefile into
 * oracle/_ref/) and dumps what the parity tests need at full precision.
 *
 * Nothing here re-implements reference arithmetic: every number comes from
 * the reference's own read_file / initialize_model / em / em_step /
 * em_2_steps / accelerated_em_step / log_likelihood.  The harness only
 *   (a) replays main()'s setup sequence (multiclust.c:67-141),
 *   (b) optionally walks the em() driver loop (em_alg.c:44-90) one top-level
 *       step at a time so parameters can be dumped between steps,
 *   (c) can fill `data` from an MCB1 file instead of the O(n^2) text parser,
 *   (d) times em_step for the CPU baseline.
 *
 * usage: ref_harness [--dump PREFIX] [--steps] [--mcb FILE] [--time N]
 *                    [--parse-only] -- <multiclust command line>
 */
#include <limits.h>
#include <time.h>

#include "multiclust.h"	/* the reference header, via -I/root/reference */
#include "mc_format.h"

/* defined in the reference's multiclust.c but not declared in its header */
int make_options(options **opt);
int make_data(data **dat);
int make_model(model **mod);
int parse_options(options *opt, data *dat, int argc, const char **argv);
int allocate_model_for_k(options *opt, model *mod, data *dat);
int synchronize(options *opt, data *dat, model *mod);
void free_model_data(model *mod, options *opt);

/* ---- log likelihood capture (see ref_hooks.h) ---- */
static double *g_ll;
static size_t g_nll, g_cll;
static FILE *g_trace;
static int g_K, g_init;

void ref_hook_ll(double ll)
{
	if (g_nll == g_cll) {
		g_cll = g_cll ? 2 * g_cll : 1024;
		g_ll = realloc(g_ll, g_cll * sizeof *g_ll);
	}
	g_ll[g_nll++] = ll;
	if (g_trace) {
		fprintf(g_trace, "ll %d %d %zu %.17g\n", g_K, g_init, g_nll, ll);
		fflush(g_trace);	/* the reference may exit(0) at any step */
	}
}

static int total_slots(data *dat)
{
	int l, t = 0;
	for (l = 0; l < dat->L; l++)
		t += dat->uniquealleles[l];
	return t;
}

/* write parameters of slot `which` + posterior sums to PREFIX.<tag>.bin */
static void dump_state(const char *prefix, const char *tag, options *opt,
	data *dat, model *mod, int which)
{
	char name[4096];
	FILE *fp;
	int i, k, l, m, T = total_slots(dat);
	int per_indiv = opt->admixture && !opt->eta_constrained;
	int32_t hdr[6];
	double v;

	snprintf(name, sizeof name, "%s.%s.bin", prefix, tag);
	if (!(fp = fopen(name, "wb"))) {
		perror(name);
		exit(2);
	}
	hdr[0] = mod->K; hdr[1] = dat->I; hdr[2] = T;
	hdr[3] = per_indiv; hdr[4] = opt->admixture; hdr[5] = mod->n_iter;
	fwrite(hdr, 4, 6, fp);
	v = mod->logL;
	fwrite(&v, 8, 1, fp);
	if (per_indiv)
		for (i = 0; i < dat->I; i++)
			fwrite(mod->vetaik[which][i], 8, mod->K, fp);
	else
		fwrite(mod->vetak[which], 8, mod->K, fp);
	for (k = 0; k < mod->K; k++)
		for (l = 0; l < dat->L; l++)
			fwrite(mod->vpklm[which][k][l], 8, dat->uniquealleles[l], fp);
	/* what the writers consume: D_ik (write_file.c:359-381) or vik */
	for (i = 0; i < dat->I; i++)
		for (k = 0; k < mod->K; k++) {
			if (opt->admixture) {
				v = 0;
				for (l = 0; l < dat->L; l++)
					for (m = 0; m < dat->uniquealleles[l]; m++)
						v += mod->diklm[i][k][l][m];
			} else {
				v = mod->vik[i][k];
			}
			fwrite(&v, 8, 1, fp);
		}
	fclose(fp);
}

/* dump the reference's parse as an MCB1 file plus the ILM count table */
static void dump_parse(const char *prefix, data *dat)
{
	char name[4096];
	mcb_data d;
	FILE *fp;
	int i, l, a, m, nhap = dat->I * dat->ploidy;
	int64_t nlab = 0, pos;

	memset(&d, 0, sizeof d);
	d.I = dat->I; d.L = dat->L; d.P = dat->ploidy; d.npops = dat->numpops;
	d.J = malloc(4 * (size_t)d.L);
	d.nreal = malloc(4 * (size_t)d.L);
	d.locale = malloc(4 * (size_t)d.I);
	d.codes = malloc((size_t)d.I * d.L * d.P);
	for (l = 0; l < d.L; l++) {
		int miss = 0;
		for (i = 0; i < nhap; i++)
			if (dat->IL[i][l] == MISSING)
				miss = 1;
		d.J[l] = dat->uniquealleles[l];
		d.nreal[l] = d.J[l] ? d.J[l] - miss : 0;
		nlab += d.nreal[l];
	}
	d.labels = malloc(4 * (size_t)(nlab ? nlab : 1));
	pos = 0;
	for (l = 0; l < d.L; l++)
		for (m = 0; m < d.nreal[l]; m++)
			d.labels[pos++] = dat->L_alleles[l][m];
	for (i = 0; i < d.I; i++) {
		d.locale[i] = dat->idv[i].locale;
		for (l = 0; l < d.L; l++)
			for (a = 0; a < d.P; a++) {
				int al = dat->IL[i * d.P + a][l], code = MCB_MISSING;
				for (m = 0; m < d.nreal[l]; m++)
					if (dat->L_alleles[l][m] == al)
						code = m;
				d.codes[((size_t)i * d.L + l) * d.P + a] = (uint8_t)code;
			}
	}
	snprintf(name, sizeof name, "%s.parse.mcb", prefix);
	if (mcb_write(name, &d)) {
		perror(name);
		exit(2);
	}
	snprintf(name, sizeof name, "%s.parse.ilm", prefix);
	if (!(fp = fopen(name, "wb"))) {
		perror(name);
		exit(2);
	}
	for (i = 0; i < d.I; i++)
		for (l = 0; l < d.L; l++)
			fwrite(dat->ILM[i][l], 4, dat->uniquealleles[l], fp);
	fclose(fp);
	/* names and locale strings, one individual per line */
	snprintf(name, sizeof name, "%s.parse.names", prefix);
	if ((fp = fopen(name, "w"))) {
		for (i = 0; i < d.I; i++)
			fprintf(fp, "%s\t%s\n", dat->idv[i].name,
				dat->pops[dat->idv[i].locale]);
		fclose(fp);
	}
	free(d.J); free(d.nreal); free(d.locale); free(d.codes); free(d.labels);
}

/* build the reference's `data` from an MCB1 file (replaces read_file only) */
static int load_mcb(const char *path, options *opt, data *dat)
{
	mcb_data d;
	int i, l, a, m, nhap;
	char buf[64];

	if (mcb_read(path, &d)) {
		fprintf(stderr, "cannot read MCB file %s\n", path);
		return 1;
	}
	dat->I = d.I; dat->L = d.L; dat->ploidy = d.P;
	nhap = d.I * d.P;
	MAKE_2ARRAY(dat->IL, nhap, dat->L);
	MAKE_1ARRAY(dat->idv, dat->I);
	MAKE_1ARRAY(dat->I_K, dat->I);
	if (opt->admixture)
		MAKE_2ARRAY(dat->IL_K, nhap, dat->L);
	CMAKE_1ARRAY(dat->uniquealleles, dat->L);
	dat->M = 0;
	dat->missing_data = 0;
	for (l = 0; l < d.L; l++) {
		dat->uniquealleles[l] = d.J[l];
		if (d.J[l] > dat->M)
			dat->M = d.J[l];
		if (d.J[l] > d.nreal[l])
			dat->missing_data = 1;
	}
	MAKE_2JAGGED_ARRAY(dat->L_alleles, dat->L, dat->uniquealleles);
	for (l = 0; l < d.L; l++) {
		for (m = 0; m < d.nreal[l]; m++)
			dat->L_alleles[l][m] = d.labels[d.lab_off[l] + m];
		/* phantom slot: the reference leaves this label unset
		 * (read_file.c:580-585); use a value no allele can equal */
		for (; m < d.J[l]; m++)
			dat->L_alleles[l][m] = INT_MIN + 7;
	}
	for (i = 0; i < d.I; i++)
		for (l = 0; l < d.L; l++)
			for (a = 0; a < d.P; a++) {
				int c = d.codes[((size_t)i * d.L + l) * d.P + a];
				dat->IL[i * d.P + a][l] = c == MCB_MISSING
					? MISSING : d.labels[d.lab_off[l] + c];
			}
	/* counts exactly as sufficient_statistics() does for labelled
	 * alleles (read_file.c:651-657): missing copies are counted nowhere */
	CMAKE_3JAGGED_ARRAY(dat->ILM, dat->I, dat->L, dat->uniquealleles);
	for (i = 0; i < d.I; i++)
		for (l = 0; l < d.L; l++)
			for (a = 0; a < d.P; a++) {
				int c = d.codes[((size_t)i * d.L + l) * d.P + a];
				if (c != MCB_MISSING)
					dat->ILM[i][l][c]++;
			}
	dat->numpops = d.npops;
	dat->pops = malloc(sizeof *dat->pops * (size_t)d.npops);
	CMAKE_1ARRAY(dat->i_p, d.npops);
	for (i = 0; i < d.npops; i++) {
		snprintf(buf, sizeof buf, "pop%d", i);
		dat->pops[i] = strdup(buf);
	}
	for (i = 0; i < d.I; i++) {
		snprintf(buf, sizeof buf, "ind%d", i);
		dat->idv[i].name = strdup(buf);
		dat->idv[i].locale = d.locale[i];
		dat->i_p[d.locale[i]]++;
	}
	mcb_free(&d);
	return 0;
}

/* --bootstrap-sample seed (needs -b n on the multiclust command line so that the
 * reference allocates mle_*): the fit's final parameters become the H0 estimates
 * exactly as maximize_likelihood() copies them (multiclust.c:562-581), the
 * generator is re-seeded so that the draws are a known stream, the reference's
 * own parametric_bootstrap() (bootstrap.c:31-52) makes one sample, and the
 * sample's allele counts are written as int32 [I][sum of uniquealleles] */
static void dump_bootstrap_sample(const char *prefix, int seed, options *opt,
	data *dat, model *mod)
{
	char name[4096];
	FILE *fp;
	int i, l, m, k;

	if (!opt->n_bootstrap || !mod->mle_pKLM) {
		fprintf(stderr, "ref_harness: --bootstrap-sample needs -b n\n");
		exit(2);
	}
	for (k = 0; k < mod->K; k++)
		for (l = 0; l < dat->L; l++)
			for (m = 0; m < dat->uniquealleles[l]; m++)
				mod->mle_pKLM[k][l][m] = mod->vpklm[mod->pindex][k][l][m];
	if (!opt->admixture || opt->eta_constrained)
		for (k = 0; k < mod->K; k++)
			mod->mle_etak[k] = mod->vetak[mod->pindex][k];
	else
		for (i = 0; i < dat->I; i++)
			for (k = 0; k < mod->K; k++)
				mod->mle_etaik[i][k] = mod->vetaik[mod->pindex][i][k];
	srand((unsigned)seed);
	if (parametric_bootstrap(opt, dat, mod))
		exit(3);
	snprintf(name, sizeof name, "%s.bootstrap.bin", prefix);
	if (!(fp = fopen(name, "wb")))
		exit(4);
	for (i = 0; i < dat->I; i++)
		for (l = 0; l < dat->L; l++)
			for (m = 0; m < dat->uniquealleles[l]; m++) {
				const int32_t c = dat->ILM[i][l][m];
				fwrite(&c, sizeof c, 1, fp);
			}
	fclose(fp);
	cleanup_parametric_bootstrap(dat);
}

static double now_sec(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* em() walked one top-level step at a time (same calls, same order as
 * em_alg.c:44-90) so that the state can be dumped between steps */
static void em_stepwise(const char *prefix, options *opt, data *dat, model *mod)
{
	char tag[128];
	int i, stop = 0, nstep = 0;

	if (mod->K == 1) {
		em_step(opt, dat, mod);
		mod->logL = log_likelihood(opt, dat, mod, mod->tindex);
		return;
	}
	while (mod->n_iter < opt->n_init_iter && !stop)
		stop = em_step(opt, dat, mod);
	for (i = 1; i < opt->q; i++) {
		em_2_steps(mod, dat, opt);
		mod->pindex = mod->findex;
	}
	if (mod->converged)
		return;
	do {
		stop = opt->accel_scheme ? accelerated_em_step(opt, dat, mod)
					 : em_step(opt, dat, mod);
		nstep++;
		snprintf(tag, sizeof tag, "K%d.init%d.step%d", g_K, g_init, nstep);
		dump_state(prefix, tag, opt, dat, mod, mod->pindex);
		if (g_trace)
			fprintf(g_trace, "step %d %d %d n_iter=%d pindex=%d "
				"logL=%.17g\n", g_K, g_init, nstep, mod->n_iter,
				mod->pindex, mod->logL);
	} while (!stop);
}

int main(int argc, const char **argv)
{
	const char *prefix = NULL, *mcb = NULL;
	int steps = 0, ntime = 0, parse_only = 0, a = 1, err, i;
	int boot_seed = -1;
	options *opt = NULL;
	data *dat = NULL;
	model *mod = NULL;
	char name[4096], tag[128];

	for (; a < argc && strcmp(argv[a], "--"); a++) {
		if (!strcmp(argv[a], "--dump") && a + 1 < argc)
			prefix = argv[++a];
		else if (!strcmp(argv[a], "--mcb") && a + 1 < argc)
			mcb = argv[++a];
		else if (!strcmp(argv[a], "--time") && a + 1 < argc)
			ntime = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--steps"))
			steps = 1;
		else if (!strcmp(argv[a], "--parse-only"))
			parse_only = 1;
		else if (!strcmp(argv[a], "--bootstrap-sample") && a + 1 < argc)
			boot_seed = atoi(argv[++a]);
		else {
			fprintf(stderr, "ref_harness: bad option %s\n", argv[a]);
			return 2;
		}
	}
	if (a == argc) {
		fprintf(stderr, "usage: ref_harness [opts] -- <multiclust args>\n");
		return 2;
	}
	/* argv[a] == "--" plays the role of argv[0] for parse_options */
	if ((err = make_options(&opt)) || (err = make_data(&dat))
		|| (err = make_model(&mod)))
		return err;
	if ((err = parse_options(opt, dat, argc - a, argv + a)))
		return err;
	if (mcb)
		err = load_mcb(mcb, opt, dat);
	else
		err = read_file(opt, dat);
	if (err)
		return err;
	if (prefix) {
		dump_parse(prefix, dat);
		snprintf(name, sizeof name, "%s.trace.txt", prefix);
		g_trace = fopen(name, "w");
	}
	if (parse_only)
		return 0;
	if ((err = synchronize(opt, dat, mod)))
		return err;
	if (g_trace)
		fprintf(g_trace, "bounds %.17g %.17g q=%d n_init=%d\n",
			opt->eta_lower_bound, opt->p_lower_bound, opt->q,
			opt->n_init);

	dat->max_M = dat->M;
	mod->max_logL = -INFINITY;
	for (mod->K = opt->min_K; mod->K <= opt->max_K; mod->K++) {
		g_K = mod->K;
		if (dat->max_M < mod->K)
			dat->max_M = mod->K;
		if ((err = allocate_model_for_k(opt, mod, dat)))
			return err;
		mod->start = clock();
		for (i = 0; i < opt->n_init; i++) {
			g_init = i;
			/* per-initialisation resets of maximize_likelihood()
			 * (multiclust.c:518-524) */
			mod->current_i = mod->current_l = mod->current_k = 0;
			mod->logL = 0.0;
			mod->converged = mod->stopped = mod->iter_stop = 0;
			if ((err = initialize_model(opt, dat, mod)))
				return err;
			if (prefix) {
				snprintf(tag, sizeof tag, "K%d.init%d.start", g_K, i);
				dump_state(prefix, tag, opt, dat, mod, 0);
			}
			if (ntime > 0) {
				double t0, t1;
				int s;
				em_step(opt, dat, mod);	/* warm-up */
				t0 = now_sec();
				for (s = 0; s < ntime; s++)
					em_step(opt, dat, mod);
				t1 = now_sec();
				printf("{\"kind\": \"reference\", \"K\": %d, "
					"\"I\": %d, \"L\": %d, \"P\": %d, "
					"\"T\": %d, \"admixture\": %d, "
					"\"steps\": %d, \"sec_per_step\": %.9g, "
					"\"logL\": %.17g}\n", mod->K, dat->I,
					dat->L, dat->ploidy, total_slots(dat),
					opt->admixture, ntime,
					(t1 - t0) / ntime, mod->logL);
				fflush(stdout);
				break;
			}
			if (steps && prefix)
				em_stepwise(prefix, opt, dat, mod);
			else
				em(opt, dat, mod);
			if (g_trace)
				fprintf(g_trace, "fit %d %d logL=%.17g converged=%d "
					"stopped=%d iter_stop=%d n_iter=%d "
					"pindex=%d\n", g_K, i, mod->logL,
					mod->converged, mod->stopped,
					mod->iter_stop, mod->n_iter, mod->pindex);
			if (prefix) {
				snprintf(tag, sizeof tag, "K%d.init%d.final", g_K, i);
				dump_state(prefix, tag, opt, dat, mod, mod->pindex);
			}
			if (boot_seed >= 0 && prefix)
				dump_bootstrap_sample(prefix, boot_seed, opt, dat, mod);
			if (mod->K == 1)
				break;
		}
		free_model_data(mod, opt);
	}
	if (g_trace)
		fclose(g_trace);
	return 0;
}
