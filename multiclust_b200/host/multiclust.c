/*
 * multiclust.c -- command line and orchestration of the B200 drop-in
 * (reference multiclust.c:67-160, 365-660, 715-978, 1181-1279, 1396-1735).
 *
 * Same command line as the reference for the EM path: -a -c -k -1 -2 -m -n -s
 * -p --missing -f -d -o -e -E -g -i -r -t -T -v -w -M -R --projection --bound
 * -b (parametric bootstrap), plus -C (the iteration cap the reference documents
 * but only implements as -T, README.md:42 vs multiclust.c:1634) and three
 * additions: --device, --gpus and --trace.  Options of the reference that
 * belong to parts outside the EM path (-x block relaxation, --simulate, -I
 * index input, --impute, -P/-Q warm start, -A, -u) are recognised and refused
 * with a clear message instead of being half-implemented.
 *
 * The K loop, the initialisation loop with its best-so-far bookkeeping and the
 * per-initialisation / summary output lines keep the reference's order and
 * text; one shared rand() stream runs across all initialisations and all K.
 */
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "mc_format.h"
#include "multiclust.h"

/* --timing: wall-clock seconds of every phase on stderr */
static double wall_now(void)
{
	struct timespec ts;

	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static int g_timing;
static double g_t_init, g_t_em, g_t_write, g_t_plan, g_t_record;

/* the initialisers' random stream: glibc's rand() state after no srand()
 * call equals srand(1) */
static mcr_state g_rng;

static const char *accel_abbrev[NUM_ACCELERATION_METHODS] = { "EM", "S1", "S2", "S3", "Q" };
static const char *accel_names[NUM_ACCELERATION_METHODS] = {
	"No acceleration", "SQUAREM version 1", "SQUAREM version 2",
	"SQUAREM version 3", "Quasi Newton",
};

/* formatted message with the reference's prefixes (message.c:26-126) */
int message(FILE *fp, const char *file, const char *fxn, int line, int type,
	int id, const char *fmt, ...)
{
	static const char *kind[] = { "", "INFO", "DEBUG", "WARNING", "ERROR" };
	va_list ap;

	fprintf(fp, "%s [%s::%s(%d)]: ", kind[type < 1 || type > 4 ? 4 : type],
		file, fxn, line);
	switch (id) {
	case MEMORY_ALLOCATION:
		fprintf(fp, "could not allocate ");
		break;
	case INVALID_CMD_OPTION:
		fprintf(fp, "unrecognized command option: ");
		break;
	case INVALID_CMD_ARGUMENT:
		fprintf(fp, "invalid argument to command option: ");
		break;
	case INVALID_CMDLINE:
		fprintf(fp, "[invalid command line] ");
		break;
	case INVALID_USER_SETUP:
		fprintf(fp, "[invalid user choice] ");
		break;
	case FILE_OPEN_ERROR:
		fprintf(fp, "could not open file \"%s\"\n", fmt);
		return id;
	case END_OF_FILE:
		fprintf(fp, "unexpected end of file in file \"%s\"\n", fmt);
		return id;
	case FILE_FORMAT_ERROR:
		fprintf(fp, "invalid file format: ");
		break;
	case GPU_ERROR:
		fprintf(fp, "[device] ");
		break;
	default:
		break;
	}
	if (fmt) {
		va_start(ap, fmt);
		vfprintf(fp, fmt, ap);
		va_end(ap);
	}
	return id;
}

/* ------------------------------------------------------------- objects */

int make_options(options **out)
{
	options *opt = calloc(1, sizeof *opt);

	if (!opt)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "options object\n");
	/* defaults of reference multiclust.c:912-975 */
	opt->path = "./";
	opt->missing_value = MISSING;
	opt->seed = 1234567;
	opt->n_init = 50;
	opt->abs_error = 1e-4;
	opt->min_K = opt->max_K = 6;
	opt->n_rand_em_init = 50;
	opt->lower_bound = opt->eta_lower_bound = opt->p_lower_bound = 1e-8;
	opt->do_projection = 1;
	opt->q = 1;
	opt->verbosity = MINIMAL;
	opt->compact = 1;
	opt->n_repeat = 1;
	opt->write_files = 1;
	opt->n_gpus = 1;
	opt->fits_per_gpu = 1;
	*out = opt;
	return NO_ERROR;
}

int make_data(data **out)
{
	data *dat = calloc(1, sizeof *dat);

	if (!dat)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "data object\n");
	dat->ploidy = 2;
	*out = dat;
	return NO_ERROR;
}

int make_model(model **out)
{
	model *mod = calloc(1, sizeof *mod);

	if (!mod)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "model object\n");
	mod->K = 1;
	mod->max_logL = -INFINITY;
	mcr_seed(&g_rng, 1);
	mod->rng = &g_rng;
	*out = mod;
	return NO_ERROR;
}

void free_options(options *opt)
{
	free(opt);
}

void free_data(data *dat)
{
	if (!dat)
		return;
	for (int i = 0; dat->idv && i < dat->I; i++)
		free(dat->idv[i].name);
	for (int n = 0; n < dat->numpops; n++)
		free(dat->pops[n]);
	free(dat->pops); free(dat->idv); free(dat->i_p); free(dat->I_K);
	free(dat->uniquealleles); free(dat->nreal); free(dat->allele_off);
	free(dat->labels); free(dat->label_off); free(dat->codes); free(dat->codes_orig);
	free(dat);
}

void free_model_data(model *mod, options *opt)
{
	(void)opt;
	free(mod->count_K);
	free(mod->eta_host); free(mod->p_host); free(mod->post_host); free(mod->popq_host);
	if (mod->comm)
		mc_comm_destroy(mod->comm);
	mod->comm = NULL;
	mod->count_K = NULL;
	mod->eta_host = mod->p_host = mod->post_host = mod->popq_host = NULL;
}

void free_model(model *mod, options *opt)
{
	if (!mod)
		return;
	free_model_data(mod, opt);
	free(mod->A); free(mod->Ainv); free(mod->cutu);
	for (int r = 0; r < mod->n_gpus; r++)
		if (mod->gpus[r])
			mc_destroy(mod->gpus[r]);
	free(mod->gpus);
	free(mod->row_first);
	if (mod->trace)
		fclose(mod->trace);
	free(mod);
}

/* ------------------------------------------------------- command line */

void fprint_usage(FILE *fp, const char *cmd)
{
	fprintf(fp,
"usage: %s -f <structure file> [options]\n\n"
"model\n"
"  -a            admixture model (default: mixture model)\n"
"  -c            one pooled mixing-proportion vector (with -a)\n"
"  -k <K>        fit K subpopulations; or -1 <Kmin> -2 <Kmax> for a sweep\n"
"  -p <ploidy>   allele copies per locus (default 2)\n"
"  --missing <v> allele value that marks missing data (default -9)\n"
"  -R            first line carries two extra column names\n"
"fitting\n"
"  -n <n>        random initialisations per K (default 50)\n"
"  -m <n>        Rand-EM initialisations (accepted; inert as in the reference)\n"
"  -s <0..6>     0 EM, 1-3 SQUAREM, 4-6 quasi-Newton with q = s-3 secants\n"
"  -C <n>, -T <n>  stop after more than n iterations\n"
"  -e <x> / -E <x>  relative / absolute log likelihood convergence error\n"
"  -i <n>        plain EM iterations before accelerating\n"
"  -g <n>        step-size halvings tried when an accelerated step fails\n"
"  -t <min>      time limit; -r <seed>  seed rand(); --bound <x>  parameter floor\n"
"  -b <n>        parametric bootstrap: n samples under H0: K-1, tested against Ha: K\n"
"  --projection  do not project onto the simplex\n"
"output\n"
"  -d <dir>      directory of the result files; -o <prefix> their name prefix\n"
"  -v <level>    verbosity (4: one line per iteration); -w n <r> [t <min>] [m <min>]  timed repeats\n"
"  -M            print only the maximum log likelihood\n"
"device\n"
"  --device <d>  first CUDA device ordinal (default 0)\n"
"  --gpus <n>    shard the individuals of one fit over n devices (NCCL exchange)\n"
"  --shard-fits  with --gpus: deal whole fits (K, initialisation) to the devices\n"
"  --fits-per-gpu <m>  with --shard-fits: m fits in flight on every device\n"
"  --trace <f>   write every log likelihood at full precision to <f>\n"
"  --timing      wall-clock seconds of every phase on stderr\n"
"  --dump <pre>  binary parameters before / after every fit to <pre>.K*.init*.bin\n"
"  --parse-only <f>  read and recode the data, write it as MCB1 to <f>, stop\n", cmd);
}

static void usage_error(const char **argv, int i)
{
	mmessage(ERROR_MSG, INVALID_CMD_ARGUMENT, "'%s'\n", argv[i]);
	fprintf(stderr, "Try '%s -h' for the list of options.\n", argv[0]);
}

static int unsupported(const char *arg, const char *what)
{
	return mmessage(ERROR_MSG, INVALID_USER_SETUP, "option '%s' (%s) belongs to a "
		"part of MULTICLUST outside the EM path this build accelerates\n",
		arg, what);
}

static int read_int_arg(int argc, const char **argv, int i, long lo, int *out)
{
	char *end;
	long v;

	if (i >= argc)
		return 1;
	errno = 0;
	v = strtol(argv[i], &end, 0);
	if (errno || end == argv[i] || *end || v < lo || v > 2147483647L)
		return 1;
	*out = (int)v;
	return 0;
}

static int read_double_arg(int argc, const char **argv, int i, double *out)
{
	char *end;

	if (i >= argc)
		return 1;
	errno = 0;
	*out = strtod(argv[i], &end);
	return errno || end == argv[i] || *end || *out < 0;
}

/* keyed on the first letter after the dashes, with the reference's two-letter
 * disambiguations (multiclust.c:1402-1718) */
int parse_options(options *opt, data *dat, int argc, const char **argv)
{
	int i, tmp;

	for (i = 1; i < argc; i++) {
		const char *name;
		size_t j = 1;

		if (strlen(argv[i]) < 2) {
			usage_error(argv, i);
			return INVALID_CMD_ARGUMENT;
		}
		while (argv[i][j] == '-' && argv[i][j + 1])
			j++;
		name = argv[i] + j;
		switch (name[0]) {
		case 'a':
			opt->admixture = 1;
			break;
		case 'A':
			return unsupported(argv[i], "partition file for the adjusted Rand index");
		case 'b':
			if (!strncmp(name, "bou", 3)) {
				if (read_double_arg(argc, argv, ++i, &opt->lower_bound))
					goto bad_arg;
				break;
			}
			/* -b n: parametric bootstrap (reference multiclust.c:1427-1434) */
			if (read_int_arg(argc, argv, ++i, 0, &opt->n_bootstrap))
				goto bad_arg;
			break;
		case 'c':
			opt->eta_constrained = 1;
			break;
		case 'C':	/* README.md:42; the reference binary only knows -T */
		case 'T':
			if (read_int_arg(argc, argv, ++i, 0, &opt->max_iter))
				goto bad_arg;
			break;
		case 'd':
			if (!strncmp(name, "dev", 3)) {
				if (read_int_arg(argc, argv, ++i, 0, &opt->device))
					goto bad_arg;
				break;
			}
			if (!strncmp(name, "du", 2)) {
				if (++i >= argc)
					goto bad_arg;
				opt->dump_prefix = argv[i];
				break;
			}
			if (++i >= argc)
				goto bad_arg;
			opt->path = argv[i];
			break;
		case 'e':
			if (read_double_arg(argc, argv, ++i, &opt->rel_error))
				goto bad_arg;
			break;
		case 'E':
			if (read_double_arg(argc, argv, ++i, &opt->abs_error))
				goto bad_arg;
			break;
		case 'f':
			if (!strncmp(name, "fi", 2)) {	/* --fits-per-gpu */
				if (read_int_arg(argc, argv, ++i, 1, &opt->fits_per_gpu))
					goto bad_arg;
				break;
			}
			if (!strncmp(name, "fo", 2))
				return unsupported(argv[i], "output format of imputed data");
			if (++i >= argc)
				goto bad_arg;
			opt->filename = argv[i];
			opt->filename_file = strrchr(argv[i], '/');
			opt->filename_file = opt->filename_file ? opt->filename_file + 1 : argv[i];
			break;
		case 'g':
			if (!strncmp(name, "gpu", 3)) {
				if (read_int_arg(argc, argv, ++i, 1, &opt->n_gpus))
					goto bad_arg;
				break;
			}
			if (read_int_arg(argc, argv, ++i, 0, &opt->adjust_step))
				goto bad_arg;
			break;
		case 'h':
			fprint_usage(stdout, argv[0]);
			return CUSTOM_ERROR;
		case 'i':
			if (!strncmp(name, "im", 2))
				return unsupported(argv[i], "imputation of missing data");
			if (read_int_arg(argc, argv, ++i, 0, &opt->n_init_iter))
				goto bad_arg;
			break;
		case 'I':
			return unsupported(argv[i], "alleles given as indices");
		case '1':
			if (read_int_arg(argc, argv, ++i, 1, &opt->min_K))
				goto bad_arg;
			break;
		case '2':
			if (read_int_arg(argc, argv, ++i, 1, &opt->max_K))
				goto bad_arg;
			break;
		case 'k':
			if (read_int_arg(argc, argv, ++i, 1, &opt->max_K))
				goto bad_arg;
			opt->min_K = opt->max_K;
			break;
		case 'm':
			if (!strncmp(name, "mi", 2)) {
				if (read_int_arg(argc, argv, ++i, -2147483647L, &opt->missing_value))
					goto bad_arg;
			} else if (read_int_arg(argc, argv, ++i, 0, &opt->n_rand_em_init)) {
				goto bad_arg;
			}
			break;
		case 'M':
			opt->parallel = 1;
			opt->n_repeat = 1;
			opt->verbosity = SILENT;
			break;
		case 'n':
			if (read_int_arg(argc, argv, ++i, -2147483647L, &opt->n_init))
				goto bad_arg;
			if (opt->n_init == 0)
				opt->n_repeat = 0;
			break;
		case 'o':
			if (++i >= argc)
				goto bad_arg;
			opt->outfile_name = argv[i];
			break;
		case 'p':
			if (!strncmp(name, "pa", 2)) {	/* --parse-only <out.mcb> */
				if (++i >= argc)
					goto bad_arg;
				opt->parse_only = argv[i];
			} else if (!strncmp(name, "pr", 2)) {
				opt->do_projection = 0;
			} else if (!strncmp(name, "pl", 2)) {
				return unsupported(argv[i], "rewriting the data file");
			} else if (read_int_arg(argc, argv, ++i, 1, &dat->ploidy)) {
				goto bad_arg;
			}
			break;
		case 'P':
		case 'Q':
			return unsupported(argv[i], "warm start from P/Q files");
		case 'R':
			opt->R_format = 1;
			break;
		case 'r':
			if (read_int_arg(argc, argv, ++i, 0, &tmp))
				goto bad_arg;
			opt->seed = (unsigned int)tmp;
			mcr_seed(&g_rng, opt->seed);	/* srand(seed), multiclust.c:1595 */
			break;
		case 's':
			if (!strncmp(name, "sh", 2)) {	/* --shard-fits */
				opt->shard_fits = 1;
				break;
			}
			if (!strncmp(name, "si", 2))
				return unsupported(argv[i], "data simulation");
			if (read_int_arg(argc, argv, ++i, 0, &opt->accel_scheme))
				goto bad_arg;
			break;
		case 't':
			if (!strncmp(name, "tr", 2)) {
				if (++i >= argc)
					goto bad_arg;
				opt->trace_file = argv[i];
				break;
			}
			if (!strncmp(name, "ti", 2)) {	/* --timing */
				opt->timing = 1;
				break;
			}
			if (read_int_arg(argc, argv, ++i, 0, &tmp))
				goto bad_arg;
			opt->n_seconds = 60u * (unsigned int)tmp;
			break;
		case 'u':
			return unsupported(argv[i], "target log likelihood search");
		case 'v':
			if (i + 1 == argc || read_int_arg(argc, argv, i + 1, 0, &opt->verbosity))
				opt->verbosity = VERBOSE;
			else
				i++;
			break;
		case 'w':
			while (++i < argc && argv[i][0] != '-') {
				if (argv[i][0] == 'n') {
					if (read_int_arg(argc, argv, ++i, 1, &opt->n_repeat))
						goto bad_arg;
				} else if (argv[i][0] == 't' || argv[i][0] == 'm') {
					if (read_int_arg(argc, argv, ++i, 0, &tmp))
						goto bad_arg;
					if (argv[i - 1][0] == 't')
						opt->repeat_seconds = 60 * tmp;
					else
						opt->max_repeat_seconds = 60 * tmp;
				} else {	/* reference multiclust.c:1706-1708 */
					usage_error(argv, i);
					return INVALID_CMD_OPTION;
				}
			}
			i--;
			opt->write_files = 0;
			break;
		case 'x':
			return unsupported(argv[i], "block relaxation");
		default:
			mmessage(ERROR_MSG, INVALID_CMD_OPTION, "'%s'\n", argv[i]);
			return INVALID_CMD_OPTION;
		}
	}
	if (!opt->filename)
		return message(stderr, __FILE__, __func__, __LINE__, ERROR_MSG,
			INVALID_CMDLINE, "You must specify the data file with command "
			"line option '-f'.  Try '-h' for help.\n");
	return NO_ERROR;
bad_arg:
	usage_error(argv, i < argc ? i : argc - 1);
	return INVALID_CMD_ARGUMENT;
}

/* reference multiclust.c:807-893 */
int synchronize(options *opt, data *dat, model *mod)
{
	const double floor_ = 1.0 / dat->I / dat->ploidy - 0.5 / dat->I / dat->ploidy;

	if (floor_ < opt->lower_bound)
		opt->lower_bound = floor_;
	opt->eta_lower_bound = opt->p_lower_bound = opt->lower_bound;

	if (opt->accel_scheme >= QN) {
		opt->adjust_step = 0;
		opt->q = opt->accel_scheme - SQS3;
		if (opt->q > 3)
			return mmessage(ERROR_MSG, INVALID_USER_SETUP, "Cannot use "
				"acceleration methods greater than 6 (QN3) without "
				"linking to lapack.\n");
		snprintf(opt->accel_abbreviation, sizeof opt->accel_abbreviation,
			"Q%d", opt->q);
		snprintf(opt->accel_name, sizeof opt->accel_name, "%s (q=%d)",
			accel_names[QN], opt->q);
		mod->A = calloc((size_t)opt->q * opt->q, sizeof(double));
		mod->Ainv = calloc((size_t)opt->q * opt->q, sizeof(double));
		mod->cutu = calloc((size_t)opt->q, sizeof(double));
	} else {
		snprintf(opt->accel_abbreviation, sizeof opt->accel_abbreviation,
			"%s", accel_abbrev[opt->accel_scheme]);
		snprintf(opt->accel_name, sizeof opt->accel_name, "%s",
			accel_names[opt->accel_scheme]);
	}
	if (dat->I < opt->max_K)
		return mmessage(ERROR_MSG, INVALID_USER_SETUP, "Maximum number of "
			"clusters (%d) (set with command-line argument -k) cannot exceed "
			"the number of individuals (%d)\n", opt->max_K, dat->I);
	/* reference multiclust.c:869-877 */
	if (opt->n_bootstrap && opt->max_K <= 1)
		return mmessage(ERROR_MSG, INVALID_USER_SETUP, "When bootstrapping, maximum "
			"K (%d) (set with command-line argument -k) must exceed 1.", opt->max_K);
	if (opt->n_bootstrap) {
		mod->null_K = opt->max_K - 1;
		mod->alt_K = opt->max_K;
		if (opt->shard_fits || opt->fits_per_gpu > 1)
			return mmessage(ERROR_MSG, INVALID_USER_SETUP, "The parametric "
				"bootstrap (-b) fits one model at a time: drop --shard-fits / "
				"--fits-per-gpu (--gpus N shards the individuals).\n");
		if (opt->n_repeat != 1)
			return mmessage(ERROR_MSG, INVALID_USER_SETUP, "The parametric "
				"bootstrap (-b) cannot be timed (-w).\n");
	}
	if (!opt->n_seconds && !opt->n_init)
		opt->n_init = 1;
	if (opt->min_K > opt->max_K)
		return mmessage(ERROR_MSG, INVALID_USER_SETUP, "Minimum K (%d) must not "
			"exceed maximum K (%d).", opt->min_K, opt->max_K);
	return NO_ERROR;
}

/* reference multiclust.c:1181-1279: device buffers for K clusters */
int allocate_model_for_k(options *opt, model *mod, data *dat)
{
	int rc;

	mod->count_K = calloc((size_t)mod->K, sizeof *mod->count_K);
	if (!mod->count_K)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "cluster sizes\n");
	for (int r = 0; r < mod->n_gpus; r++) {
		rc = mc_alloc_model(mod->gpus[r], mod->K, opt->admixture,
			opt->eta_constrained, opt->accel_scheme ? opt->q : 0,
			opt->eta_lower_bound, opt->p_lower_bound, opt->do_projection);
		if (rc)
			return mmessage(ERROR_MSG, GPU_ERROR, "%s\n",
				mc_last_error(mod->gpus[r]));
	}
	if (mod->n_gpus > 1 && mc_comm_create(&mod->comm, mod->gpus, mod->n_gpus))
		return mmessage(ERROR_MSG, GPU_ERROR, "%s\n", mc_comm_last_error(NULL));
	mod->eta_len = (opt->admixture && !opt->eta_constrained)
		? (int64_t)dat->I * mod->K : mod->K;
	/* parameter count: phantom slots included (multiclust.c:1268-1276) */
	mod->no_parameters = (!opt->admixture || opt->eta_constrained)
		? mod->K - 1 : dat->I * (mod->K - 1);
	for (int l = 0; l < dat->L; l++)
		mod->no_parameters += (dat->uniquealleles[l] - 1) * mod->K;
	return NO_ERROR;
}

/* --dump: parameters of `slot` and the posterior sums at full precision, in
 * the layout oracle/ref_harness.c uses for the reference (tests compare them) */
static int dump_state(options *opt, data *dat, model *mod, int init, const char *tag,
	int slot)
{
	char name[4096];
	int32_t hdr[6];
	size_t np = (size_t)mod->K * mod->T;
	double *eta = malloc(sizeof(double) * (size_t)mod->eta_len);
	double *p = malloc(sizeof(double) * (np ? np : 1));
	double *post = malloc(sizeof(double) * (size_t)dat->I * mod->K);
	FILE *fp;

	snprintf(name, sizeof name, "%s.K%d.init%d.%s.bin", opt->dump_prefix, mod->K,
		init, tag);
	if (!eta || !p || !post || !(fp = fopen(name, "wb")))
		return message(stderr, __FILE__, __func__, __LINE__, ERROR_MSG,
			FILE_OPEN_ERROR, name);
	gather_state(opt, dat, mod, slot, eta, p, post);
	hdr[0] = mod->K; hdr[1] = dat->I; hdr[2] = (int32_t)mod->T;
	hdr[3] = opt->admixture && !opt->eta_constrained;
	hdr[4] = opt->admixture; hdr[5] = mod->n_iter;
	fwrite(hdr, 4, 6, fp);
	fwrite(&mod->logL, 8, 1, fp);
	fwrite(eta, 8, (size_t)mod->eta_len, fp);
	fwrite(p, 8, np, fp);
	fwrite(post, 8, (size_t)dat->I * mod->K, fp);
	fclose(fp);
	free(eta); free(p); free(post);
	return NO_ERROR;
}

/* bookkeeping of one finished fit (reference multiclust.c:534-627): counters,
 * best-so-far test, result files, the per-initialisation stdout line.
 * `write_best` writes the files of the current fit (it is called only when the
 * fit improves on everything seen so far, all K included) */
static int record_fit(options *opt, data *dat, model *mod, int i, int bootstrap,
	int (*write_best)(options *, data *, model *, void *), void *ctx)
{
	int err;

	if (mod->converged)
		mod->ever_converged = 1;
	if (mod->converged || (!mod->n_init && mod->time_stop)) {
		mod->n_total_iter += mod->n_iter;
		if (mod->n_max_iter < mod->n_iter)
			mod->n_max_iter = mod->n_iter;
		mod->n_init++;
	}
	if (mod->converged && converged(opt, mod, mod->first_max_logL)) {
		mod->n_maxll_times++;
	} else if (mod->converged && mod->logL > mod->first_max_logL) {
		mod->n_maxll_times = 1;
		mod->first_max_logL = mod->logL;
		mod->n_maxll_init = mod->n_init;
	}
	if (mod->trace)
		fprintf(mod->trace, "fit %d %d logL=%.17g converged=%d stopped=%d "
			"iter_stop=%d n_iter=%d pindex=%d\n", mod->K, i, mod->logL,
			mod->converged, mod->stopped, mod->iter_stop, mod->n_iter,
			mod->pindex);
	if (mod->logL > mod->max_logL) {
		mod->max_logL = mod->logL;
		mod->aic = aic(mod);
		mod->bic = bic(dat, mod);
		/* save the estimates if this is H0 fitted to the observed data
		 * (reference multiclust.c:562-581) */
		if (!bootstrap && opt->n_bootstrap && mod->K == mod->null_K)
			for (int r = 0; r < mod->n_gpus; r++)
				gpu_check(mod, mc_save_mle(mod->gpus[r], mod->pindex), "mc_save_mle");
		if (!bootstrap && opt->write_files && (err = write_best(opt, dat, mod, ctx)))
			return err;
	}
	if (!bootstrap && opt->verbosity > QUIET && opt->write_files)
		fprintf(stdout, "K = %d, initialization = %d: %f (%s) in %3d "
			"iterations, %02d:%02d:%02d (%f; %d), seed: %u\n",
			mod->K, i, mod->logL,
			mod->converged ? "converged" : "not converged",
			mod->n_iter, (int)(mod->seconds_run / 3600),
			(int)((((int)mod->seconds_run) % 3600) / 60),
			(((int)mod->seconds_run) % 60), mod->max_logL,
			mod->n_maxll_times, opt->seed);
	return NO_ERROR;
}

/* the result files of a fit whose state is in mod->eta_host / p_host /
 * post_host, dat->I_K and mod->count_K (write_file.c:203-732) */
static int write_result_files(options *opt, data *dat, model *mod)
{
	int err;

	if (opt->admixture) {
		if ((err = write_file_detail(opt, dat, mod))
			|| (err = popq_admix(opt, dat, mod))
			|| (err = indivq_admix(opt, dat, mod)))
			return err;
	} else {
		if ((err = write_file_detail(opt, dat, mod))
			|| (err = popq_mix(opt, dat, mod))
			|| (err = indivq_mix(opt, dat, mod)))
			return err;
	}
	return NO_ERROR;
}

/* sequential mode: the fit's state is still on the device */
static int write_best_from_device(options *opt, data *dat, model *mod, void *ctx)
{
	int err;

	const double t0 = wall_now();

	(void)ctx;
	if ((err = fetch_results(opt, dat, mod)))
		return err;
	if (opt->admixture)
		partition_admixture(dat, mod);
	else
		partition_mixture(dat, mod);
	err = write_result_files(opt, dat, mod);
	g_t_write += wall_now() - t0;
	return err;
}

int record_fit_public(options *opt, data *dat, model *mod, int i,
	int (*write_best)(options *, data *, model *, void *), void *ctx)
{
	return record_fit(opt, dat, mod, i, 0, write_best, ctx);
}

int write_result_files_public(options *opt, data *dat, model *mod)
{
	return write_result_files(opt, dat, mod);
}

int dump_state_public(options *opt, data *dat, model *mod, int init, const char *tag,
	int slot)
{
	return dump_state(opt, dat, mod, init, tag, slot);
}

/* reference multiclust.c:471-660 */
int maximize_likelihood(options *opt, data *dat, model *mod, int bootstrap)
{
	int err;

	mod->first_max_logL = -INFINITY;
	mod->n_init = 0;
	mod->n_total_iter = 0;
	mod->n_maxll_times = 0;
	mod->n_maxll_init = -1;
	mod->n_max_iter = 0;
	mod->time_stop = 0;
	mod->ever_converged = 0;
	mod->start = clock();

	for (int i = 0; opt->n_seconds || i < opt->n_init; i++) {
		mod->logL = 0.0;
		mod->converged = 0;
		mod->stopped = 0;
		mod->iter_stop = 0;
		if (mod->trace)
			fprintf(mod->trace, "init %d %d\n", mod->K, i);
		double t0 = wall_now();
		if ((err = initialize_model(opt, dat, mod)))
			return err;
		g_t_init += wall_now() - t0;
		if (opt->dump_prefix && (err = dump_state(opt, dat, mod, i, "start", mod->tindex)))
			return err;
		t0 = wall_now();
		em(opt, dat, mod);
		g_t_em += wall_now() - t0;
		if (opt->dump_prefix && (err = dump_state(opt, dat, mod, i, "final", mod->pindex)))
			return err;

		t0 = wall_now();
		if ((err = record_fit(opt, dat, mod, i, bootstrap, write_best_from_device, NULL)))
			return err;
		g_t_record += wall_now() - t0;
		if (mod->K == 1)
			break;
		if (mod->time_stop)
			break;
	}
	return NO_ERROR;
}

/* reference multiclust.c:715-793 */
void print_model_state(options *opt, data *dat, model *mod, int diff, int newline)
{
	if (opt->compact) {
		fprintf(stdout, "%s %s %s %d %u %e %e %e %e %f %f %f ", opt->filename,
			opt->accel_abbreviation, opt->admixture ? "admix" : "mix",
			mod->K, opt->seed, opt->eta_lower_bound, opt->p_lower_bound,
			opt->abs_error, opt->rel_error, mod->max_logL, aic(mod),
			bic(dat, mod));
		fprintf(stdout, "ND ");
		fprintf(stdout, "%s %02d:%02d:%02d %d %d %d %d",
			mod->ever_converged ? "converged" : "not", diff / 3600,
			(diff % 3600) / 60, diff % 60, mod->n_total_iter, mod->n_init,
			mod->n_maxll_init, mod->n_maxll_times);
		if (mod->time_stop)
			fprintf(stdout, " time");
		if (newline)
			fprintf(stdout, "\n");
		return;
	}
	fprintf(stdout, "Dataset: %s\n", opt->filename);
	fprintf(stdout, "Method/Model: %s, %s, K=%d\n", opt->accel_abbreviation,
		opt->admixture ? "admix" : "mix", mod->K);
	fprintf(stdout, "Convergence: ae=%e, re=%e\n", opt->abs_error, opt->rel_error);
	fprintf(stdout, "Bounds: e=%e, p=%e\n", opt->eta_lower_bound, opt->p_lower_bound);
	fprintf(stdout, "Total number of iterations: %d\n", mod->n_total_iter);
	fprintf(stdout, "Total time: %02d:%02d:%02d\n", diff / 3600, (diff % 3600) / 60,
		diff % 60);
	fprintf(stdout, "Iteration of max log likelihood: %d of %d\n",
		mod->n_maxll_init, mod->n_init);
	fprintf(stdout, "Number of times reach max log likelihood: %d\n", mod->n_maxll_times);
	fprintf(stdout, "Maximum log likelihood: %f\n", mod->max_logL);
	fprintf(stdout, "AIC: %f\n", aic(mod));
	fprintf(stdout, "BIC: %f\n", bic(dat, mod));
	fprintf(stdout, "Converged: %s\n", mod->ever_converged ? "yes" : "no");
	if (mod->time_stop)
		fprintf(stdout, "WARNING: Fitting stopped because ran out of time\n");
}

/* reference multiclust.c:365-450 */
int estimate_model(options *opt, data *dat, model *mod, int bootstrap)
{
	const clock_t start = clock();
	double min_aic = INFINITY, min_bic = INFINITY;
	int err;

	mod->max_logL = -INFINITY;
	mod->max_logL_H0 = -INFINITY;
	/* bootstrapping compares H0: K = null_K with Ha: K = alt_K */
	mod->K = opt->n_bootstrap ? mod->null_K : opt->min_K;
	dat->max_M = dat->M;
	for (;;) {
		if (dat->max_M < mod->K)
			dat->max_M = mod->K;
		double t0 = wall_now();
		if ((err = allocate_model_for_k(opt, mod, dat)))
			return err;
		g_t_plan += wall_now() - t0;
		if ((err = maximize_likelihood(opt, dat, mod, bootstrap)))
			return err;
		if (opt->n_repeat == 1 && opt->verbosity)
			print_model_state(opt, dat, mod,
				(int)(((double)clock() - start) / CLOCKS_PER_SEC), 1);
		if (opt->n_bootstrap && mod->K == mod->null_K)
			mod->max_logL_H0 = mod->max_logL;
		if (min_aic > mod->aic) {
			min_aic = mod->aic;
			mod->aic_K = mod->K;
		}
		if (min_bic > mod->bic) {
			min_bic = mod->bic;
			mod->bic_K = mod->K;
		}
		t0 = wall_now();
		free_model_data(mod, opt);
		g_t_plan += wall_now() - t0;
		if (opt->n_bootstrap && mod->K == mod->null_K)
			mod->K = mod->alt_K;
		else if (!opt->n_bootstrap && mod->K < opt->max_K)
			mod->K++;
		else
			break;
	}
	/* the test statistic (reference multiclust.c:434-449) */
	if (opt->n_bootstrap) {
		const double diff = mod->max_logL - mod->max_logL_H0;

		if (diff <= 0)
			return mmessage(ERROR_MSG, INTERNAL_ERROR, "Null hypothesis likelihood "
				"exceeds alternative hypothesis likelihood.  Try increasing "
				"number of initializations (command-line option -n)\n");
		if (!bootstrap)
			mod->ts_obs = diff;
		else
			mod->ts_bs = diff;
	}
	return NO_ERROR;
}

/* reference multiclust.c:675-708 */
static int run_bootstrap(options *opt, data *dat, model *mod)
{
	int ntime = 0, err = NO_ERROR;

	for (int i = 0; i < opt->n_bootstrap; i++) {
		fprintf(stdout, "Bootstrap dataset %d (of %d):", i + 1, opt->n_bootstrap);
		if ((err = parametric_bootstrap(opt, dat, mod)))
			break;
		if ((err = estimate_model(opt, dat, mod, 1)))
			break;
		if (mod->ts_bs >= mod->ts_obs)
			ntime++;
		fprintf(stdout, " test statistics bs=%f obs=%f (%f)\n", mod->ts_bs, mod->ts_obs,
			(double)ntime / (i + 1));
	}
	/* integer division, as in the reference (multiclust.c:703) */
	mod->pvalue = ntime / opt->n_bootstrap;
	const int err2 = cleanup_parametric_bootstrap(dat, mod);
	return err ? err : err2;
}

/* reference multiclust.c:201-347: `-w n <r> [t <min>] [m <min>]` repeats the
 * whole estimation (no result files) at least r times and until t minutes have
 * passed, at most m minutes, prints one compact line per repetition and the
 * statistics over the repetitions.  Features outside this program's flag set
 * keep their neutral values in the output: no target log likelihood (-u: "NA",
 * "u=(0.000000,0)", 0 "reach target") and no adjusted Rand index (-A: 0). */
static int timed_model_estimation(options *opt, data *dat, model *mod)
{
	const clock_t start = clock();
	int enough_time = opt->repeat_seconds ? 0 : 1;
	double esec = 0;
	int max_iter = 0, max_init = 0, err;
	double sum_init = 0, sum_init2 = 0, sum_iter = 0, sum_iter2 = 0;
	double sum_aic_K = 0, sum_aic_K2 = 0, sum_bic_K = 0, sum_bic_K2 = 0;
	double sum_ll = 0, sum_ll2 = 0;
	const double sum_ar = 0, sum_ar2 = 0, arand = 0;
	double max_ll = -INFINITY, min_aic = 0, min_bic = 0, max_ar = -1;
	double first_ll = -INFINITY, max_ll_rand = 0;
	int first_hit_index = 0, converged_repeats = 0, n_repeats = 0;
	const int target_reached = 0;

	while (n_repeats < opt->n_repeat || !enough_time) {
		if ((err = estimate_model(opt, dat, mod, 0)))
			return err;
		if (mod->n_init > max_init)
			max_init = mod->n_init;
		if (mod->n_max_iter > max_iter)
			max_iter = mod->n_max_iter;
		if (mod->max_logL > max_ll) {
			max_ll = mod->max_logL;
			min_aic = mod->aic;
			min_bic = mod->bic;
			max_ll_rand = arand;
			if (!converged(opt, mod, first_ll)) {
				first_ll = mod->max_logL;
				first_hit_index = n_repeats;
			}
		}
		if (arand > max_ar)
			max_ar = arand;
		sum_init += mod->n_init;
		sum_init2 += mod->n_init * mod->n_init;
		sum_iter += mod->n_total_iter;
		sum_iter2 += mod->n_total_iter * mod->n_total_iter;
		sum_aic_K += mod->aic_K;
		sum_aic_K2 += mod->aic_K * mod->aic_K;
		sum_bic_K += mod->bic_K;
		sum_bic_K2 += mod->bic_K * mod->bic_K;
		sum_ll += mod->max_logL;
		sum_ll2 += mod->max_logL * mod->max_logL;
		n_repeats++;
		if (mod->ever_converged)
			converged_repeats++;

		esec = ((double)clock() - start) / CLOCKS_PER_SEC;
		if (opt->verbosity > SILENT) {
			print_model_state(opt, dat, mod,
				(int)(((double)clock() - start) / CLOCKS_PER_SEC), 0);
			fprintf(stdout, " %f %f %d %d", esec, esec / n_repeats,
				target_reached, converged_repeats);
			fprintf(stdout, " %f", max_ll);
			fprintf(stdout, " NA");
			fprintf(stdout, " %d %d %d %d", 0, n_repeats, opt->n_repeat,
				opt->repeat_seconds);
			fprintf(stdout, "\n");
		}
		if (!enough_time || opt->max_repeat_seconds) {
			if (!enough_time && esec > opt->repeat_seconds)
				enough_time = 1;
			if (opt->max_repeat_seconds && esec > opt->max_repeat_seconds)
				break;
		}
	}

	if (opt->verbosity >= SILENT) {
		fprintf(stdout, "Data, Method, Model: %s, %s, %s\n", opt->filename,
			opt->accel_abbreviation,
			opt->admixture && opt->eta_constrained
			? "admix constrained" : opt->admixture ? "admix" : "mix");
		fprintf(stdout, "Run: %e %e %e %e n=%d i=%d u=(%f,%d) w=(%d,%d)\n",
			opt->abs_error, opt->rel_error, opt->eta_lower_bound,
			opt->p_lower_bound, opt->n_init, opt->n_init_iter, 0.0, 0,
			opt->n_repeat, opt->repeat_seconds);
		fprintf(stdout, "Number of repetitions: %d of %d requested, %d converged, "
			"%d reach target\n", n_repeats, opt->n_repeat, converged_repeats,
			target_reached);
		fprintf(stdout, "Average time: %fs (total: %fs; target: %d)\n",
			esec / n_repeats, esec, opt->repeat_seconds);
		fprintf(stdout, "Average log likelihood: %f (+/- %f)\n", sum_ll / n_repeats,
			sqrt((sum_ll2 - sum_ll * sum_ll / n_repeats) / (n_repeats - 1)));
		fprintf(stdout, "Maximum log likelihood: %f first hit at run %d (AIC %f; "
			"BIC %f; RAND: %f)\n", max_ll, first_hit_index, min_aic, min_bic,
			max_ll_rand);
		fprintf(stdout, "Adjusted RAND: avg = %f +/- %f; max = %f\n",
			sum_ar / n_repeats,
			sqrt((sum_ar2 - sum_ar * sum_ar / n_repeats) / (n_repeats - 1)), max_ar);
		if (opt->max_K != opt->min_K) {
			fprintf(stdout, "Average K (AIC): %f (+/- %f)\n", sum_aic_K / n_repeats,
				sqrt((sum_aic_K2 - sum_aic_K * sum_aic_K / n_repeats)
					/ (n_repeats - 1)));
			fprintf(stdout, "Average K (BIC): %f (+/- %f)\n", sum_bic_K / n_repeats,
				sqrt((sum_bic_K2 - sum_bic_K * sum_bic_K / n_repeats)
					/ (n_repeats - 1)));
		} else {
			fprintf(stdout, "Total initializations, iterations: %d, %d\n",
				(int)sum_init, (int)sum_iter);
			fprintf(stdout, "Average initializations: %f (+/- %f) [%e, %e]\n",
				sum_init / n_repeats,
				sqrt((sum_init2 - sum_init * sum_init / n_repeats)
					/ (n_repeats - 1)), sum_init2, sum_init);
			fprintf(stdout, "Average iterations: %f (+/- %f) [%e %e]\n",
				sum_iter / sum_init,
				sqrt((sum_iter2 - sum_iter * sum_iter / sum_init)
					/ (sum_init - 1)), sum_iter2, sum_iter);
			fprintf(stdout, "Maximum initializations: %d\n", max_init);
			fprintf(stdout, "Maximum iterations: %d\n", max_iter);
		}
	}
	return NO_ERROR;
}

#ifndef MC_HOST_NO_MAIN
int main(int argc, const char **argv)
{
	options *opt = NULL;
	data *dat = NULL;
	model *mod = NULL;
	int err;

	if ((err = make_options(&opt)) || (err = make_data(&dat))
		|| (err = make_model(&mod)))
		goto done;
	if ((err = parse_options(opt, dat, argc, argv)))
		goto done;
	g_timing = opt->timing;
	const double t_start = wall_now();
	start_device_contexts(opt);
	if ((err = read_file(opt, dat)))
		goto done;
	const double t_read = wall_now();
	if (opt->verbosity >= TALKATIVE)
		mmessage(INFO_MSG, NO_ERROR, "Finished reading data: %u %u-ploid "
			"individuals at %u loci.\n", dat->I, dat->ploidy, dat->L);
	if (opt->parse_only) {
		/* the recoded genotypes, without touching a device */
		mcb_data d = { dat->I, dat->L, dat->ploidy, dat->numpops,
			dat->uniquealleles, dat->nreal, dat->labels, dat->label_off,
			NULL, dat->codes };
		d.locale = malloc(sizeof(int32_t) * (size_t)dat->I);
		for (int i = 0; i < dat->I; i++)
			d.locale[i] = dat->idv[i].locale;
		err = mcb_write(opt->parse_only, &d) ? FILE_OPEN_ERROR : NO_ERROR;
		free(d.locale);
		goto done;
	}
	if ((err = synchronize(opt, dat, mod)))
		goto done;
	if (opt->trace_file && !(mod->trace = fopen(opt->trace_file, "w"))) {
		err = message(stderr, __FILE__, __func__, __LINE__, ERROR_MSG,
			FILE_OPEN_ERROR, opt->trace_file);
		goto done;
	}
	if (opt->shard_fits && (opt->n_gpus > 1 || opt->fits_per_gpu > 1)) {
		err = estimate_model_sharded(opt, dat, mod);
		if (!err && opt->parallel)
			printf("%f\n", mod->max_logL);
		goto done;
	}
	if ((err = upload_data(opt, dat, mod)))
		goto done;
	const double t_upload = wall_now();
	if (opt->n_repeat > 1)
		err = timed_model_estimation(opt, dat, mod);
	else if (opt->n_repeat == 1)
		err = estimate_model(opt, dat, mod, 0);
	if (!err && opt->parallel)
		printf("%f\n", mod->max_logL);
	/* optionally run a bootstrap (reference multiclust.c:147-155) */
	if (!err && opt->n_bootstrap && !(err = run_bootstrap(opt, dat, mod)))
		fprintf(stdout, "p-value to reject H0: K=%d is %f\n", mod->null_K, mod->pvalue);
	if (g_timing)
		fprintf(stderr, "timing (s): read %.3f, upload %.3f, plan per K %.3f, other %.3f, "
			"initialise %.3f, em %.3f (%d iterations), fetch+write %.3f, total %.3f\n",
			t_read - t_start, t_upload - t_read, g_t_plan,
			wall_now() - t_upload - g_t_init - g_t_em - g_t_record - g_t_plan, g_t_init,
			g_t_em, mod->n_total_iter ? mod->n_total_iter : mod->n_iter, g_t_write,
			wall_now() - t_start);
done:
	free_model(mod, opt);
	free_options(opt);
	free_data(dat);
	return err;
}
#endif
