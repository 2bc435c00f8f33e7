/*
 * multiclust.h -- host side of the B200 MULTICLUST drop-in (C17).
 *
 * Mirrors the interface of the reference's hot path and of the code either
 * side of it (reference multiclust.h:146-388): the same three objects
 * `options`, `data`, `model`, the same function names and argument order, the
 * same error codes.  What differs is the storage: genotypes are 8-bit allele
 * codes on a flat [I][L][ploidy] layout, parameters live in HBM behind the
 * C ABI of include/mc_cuda.h (one `mc_ctx` per model), and the four-dimensional
 * scratch array diklm (reference multiclust.c:1197) does not exist.
 *
 * Everything that is policy stays here in C exactly as in the reference: the
 * rotation of the three parameter slots, accept / reject of accelerated steps,
 * stop() / converged(), the exit(0) rules.  Everything that touches I*L data
 * or the parameter vectors is a call into libmc_cuda.so.  There is no CPU
 * implementation of the EM path in this program.
 */
#ifndef MC_HOST_MULTICLUST_H
#define MC_HOST_MULTICLUST_H

#include <stdint.h>
#include <stdio.h>
#include <time.h>

#include "mc_cuda.h"
#include "mc_comm.h"
#include "mc_rand.h"

/* error codes: same numbering as reference message.h:17-42 */
enum {
	NO_ERROR, CUSTOM_ERROR, NO_DATA, MEMORY_ALLOCATION, FILE_NOT_FOUND,
	FILE_OPEN_ERROR, END_OF_FILE, FILE_FORMAT_ERROR, INVALID_CMDLINE,
	INVALID_CMD_OPTION, INVALID_CMD_ARGUMENT, INVALID_USER_SETUP,
	INTERNAL_MISMATCH, INTERNAL_ERROR, CLUSTER_SIZE_OVERFLOW,
	STATE_SPACE_OVERFLOW, OUT_OF_TIME, MEMORY_USAGE_LIMIT, MEMCPY_ERROR,
	GPU_ERROR,		/* new: a libmc_cuda call failed */
	NUM_ERRORS
};
enum { NO_MSG, INFO_MSG, DEBUG_MSG, WARNING_MSG, ERROR_MSG };
/* verbosity levels, reference message.h:70-78 */
enum { ABSOLUTE_SILENCE, SILENT, QUIET, MINIMAL, RESTRAINED, TALKATIVE, VERBOSE, DEBUG };

/* acceleration schemes (-s): 0 none, 1-3 SQUAREM, 4-6 quasi-Newton q = s - 3 */
enum { NONE, SQS1, SQS2, SQS3, QN, NUM_ACCELERATION_METHODS };

#define MISSING (-9)		/* default missing allele in the input */
#define MC_CODE_MISSING 255	/* its 8-bit code */

int message(FILE *fp, const char *file, const char *fxn, int line, int type,
	int id, const char *fmt, ...);
#define mmessage(type, err, ...) \
	message(stderr, __FILE__, __func__, __LINE__, (type), (err), __VA_ARGS__)

typedef struct _options options;
typedef struct _data data;
typedef struct _model model;
typedef struct _indiv indiv;

/* run options (reference multiclust.h:156-217, the fields this path uses) */
struct _options {
	int eta_constrained;		/* -c */
	int admixture;			/* -a */
	int n_init;			/* -n */
	int n_rand_em_init;		/* -m: accepted, inert (SURVEY.md finding 2) */
	int max_iter;			/* -T / -C */
	int min_K, max_K;		/* -1 / -2 / -k */
	int missing_value;		/* --missing */
	double lower_bound;		/* --bound */
	double rel_error, abs_error;	/* -e / -E */
	const char *filename;		/* -f */
	const char *filename_file;	/* basename of -f */
	const char *path;		/* -d */
	const char *outfile_name;	/* -o */
	int R_format;			/* -R */
	int interleaved;
	unsigned int seed;		/* -r */
	int accel_scheme;		/* -s */
	int do_projection;		/* 0 after --projection */
	char accel_name[64];
	char accel_abbreviation[16];
	int q;
	int n_init_iter;		/* -i */
	unsigned int n_seconds;		/* -t */
	int adjust_step;		/* -g */
	int verbosity;			/* -v */
	int compact;
	double eta_lower_bound, p_lower_bound;
	int n_bootstrap;		/* -b n: parametric bootstrap of H0: K-1 against Ha: K */
	int n_repeat;			/* -w n */
	int repeat_seconds;		/* -w t <minutes>: keep repeating until then */
	int max_repeat_seconds;		/* -w m <minutes>: stop repeating after */
	int write_files;
	int parallel;			/* -M */
	/* new in this program */
	int device;			/* --device: CUDA ordinal (default 0) */
	int n_gpus;			/* --gpus: shard individuals over n GPUs */
	const char *trace_file;		/* --trace: every log likelihood, %.17g */
	const char *dump_prefix;	/* --dump: binary parameters per fit */
	const char *parse_only;		/* --parse-only: MCB1 file to write */
	int timing;			/* --timing: wall-clock seconds per phase on stderr */
	int shard_fits;			/* --shard-fits: with --gpus N, deal whole fits
					 * (K, initialisation) to the devices instead of
					 * sharding the individuals of one fit */
	int fits_per_gpu;		/* --fits-per-gpu M: fits in flight on every device
					 * (one host thread, context and stream each): the
					 * launch latency of one small fit hides behind the
					 * kernels of the others */
};

struct _indiv {
	char *name;
	int locale;
};

/* the data (reference multiclust.h:225-251) on the flat layout */
struct _data {
	int I, L, M, ploidy, missing_data;
	int32_t *uniquealleles;	/* [L] allele slots, incl. the phantom slot */
	int32_t *nreal;		/* [L] labelled alleles */
	int32_t *allele_off;	/* [L+1] prefix sums of uniquealleles */
	int32_t *labels;	/* ascending allele labels, locus after locus */
	int64_t *label_off;	/* [L+1] prefix sums of nreal */
	uint8_t *codes;		/* [I][L][ploidy]: 0..nreal-1, 255 = missing */
	uint8_t *codes_orig;	/* the observed data while `codes` is a bootstrap sample */
	indiv *idv;
	int *I_K;		/* partition of the individuals */
	int numpops;
	char **pops;
	int *i_p;
	int max_M;
};

/* the model (reference multiclust.h:259-357) */
struct _model {
	int K;
	int no_parameters;
	int pindex, findex, tindex;	/* parameter slots on the device */
	int delta_index;
	double *A, *Ainv, *cutu;	/* quasi-Newton q x q work space */
	double logL;
	int accel_step;
	int converged, stopped;
	int *count_K;
	int n_iter;
	int ever_converged;
	double max_logL, first_max_logL;
	double aic, bic;
	int n_init, n_total_iter, n_maxll_init, n_maxll_times;
	int time_stop, iter_stop, n_max_iter;
	clock_t start;
	double seconds_run;
	int aic_K, bic_K;
	/* parametric bootstrap (reference multiclust.h:340-352) */
	int null_K, alt_K;
	double max_logL_H0, ts_obs, ts_bs, pvalue;
	/* device side */
	mc_ctx *gpu;			/* = gpus[0] */
	mc_ctx **gpus;			/* --gpus: one context per device, individuals
					 * [row_first[r], row_first[r+1]) on device r */
	int n_gpus;
	int *row_first;
	mc_comm *comm;			/* NCCL exchange (n_gpus > 1) */
	int64_t T;			/* sum of allele slots */
	int64_t eta_len;		/* I*K or K */
	/* host copies fetched for the writers */
	double *eta_host, *p_host, *post_host;
	double *popq_host;		/* [numpops][K] locale sums of the posterior (device) */
	FILE *trace;			/* --trace */
	mcr_state *rng;			/* the rand() stream of the initialisers */
	int no_exit;			/* 1: record the reference's exit(0) conditions in
					 * `aborted` (1 NaN, 2 decrease) instead of exiting */
	int aborted;
	double abort_ll, abort_prev;
};

/* ---- objects (reference multiclust.c:902-1380) ---- */
int make_options(options **opt);
int make_data(data **dat);
int make_model(model **mod);
void free_options(options *opt);
void free_data(data *dat);
void free_model(model *mod, options *opt);
void free_model_data(model *mod, options *opt);
int parse_options(options *opt, data *dat, int argc, const char **argv);
int synchronize(options *opt, data *dat, model *mod);
int allocate_model_for_k(options *opt, model *mod, data *dat);
int estimate_model(options *opt, data *dat, model *mod, int bootstrap);
int maximize_likelihood(options *opt, data *dat, model *mod, int bootstrap);
void print_model_state(options *opt, data *dat, model *mod, int diff, int newline);
/* --gpus N --shard-fits: whole fits dealt to the devices (shard_fits.c) */
int estimate_model_sharded(options *opt, data *dat, model *mod);
int record_fit_public(options *opt, data *dat, model *mod, int i,
	int (*write_best)(options *, data *, model *, void *), void *ctx);
int write_result_files_public(options *opt, data *dat, model *mod);
int dump_state_public(options *opt, data *dat, model *mod, int init, const char *tag,
	int slot);
void fprint_usage(FILE *fp, const char *cmd);

/* ---- input (reference read_file.c) ---- */
int read_file(options *opt, data *dat);
int upload_data(options *opt, data *dat, model *mod);
void start_device_contexts(options *opt);	/* CUDA start-up overlapped with the parse */

/* ---- EM hot path (reference multiclust.h:371-388) ---- */
int initialize_model(options *opt, data *dat, model *mod);
/* bootstrap.c:31-66: one bootstrap sample in place of the data / the data back */
int parametric_bootstrap(options *opt, data *dat, model *mod);
int cleanup_parametric_bootstrap(data *dat, model *mod);
void em(options *opt, data *dat, model *mod);
int em_step(options *opt, data *dat, model *mod);
int em_2_steps(model *mod, data *dat, options *opt);
int stop(options *opt, model *mod, double loglik);
int converged(options *opt, model *mod, double loglik);
double log_likelihood(options *opt, data *dat, model *mod, int which);
int accelerated_em_step(options *opt, data *dat, model *mod);
double aic(model *mod);
double bic(data *dat, model *mod);

/* ---- output (reference write_file.c) ---- */
int gather_state(options *opt, data *dat, model *mod, int slot, double *eta,
	double *p, double *post);
int fetch_results(options *opt, data *dat, model *mod);
int write_file_detail(options *opt, data *dat, model *mod);
void partition_admixture(data *dat, model *mod);
void partition_mixture(data *dat, model *mod);
int popq_admix(options *opt, data *dat, model *mod);
int indivq_admix(options *opt, data *dat, model *mod);
int popq_mix(options *opt, data *dat, model *mod);
int indivq_mix(options *opt, data *dat, model *mod);

/* abort with the library's message when a device call fails */
void gpu_check(model *mod, int rc, const char *what);

#endif
