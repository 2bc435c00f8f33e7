/*
 * shard_fits.c -- `--gpus N --shard-fits`: the independent fits of a run (one
 * per K of the -1/-2 sweep and per random initialisation -n) dealt to N
 * devices, results identical to the sequential program.
 *
 * The reference runs the fits one after the other (estimate_model,
 * multiclust.c:365-450; maximize_likelihood, 471-660) and couples them in two
 * places only:
 *   1. the initialisers draw from ONE rand() stream that is never reseeded
 *      between fits (multiclust.c:516-531), so fit n+1 starts where fit n
 *      stopped drawing;
 *   2. the best-so-far bookkeeping (max_logL over all K, result files
 *      rewritten on every improvement, the per-fit stdout line, the
 *      n_maxll_* counters; multiclust.c:534-627) runs in fit order.
 * Both are replayed here on the host:
 *   1. the number of draws of a fit does not depend on its outcome --
 *      admixture: I*L*P draws whatever K (rnd_init.c:460-481); mixture: K
 *      centre draws plus re-draws on collision (rnd_init.c:205-217), which
 *      only need the draws themselves -- so the master walks the stream once,
 *      snapshots the generator state (mc_rand.h) in front of every fit, and
 *      each worker regenerates its own draws from its snapshot;
 *   2. workers store the outcome of every fit; the master consumes them in
 *      (K, initialisation) order through the same record_fit() the sequential
 *      loop uses.  A fit that cannot be a new maximum (its log likelihood does
 *      not exceed every earlier fit of the same worker) does not keep its
 *      parameters.
 * One host thread and one mc_ctx per worker (the C ABI's rule); a device may
 * carry several workers (--fits-per-gpu), each with its own stream, so that the
 * launch latency of one small fit hides behind the kernels of the others.  The
 * genotype codes are replicated for every worker, no collective.  The reference's two
 * exit(0) conditions (NaN, log-likelihood decrease; em_alg.c:113-121) are
 * recorded by the worker and raised by the master when the replay reaches
 * that fit, so everything an aborted sequential run would have written
 * before the abort is written here too.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "multiclust.h"

typedef struct {
	int K, init;
	mcr_state rng;			/* generator state in front of the fit */
	/* outcome */
	int done, err;
	double logL, seconds_run;
	int converged, stopped, iter_stop, n_iter, pindex;
	int aborted;
	double abort_ll, abort_prev;
	char *trace;			/* --trace text of the fit */
	size_t trace_len;
	/* kept only when the fit may be a new maximum */
	double *eta, *p, *post, *popq;
	int *I_K, *count_K;
} fit_job;

typedef struct {
	options *opt;
	data *dat;
	model *master;
	fit_job *jobs;
	int n_jobs, n_workers;
	pthread_mutex_t lock;
	pthread_cond_t cond;
} fit_pool;

typedef struct {
	fit_pool *pool;
	int rank;
	pthread_t thread;
	double t_ready, t_done;		/* --timing: context + data up, last fit finished */
} fit_worker;

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void job_release(fit_job *job)
{
	free(job->eta); free(job->p); free(job->post); free(job->popq);
	free(job->I_K); free(job->count_K); free(job->trace);
	job->eta = job->p = job->post = job->popq = NULL;
	job->I_K = job->count_K = NULL;
	job->trace = NULL;
}

/* advance `rng` over the draws of one initialisation without keeping them */
static void skip_draws(const options *opt, const data *dat, int K, mcr_state *rng)
{
	if (opt->admixture) {
		/* I*L*P draws whatever K: one matrix-vector product */
		static uint32_t jump[MCR_LAG * MCR_LAG];
		static long long jump_n = -1;
		const long long n = (long long)dat->I * dat->L * dat->ploidy;
		uint32_t h[MCR_LAG];
		if (jump_n != n) {
			mcr_jump_matrix(n, jump);
			jump_n = n;
		}
		mcr_history(rng, h);
		mcr_apply(jump, h);
		mcr_from_history(rng, h);
	} else if (K > 1) {
		/* rnd_init.c:205-217, the draws only */
		int *center = malloc(sizeof *center * (size_t)K);
		for (int k = 0; k < K; k++) {
			int again;
			center[k] = mcr_next(rng) % dat->I;
			do {
				again = 0;
				for (int j = 0; j < k; j++)
					if (center[k] == center[j]) {
						center[k] = mcr_next(rng) % dat->I;
						again = 1;
						break;
					}
			} while (again);
		}
		free(center);
	}
}

/* one worker: its own device, model object and generator */
static void *worker_main(void *arg)
{
	fit_worker *w = arg;
	fit_pool *pool = w->pool;
	options *opt = pool->opt;
	data dat = *pool->dat;		/* private partition array below */
	model *mod = calloc(1, sizeof *mod);
	mcr_state rng;
	double best = -INFINITY;
	int err = NO_ERROR, K_alloc = 0;

	dat.I_K = malloc(sizeof *dat.I_K * (size_t)dat.I);
	mod->rng = &rng;
	mod->no_exit = 1;
	mod->max_logL = -INFINITY;
	mod->n_gpus = 1;
	mod->gpus = calloc(1, sizeof *mod->gpus);
	mod->row_first = calloc(2, sizeof *mod->row_first);
	mod->row_first[1] = dat.I;
	mod->T = dat.allele_off[dat.L];
	if (opt->accel_scheme >= QN) {
		mod->A = calloc((size_t)opt->q * opt->q, sizeof(double));
		mod->Ainv = calloc((size_t)opt->q * opt->q, sizeof(double));
		mod->cutu = calloc((size_t)opt->q, sizeof(double));
	}
	/* workers r, r + n_gpus, ... share device r % n_gpus */
	if (mc_create(&mod->gpus[0], opt->device + w->rank % opt->n_gpus)) {
		err = mmessage(ERROR_MSG, GPU_ERROR, "%s\n", mc_last_error(NULL));
	} else if (mc_set_data(mod->gpus[0], dat.I, dat.L, dat.ploidy,
			dat.uniquealleles, dat.codes)) {
		err = mmessage(ERROR_MSG, GPU_ERROR, "%s\n", mc_last_error(mod->gpus[0]));
	}
	mod->gpu = mod->gpus[0];
	w->t_ready = now_s();

	/* jobs are dealt round-robin, so a worker sees its jobs in run order */
	for (int j = w->rank; j < pool->n_jobs; j += pool->n_workers) {
		fit_job *job = &pool->jobs[j];

		if (!err && job->K != K_alloc) {
			if (K_alloc)
				free_model_data(mod, opt);
			mod->K = job->K;
			err = allocate_model_for_k(opt, mod, &dat);
			K_alloc = job->K;
		}
		if (!err) {
			rng = job->rng;
			mod->logL = 0.0;
			mod->converged = mod->stopped = mod->iter_stop = 0;
			mod->aborted = 0;
			mod->start = clock();
			if (opt->trace_file)
				mod->trace = open_memstream(&job->trace, &job->trace_len);
			if (mod->trace)
				fprintf(mod->trace, "init %d %d\n", job->K, job->init);
			err = initialize_model(opt, &dat, mod);
			if (!err && opt->dump_prefix)
				err = dump_state_public(opt, &dat, mod, job->init, "start",
					mod->tindex);
			if (!err)
				em(opt, &dat, mod);
			if (!err && opt->dump_prefix && !mod->aborted)
				err = dump_state_public(opt, &dat, mod, job->init, "final",
					mod->pindex);
			if (mod->trace)
				fclose(mod->trace);
			mod->trace = NULL;
		}
		if (!err) {
			job->logL = mod->logL;
			job->seconds_run = mod->seconds_run;
			job->converged = mod->converged;
			job->stopped = mod->stopped;
			job->iter_stop = mod->iter_stop;
			job->n_iter = mod->n_iter;
			job->pindex = mod->pindex;
			job->aborted = mod->aborted;
			job->abort_ll = mod->abort_ll;
			job->abort_prev = mod->abort_prev;
			if (!mod->aborted && opt->write_files && mod->logL > best) {
				best = mod->logL;
				err = fetch_results(opt, &dat, mod);
				if (!err) {
					if (opt->admixture)
						partition_admixture(&dat, mod);
					else
						partition_mixture(&dat, mod);
					job->eta = mod->eta_host;
					job->p = mod->p_host;
					job->post = mod->post_host;
					job->popq = mod->popq_host;
					mod->eta_host = mod->p_host = mod->post_host = mod->popq_host = NULL;
					job->I_K = malloc(sizeof(int) * (size_t)dat.I);
					job->count_K = malloc(sizeof(int) * (size_t)job->K);
					memcpy(job->I_K, dat.I_K, sizeof(int) * (size_t)dat.I);
					memcpy(job->count_K, mod->count_K,
						sizeof(int) * (size_t)job->K);
				}
			}
		}
		pthread_mutex_lock(&pool->lock);
		job->err = err;
		job->done = 1;
		pthread_cond_broadcast(&pool->cond);
		pthread_mutex_unlock(&pool->lock);
	}
	w->t_done = now_s();
	if (K_alloc)
		free_model_data(mod, opt);
	if (mod->gpus[0])
		mc_destroy(mod->gpus[0]);
	free(mod->gpus); free(mod->row_first);
	free(mod->A); free(mod->Ainv); free(mod->cutu);
	free(mod);
	free(dat.I_K);
	return NULL;
}

/* record_fit's callback: the files of the fit whose kept state the master's
 * model currently points at */
static int write_best_from_job(options *opt, data *dat, model *mod, void *ctx)
{
	fit_job *job = ctx;
	int *I_K = dat->I_K, err;

	if (!job->eta)		/* cannot happen: a global maximum is a worker maximum */
		return mmessage(ERROR_MSG, INTERNAL_ERROR, "fit K=%d init=%d kept no "
			"parameters\n", job->K, job->init);
	mod->eta_host = job->eta;
	mod->p_host = job->p;
	mod->post_host = job->post;
	mod->popq_host = job->popq;
	dat->I_K = job->I_K;
	memcpy(mod->count_K, job->count_K, sizeof(int) * (size_t)job->K);
	err = write_result_files_public(opt, dat, mod);
	dat->I_K = I_K;
	mod->eta_host = mod->p_host = mod->post_host = mod->popq_host = NULL;
	return err;
}

int estimate_model_sharded(options *opt, data *dat, model *mod)
{
	const clock_t start = clock();
	double min_aic = INFINITY, min_bic = INFINITY;
	fit_pool pool = { .opt = opt, .dat = dat, .master = mod };
	fit_worker *workers;
	mcr_state rng = *mod->rng;
	int err = NO_ERROR, j = 0;

	if (opt->n_seconds)
		return mmessage(ERROR_MSG, INVALID_USER_SETUP, "--shard-fits needs a "
			"fixed number of initialisations (-n), not a time limit (-t)\n");
	if (opt->n_repeat != 1)
		return mmessage(ERROR_MSG, INVALID_USER_SETUP, "--shard-fits does not "
			"combine with -w\n");

	/* the job list in run order, with the generator state in front of each */
	for (int K = opt->min_K; K <= opt->max_K; K++)
		pool.n_jobs += K == 1 ? 1 : opt->n_init;
	pool.jobs = calloc((size_t)pool.n_jobs, sizeof *pool.jobs);
	if (!pool.jobs)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "fit table\n");
	for (int K = opt->min_K; K <= opt->max_K; K++)
		for (int i = 0; i < (K == 1 ? 1 : opt->n_init); i++, j++) {
			pool.jobs[j].K = K;
			pool.jobs[j].init = i;
			pool.jobs[j].rng = rng;
			skip_draws(opt, dat, K, &rng);
		}
	*mod->rng = rng;

	const double t_start = now_s();
	pool.n_workers = opt->n_gpus * opt->fits_per_gpu;
	if (pool.n_workers > pool.n_jobs)
		pool.n_workers = pool.n_jobs;
	pthread_mutex_init(&pool.lock, NULL);
	pthread_cond_init(&pool.cond, NULL);
	workers = calloc((size_t)pool.n_workers, sizeof *workers);
	for (int r = 0; r < pool.n_workers; r++) {
		workers[r].pool = &pool;
		workers[r].rank = r;
		pthread_create(&workers[r].thread, NULL, worker_main, &workers[r]);
	}

	/* replay of estimate_model / maximize_likelihood in fit order */
	mod->max_logL = -INFINITY;
	mod->T = dat->allele_off[dat->L];
	dat->max_M = dat->M;
	j = 0;
	for (int K = opt->min_K; K <= opt->max_K && !err; K++) {
		mod->K = K;
		if (dat->max_M < K)
			dat->max_M = K;
		/* host-side part of allocate_model_for_k */
		mod->count_K = calloc((size_t)K, sizeof *mod->count_K);
		mod->eta_len = (opt->admixture && !opt->eta_constrained)
			? (int64_t)dat->I * K : K;
		mod->no_parameters = (!opt->admixture || opt->eta_constrained)
			? K - 1 : dat->I * (K - 1);
		for (int l = 0; l < dat->L; l++)
			mod->no_parameters += (dat->uniquealleles[l] - 1) * K;
		mod->first_max_logL = -INFINITY;
		mod->n_init = mod->n_total_iter = mod->n_maxll_times = 0;
		mod->n_maxll_init = -1;
		mod->n_max_iter = mod->time_stop = mod->ever_converged = 0;

		for (int i = 0; i < (K == 1 ? 1 : opt->n_init) && !err; i++, j++) {
			fit_job *job = &pool.jobs[j];

			pthread_mutex_lock(&pool.lock);
			while (!job->done)
				pthread_cond_wait(&pool.cond, &pool.lock);
			pthread_mutex_unlock(&pool.lock);
			if ((err = job->err))
				break;
			if (mod->trace && job->trace)
				fwrite(job->trace, 1, job->trace_len, mod->trace);
			if (job->aborted) {
				/* the reference's exit(0) (em_alg.c:113-121), raised
				 * where the sequential run would have raised it */
				if (mod->trace)
					fflush(mod->trace);
				if (job->aborted == 1)
					mmessage(ERROR_MSG, CUSTOM_ERROR, "nan\n");
				else
					mmessage(ERROR_MSG, CUSTOM_ERROR, "log likelihood "
						"decrease (%f < %f; %e)\n", job->abort_ll,
						job->abort_prev, (job->abort_ll
						- job->abort_prev) / job->abort_ll);
				fflush(NULL);
				_Exit(0);
			}
			mod->logL = job->logL;
			mod->seconds_run = job->seconds_run;
			mod->converged = job->converged;
			mod->stopped = job->stopped;
			mod->iter_stop = job->iter_stop;
			mod->n_iter = job->n_iter;
			mod->pindex = job->pindex;
			err = record_fit_public(opt, dat, mod, i, write_best_from_job, job);
			job_release(job);
		}
		if (!err) {
			if (opt->verbosity)
				print_model_state(opt, dat, mod,
					(int)(((double)clock() - start) / CLOCKS_PER_SEC), 1);
			if (min_aic > mod->aic) {
				min_aic = mod->aic;
				mod->aic_K = K;
			}
			if (min_bic > mod->bic) {
				min_bic = mod->bic;
				mod->bic_K = K;
			}
		}
		free(mod->count_K);
		mod->count_K = NULL;
	}
	for (int r = 0; r < pool.n_workers; r++)
		pthread_join(workers[r].thread, NULL);
	if (opt->timing) {
		double ready = 0, done = 0;
		for (int r = 0; r < pool.n_workers; r++) {
			if (workers[r].t_ready - t_start > ready)
				ready = workers[r].t_ready - t_start;
			if (workers[r].t_done - t_start > done)
				done = workers[r].t_done - t_start;
		}
		fprintf(stderr, "timing (s): %d workers on %d devices: contexts + data up after "
			"%.3f, %d fits done after %.3f (fits phase %.3f)\n", pool.n_workers,
			opt->n_gpus, ready, pool.n_jobs, done, done - ready);
	}
	for (j = 0; j < pool.n_jobs; j++)
		job_release(&pool.jobs[j]);
	free(workers);
	free(pool.jobs);
	pthread_mutex_destroy(&pool.lock);
	pthread_cond_destroy(&pool.cond);
	return err;
}
