/*
 * em_driver.c -- the EM driver of MULTICLUST on top of the CUDA C ABI.
 *
 * Control flow, slot rotation, accept/reject and stopping rules follow the
 * reference line by line in behaviour (em_alg.c:44-207, 1072-1171;
 * accel_em.c:35-551); every pass over the data or over the parameter vectors
 * is one mc_* call (include/mc_cuda.h).  Nothing here loops over individuals,
 * loci or alleles.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "multiclust.h"

static const char *accel_abbrev[NUM_ACCELERATION_METHODS] = { "EM", "S1", "S2", "S3", "Q" };

void gpu_check(model *mod, int rc, const char *what)
{
	if (rc == MC_OK)
		return;
	mmessage(ERROR_MSG, GPU_ERROR, "%s: %s\n", what, mc_last_error(mod->gpu));
	exit(GPU_ERROR);
}
#define GPU(call) gpu_check(mod, (call), #call)
/* the same call on every device of an individual-sharded fit (--gpus); the
 * calls only launch work, so the devices run concurrently */
#define GPUS(r, call) for (int r = 0; r < mod->n_gpus; r++) gpu_check(mod, (call), #call)

static void comm_exchange(model *mod)
{
	if (mc_comm_exchange(mod->comm) != MC_OK) {
		mmessage(ERROR_MSG, GPU_ERROR, "%s\n", mc_comm_last_error(mod->comm));
		exit(GPU_ERROR);
	}
}

static int eta_is_sharded(const options *opt)
{
	return opt->admixture && !opt->eta_constrained;
}

/* reference em_alg.c:163-182 */
int converged(options *opt, model *mod, double loglik)
{
	double abs_diff = 0, rel_diff = 0;
	int done = 1;

	if (opt->abs_error)
		abs_diff = fabs(loglik - mod->logL);
	if (opt->rel_error)
		rel_diff = abs_diff / fabs(mod->logL);
	if (opt->abs_error && abs_diff > opt->abs_error)
		done = 0;
	if (opt->rel_error && rel_diff > opt->rel_error)
		done = 0;
	if (done)
		mod->converged = 1;
	return done;
}

/* reference em_alg.c:145-161 */
static int stop_condition(options *opt, model *mod, double loglik)
{
	mod->seconds_run = ((double)clock() - mod->start) / CLOCKS_PER_SEC;
	if (opt->max_iter && mod->n_iter > opt->max_iter) {
		mod->iter_stop = 1;
		return 1;
	}
	if (opt->n_seconds && mod->seconds_run > opt->n_seconds) {
		mod->time_stop = 1;
		return 1;
	}
	return converged(opt, mod, loglik);
}

/* reference em_alg.c:101-143: counts the iteration, aborts the program on a
 * NaN or a decrease of the log likelihood (exit status 0, like the reference),
 * prints the -v trace */
int stop(options *opt, model *mod, double loglik)
{
	mod->n_iter++;
	if (mod->trace) {
		fprintf(mod->trace, "ll %d %d %.17g\n", mod->K, mod->n_iter, loglik);
		fflush(mod->trace);
	}
	if (isnan(loglik)) {
		if (mod->no_exit) {	/* a worker of the sharded multi-start mode */
			mod->aborted = 1;
			return mod->stopped = 1;
		}
		mmessage(ERROR_MSG, CUSTOM_ERROR, "nan\n");
		exit(0);
	}
	mod->stopped = stop_condition(opt, mod, loglik);
	if (loglik < mod->logL && !mod->stopped) {
		if (mod->no_exit) {
			mod->aborted = 2;
			mod->abort_ll = loglik;
			mod->abort_prev = mod->logL;
			return mod->stopped = 1;
		}
		mmessage(ERROR_MSG, CUSTOM_ERROR, "log likelihood decrease (%f < %f; %e)\n",
			loglik, mod->logL, (loglik - mod->logL) / loglik);
		exit(0);
	}
	if (opt->verbosity > MINIMAL) {
		fprintf(stderr, "%4d (%s", mod->n_iter, mod->accel_step
			? accel_abbrev[opt->accel_scheme < NUM_ACCELERATION_METHODS
				? opt->accel_scheme : NUM_ACCELERATION_METHODS - 1] : "EM");
		if (mod->accel_step && opt->accel_scheme >= NUM_ACCELERATION_METHODS)
			fprintf(stderr, "%d", opt->q);
		fprintf(stderr, "): %.2f (delta): %.5g\n", loglik, loglik - mod->logL);
	}
	mod->accel_step = 0;
	mod->logL = loglik;
	return mod->stopped;
}

/* reference em_alg.c:195-207: E-step on slot findex, M-step into slot tindex;
 * the log likelihood is that of the parameters the E-step read */
int em_step(options *opt, data *dat, model *mod)
{
	double ll = 0;

	(void)dat;
	if (mod->n_gpus == 1) {
		GPU(mc_em_step(mod->gpu, mod->findex, mod->tindex, &ll));
	} else {
		/* local E-step on every device, one NCCL all-gather + rank-order
		 * sum of the sufficient statistics, identical finish everywhere */
		GPUS(r, mc_em_step_local(mod->gpus[r], mod->findex, mod->tindex));
		comm_exchange(mod);
		GPUS(r, mc_em_step_finish(mod->gpus[r], mod->tindex, NULL));
		GPU(mc_read_ll(mod->gpu, &ll));
	}
	return stop(opt, mod, ll);
}

/* reference log_likelihood.c:56-62 */
double log_likelihood(options *opt, data *dat, model *mod, int which)
{
	double ll = 0;

	(void)opt;
	(void)dat;
	if (mod->n_gpus == 1) {
		GPU(mc_loglik(mod->gpu, which, &ll));
		return ll;
	}
	GPUS(r, mc_loglik(mod->gpus[r], which, NULL));
	for (int r = 0; r < mod->n_gpus; r++) {	/* rank order: deterministic */
		double part = 0;
		GPU(mc_read_ll(mod->gpus[r], &part));
		ll += part;
	}
	return ll;
}

/* reference log_likelihood.c:70-85 */
double aic(model *mod)
{
	return -2 * mod->max_logL + 2 * mod->no_parameters;
}

double bic(data *dat, model *mod)
{
	return -2 * mod->max_logL + mod->no_parameters * log(dat->I);
}

/* reference em_alg.c:1072-1171: two EM steps from the previous iterate,
 * recording u = F(x) - x and v = F(F(x)) - F(x) in secant pair delta_index */
int em_2_steps(model *mod, data *dat, options *opt)
{
	mod->findex = mod->pindex;
	mod->tindex = (mod->findex + 1) % 3;
	for (int j = 0; j < 2; j++) {
		if (em_step(opt, dat, mod))
			return 1;
		GPUS(r, mc_delta(mod->gpus[r], j, mod->delta_index, mod->tindex, mod->findex));
		mod->findex = mod->tindex;
		mod->tindex = (mod->findex + 1) % 3;
		if (mod->tindex == mod->pindex)
			mod->tindex = (mod->tindex + 1) % 3;
	}
	mod->delta_index = (mod->delta_index + 1) % opt->q;
	return 0;
}

/* reference accel_em.c:130-243; the three sums come back from the device in
 * two parts (eta, p) and are added in the reference's order (eta first) */
static double step_size(options *opt, model *mod)
{
	double e[3] = { 0, 0, 0 }, p[3], utu, utvu, vutvu, s;

	/* p is replicated: device 0's sums; eta rows are sharded: add the parts
	 * of all devices in rank order (a pooled eta is replicated too) */
	for (int r = 0; r < mod->n_gpus; r++) {
		double er[3], pr[3];
		GPU(mc_step_dots(mod->gpus[r], mod->delta_index, er, pr));
		for (int x = 0; x < 3; x++) {
			if (r == 0)
				p[x] = pr[x];
			if (r == 0 || eta_is_sharded(opt))
				e[x] += er[x];
		}
	}
	utu = e[0] + p[0];
	utvu = e[1] + p[1];
	vutvu = e[2] + p[2];
	switch (opt->accel_scheme) {
	case SQS1:
		s = utu / utvu;
		break;
	case SQS2:
		s = utvu / vutvu;
		break;
	case SQS3:
		if (sqrt(utu) < 1e-8)
			return NAN;
		s = -sqrt(utu / vutvu);
		break;
	case QN:
		s = -utu / utvu;
		break;
	default:
		s = -1;
	}
	if (opt->accel_scheme < QN && s > -1)
		s = -1;
	return s;
}

/* reference accel_em.c:422-551: SQUAREM / QN1 extrapolation into slot tindex,
 * projections, log likelihood of the result */
static double accelerated_update(options *opt, data *dat, model *mod, double s)
{
	double ll;

	mod->delta_index = mod->delta_index ? mod->delta_index - 1 : opt->q - 1;
	GPUS(r, mc_accel_update(mod->gpus[r], opt->accel_scheme == QN, mod->tindex,
		mod->pindex, mod->delta_index, s));
	ll = log_likelihood(opt, dat, mod, mod->tindex);
	mod->delta_index = (mod->delta_index + 1) % opt->q;
	return ll;
}

/* reference accel_em.c:262-419: quasi-Newton with q = 2, 3 secant pairs */
static double qn_accelerated_update(options *opt, data *dat, model *mod)
{
	const int q = opt->q;
	const int vindex = mod->delta_index ? mod->delta_index - 1 : q - 1;
	const int uindex = vindex ? vindex - 1 : q - 1;
	double *A = mod->A, *Ai = mod->Ainv, e[2], p[2] = { 0, 0 }, det;
	int q1 = mod->delta_index, q2, j = 0, n;

	do {
		q2 = mod->delta_index;
		n = 0;
		do {
			e[0] = e[1] = 0;
			for (int r = 0; r < mod->n_gpus; r++) {
				double er[2], pr[2];
				GPU(mc_qn_dots(mod->gpus[r], q1, q2, er, pr));
				if (r == 0) {
					p[0] = pr[0];
					p[1] = pr[1];
				}
				if (r == 0 || eta_is_sharded(opt)) {
					e[0] += er[0];
					e[1] += er[1];
				}
			}
			mod->cutu[n] = e[0] + p[0];
			A[j * q + n] = (e[0] + p[0]) - (e[1] + p[1]);
			n++;
			q2 = (q2 + 1) % q;
		} while (q2 != mod->delta_index);
		q1 = (q1 + 1) % q;
		j++;
	} while (q1 != mod->delta_index);

	if (q == 1) {
		Ai[0] = 1 / A[0];
	} else if (q == 2) {
		det = A[0] * A[3] - A[1] * A[2];
		Ai[0] = A[3] / det;
		Ai[3] = A[0] / det;
		Ai[1] = -A[1] / det;
		Ai[2] = -A[2] / det;
	} else {
		/* adjugate over determinant */
		const double c00 = A[4] * A[8] - A[5] * A[7];
		const double c01 = A[8] * A[3] - A[5] * A[6];
		const double c02 = A[3] * A[7] - A[4] * A[6];
		det = A[0] * c00 - A[1] * c01 + A[2] * c02;
		Ai[0] = c00 / det;
		Ai[1] = (A[2] * A[7] - A[1] * A[8]) / det;
		Ai[2] = (A[1] * A[5] - A[2] * A[4]) / det;
		Ai[3] = (A[5] * A[6] - A[3] * A[8]) / det;
		Ai[4] = (A[0] * A[8] - A[2] * A[6]) / det;
		Ai[5] = (A[2] * A[3] - A[0] * A[5]) / det;
		Ai[6] = c02 / det;
		Ai[7] = (A[1] * A[6] - A[0] * A[7]) / det;
		Ai[8] = (A[0] * A[4] - A[1] * A[3]) / det;
	}
	GPUS(r, mc_qn_update(mod->gpus[r], mod->tindex, mod->pindex, uindex,
		mod->delta_index, Ai, mod->cutu));
	return log_likelihood(opt, dat, mod, mod->tindex);
}

/* reference accel_em.c:35-114 */
int accelerated_em_step(options *opt, data *dat, model *mod)
{
	double emll, ll = 0, s = 0;
	int n_adjust = 0;

	em_2_steps(mod, dat, opt);
	if (mod->stopped)
		return 1;
	emll = log_likelihood(opt, dat, mod, mod->findex);
	if (opt->accel_scheme <= QN) {
		s = step_size(opt, mod);
		if (isnan(s) || isinf(s))
			goto keep_em;
	}
	do {
		if (opt->accel_scheme <= QN)
			ll = accelerated_update(opt, dat, mod, s);
		else
			ll = qn_accelerated_update(opt, dat, mod);
		if (opt->adjust_step && ll < emll)
			s = (s - 1) / 2;
	} while (n_adjust++ < opt->adjust_step && ll < emll && s < -1);

	if (ll > emll) {
		mod->pindex = mod->tindex;
		mod->accel_step = 1;
		return 0;
	}
keep_em:
	mod->pindex = mod->findex;
	return 0;
}

/* reference em_alg.c:44-90 */
void em(options *opt, data *dat, model *mod)
{
	int halt = 0;

	if (mod->K == 1) {
		em_step(opt, dat, mod);
		mod->logL = log_likelihood(opt, dat, mod, mod->tindex);
		return;
	}
	while (mod->n_iter < opt->n_init_iter && !halt)
		halt = em_step(opt, dat, mod);
	for (int i = 1; i < opt->q; i++) {
		em_2_steps(mod, dat, opt);
		mod->pindex = mod->findex;
	}
	if (mod->converged)
		return;
	do {
		halt = opt->accel_scheme ? accelerated_em_step(opt, dat, mod)
			: em_step(opt, dat, mod);
	} while (!halt);
}
