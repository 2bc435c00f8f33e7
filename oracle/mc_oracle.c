/*
 * mc_oracle.c -- TEST INFRASTRUCTURE ONLY (see mc_oracle.h).
 *
 * CPU restatement of the reference's EM hot path on flat arrays.  Each
 * function cites the reference lines it follows.  Loop orders and expression
 * shapes are kept identical to the reference so that, compiled without FMA
 * contraction, results are bit-identical to the reference objects (checked by
 * tests/test_oracle_golden.py on vectors written by the reference itself).  The reference materialises d_iklm
 * (multiclust.c:1197); here the E-step folds it straight into the sums the
 * M-step needs, visiting terms in the reference's order.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mc_oracle.h"

#define MISSING_CODE 255

struct orc_fit {
	/* data */
	int I, L, P, T;
	int32_t *J, *off;
	uint8_t *codes;
	int32_t *cnt;		/* ILM, [I][T] */
	/* options */
	orc_options o;
	int q;
	double eta_lb, p_lb;
	/* model */
	int K, per_indiv, neta, n_par;
	double *p[3], *eta[3];
	double **up, **vp, **ue, **ve;	/* [q] secant pairs */
	double *A, *Ainv, *cutu;
	double *D;		/* admixture: sum_{l,m} d_iklm, [I][K] */
	double *N;		/* admixture: sum_i d_iklm, [K][T] */
	double *Dk;		/* -c: sum_{l,m,i} d_iklm, [K] */
	double *vik;		/* mixture posteriors, [I][K] */
	int pindex, findex, tindex, delta_index;
	double logL;
	int n_iter, converged, stopped, iter_stop, accel_step, aborted;
	/* every log likelihood passed to stop() */
	double *trace;
	int ntrace, ctrace;
};

/* ---------------------------------------------------------------- setup */

orc_fit *orc_create(int I, int L, int P, const int32_t *J,
	const uint8_t *codes, const orc_options *opt)
{
	orc_fit *f = calloc(1, sizeof *f);
	int i, l, a;
	double lb;

	f->I = I; f->L = L; f->P = P;
	f->J = malloc(sizeof(int32_t) * (size_t)L);
	f->off = malloc(sizeof(int32_t) * ((size_t)L + 1));
	f->off[0] = 0;
	for (l = 0; l < L; l++) {
		f->J[l] = J[l];
		f->off[l + 1] = f->off[l] + J[l];
	}
	f->T = f->off[L];
	f->codes = malloc((size_t)I * L * P);
	memcpy(f->codes, codes, (size_t)I * L * P);
	/* sufficient_statistics (read_file.c:651-657): count labelled alleles;
	 * missing copies are counted nowhere */
	f->cnt = calloc((size_t)I * (f->T ? f->T : 1), sizeof(int32_t));
	for (i = 0; i < I; i++)
		for (l = 0; l < L; l++)
			for (a = 0; a < P; a++) {
				int c = codes[((size_t)i * L + l) * P + a];
				if (c != MISSING_CODE)
					f->cnt[(size_t)i * f->T + f->off[l] + c]++;
			}
	f->o = *opt;
	/* synchronize (multiclust.c:812-820) */
	lb = 1.0 / I / P - 0.5 / I / P;
	if (opt->lower_bound < lb)
		lb = opt->lower_bound;
	f->eta_lb = f->p_lb = lb;
	f->q = 1;
	if (opt->accel_scheme >= 4) {
		f->o.adjust_step = 0;
		f->q = opt->accel_scheme - 3;
	}
	return f;
}

static void free_model(orc_fit *f)
{
	int s;
	for (s = 0; s < 3; s++) {
		free(f->p[s]); free(f->eta[s]);
		f->p[s] = f->eta[s] = NULL;
	}
	if (f->up)
		for (s = 0; s < f->q; s++) {
			free(f->up[s]); free(f->vp[s]);
			free(f->ue[s]); free(f->ve[s]);
		}
	free(f->up); free(f->vp); free(f->ue); free(f->ve);
	f->up = f->vp = f->ue = f->ve = NULL;
	free(f->A); free(f->Ainv); free(f->cutu);
	f->A = f->Ainv = f->cutu = NULL;
	free(f->D); free(f->N); free(f->Dk); free(f->vik);
	f->D = f->N = f->Dk = f->vik = NULL;
}

void orc_destroy(orc_fit *f)
{
	if (!f)
		return;
	free_model(f);
	free(f->J); free(f->off); free(f->codes); free(f->cnt); free(f->trace);
	free(f);
}

int orc_T(const orc_fit *f) { return f->T; }
double orc_lower_bound(const orc_fit *f) { return f->p_lb; }
int orc_q(const orc_fit *f) { return f->q; }
int orc_eta_len(const orc_fit *f) { return f->neta; }
int orc_n_parameters(const orc_fit *f) { return f->n_par; }

/* allocate_model_for_k, multiclust.c:1181-1279 */
int orc_alloc_model(orc_fit *f, int K)
{
	int s, l;
	size_t np;

	free_model(f);
	f->K = K;
	f->per_indiv = f->o.admixture && !f->o.eta_constrained;
	f->neta = f->per_indiv ? f->I * K : K;
	np = (size_t)K * (f->T ? f->T : 1);
	for (s = 0; s < 3; s++) {
		f->p[s] = calloc(np, sizeof(double));
		f->eta[s] = calloc((size_t)f->neta, sizeof(double));
	}
	if (f->o.accel_scheme) {
		f->up = calloc((size_t)f->q, sizeof *f->up);
		f->vp = calloc((size_t)f->q, sizeof *f->vp);
		f->ue = calloc((size_t)f->q, sizeof *f->ue);
		f->ve = calloc((size_t)f->q, sizeof *f->ve);
		for (s = 0; s < f->q; s++) {
			f->up[s] = calloc(np, sizeof(double));
			f->vp[s] = calloc(np, sizeof(double));
			f->ue[s] = calloc((size_t)f->neta, sizeof(double));
			f->ve[s] = calloc((size_t)f->neta, sizeof(double));
		}
		f->A = calloc((size_t)f->q * f->q, sizeof(double));
		f->Ainv = calloc((size_t)f->q * f->q, sizeof(double));
		f->cutu = calloc((size_t)f->q, sizeof(double));
	}
	if (f->o.admixture) {
		f->D = calloc((size_t)f->I * K, sizeof(double));
		f->N = calloc(np, sizeof(double));
		f->Dk = calloc((size_t)K, sizeof(double));
	} else {
		f->vik = calloc((size_t)f->I * K, sizeof(double));
	}
	/* parameter count, multiclust.c:1268-1276 (phantom slots included) */
	f->n_par = f->per_indiv ? f->I * (K - 1) : (K - 1);
	for (l = 0; l < f->L; l++)
		f->n_par += (f->J[l] - 1) * K;
	/* slot indices and delta_index are NOT reset here: the reference only
	 * resets p/f/t in initialize_model when accelerating
	 * (rnd_init.c:63-71) and never resets delta_index */
	return 0;
}

void orc_get_params(const orc_fit *f, int slot, double *eta, double *p)
{
	memcpy(eta, f->eta[slot], sizeof(double) * (size_t)f->neta);
	memcpy(p, f->p[slot], sizeof(double) * (size_t)f->K * f->T);
}

void orc_set_params(orc_fit *f, int slot, const double *eta, const double *p)
{
	memcpy(f->eta[slot], eta, sizeof(double) * (size_t)f->neta);
	memcpy(f->p[slot], p, sizeof(double) * (size_t)f->K * f->T);
}

void orc_set_indices(orc_fit *f, int pindex, int findex, int tindex)
{
	f->pindex = pindex; f->findex = findex; f->tindex = tindex;
}

void orc_get_state(const orc_fit *f, double *logL, int *n_iter, int *converged,
	int *stopped, int *iter_stop, int *pindex, int *aborted)
{
	*logL = f->logL; *n_iter = f->n_iter; *converged = f->converged;
	*stopped = f->stopped; *iter_stop = f->iter_stop; *pindex = f->pindex;
	*aborted = f->aborted;
}

void orc_get_posterior(const orc_fit *f, double *out)
{
	memcpy(out, f->o.admixture ? f->D : f->vik,
		sizeof(double) * (size_t)f->I * f->K);
}

int orc_trace_len(const orc_fit *f) { return f->ntrace; }
void orc_get_trace(const orc_fit *f, double *ll)
{
	memcpy(ll, f->trace, sizeof(double) * (size_t)f->ntrace);
}
void orc_reset_trace(orc_fit *f) { f->ntrace = 0; }

/* log_likelihood.c:70-85 */
double orc_aic(const orc_fit *f, double max_logL)
{
	return -2 * max_logL + 2 * f->n_par;
}
double orc_bic(const orc_fit *f, double max_logL)
{
	return -2 * max_logL + f->n_par * log((double)f->I);
}

/* ------------------------------------------------------ simplex.c:109-143 */

void orc_project(double *x, int len, double floor_)
{
	double csum, shift;
	int fixed[len > 0 ? len : 1];
	int i, n = len, done;

	for (i = 0; i < len; i++)
		fixed[i] = 0;
	while (n) {
		csum = 0.0;
		for (i = 0; i < len; i++)
			csum += x[i];
		shift = (csum - 1.0) / n;
		done = 1;
		for (i = 0; i < len; i++)
			if (!fixed[i]) {
				x[i] -= shift;
				if (x[i] < floor_) {
					x[i] = floor_;
					fixed[i] = 1;
					n--;
					done = 0;
				}
			}
		if (done)
			break;
	}
}

/* ------------------------------------------------- admixture E and M steps */

/* em_alg.c:325-433.  Also accumulates, in the order m_step_admixture_orig
 * (em_alg.c:604-752) would read d_iklm, the three sums that M-step forms. */
static double e_step_admixture(orc_fit *f)
{
	const int K = f->K, T = f->T;
	const double *eta = f->eta[f->findex], *p = f->p[f->findex];
	double t[K], tmp, ll = 0;
	int i, k, l, m;

	memset(f->D, 0, sizeof(double) * (size_t)f->I * K);
	memset(f->N, 0, sizeof(double) * (size_t)K * T);
	for (i = 0; i < f->I; i++) {
		const double *e = f->per_indiv ? eta + (size_t)i * K : eta;
		const int32_t *c = f->cnt + (size_t)i * T;
		for (l = 0; l < f->L; l++)
			for (m = f->off[l]; m < f->off[l + 1]; m++) {
				if (c[m] == 0)
					continue;	/* d = 0: adds nothing */
				tmp = 0;
				for (k = 0; k < K; k++) {
					t[k] = e[k] * p[(size_t)k * T + m];
					tmp += t[k];
				}
				for (k = 0; k < K; k++) {
					double d = c[m] * t[k] / tmp;
					f->D[(size_t)i * K + k] += d;	/* l,m order */
					f->N[(size_t)k * T + m] += d;	/* i order */
				}
				ll += c[m] * log(tmp);
			}
	}
	if (f->o.eta_constrained) {
		/* em_alg.c:604-629 sums d in (l, m, i) order for each k */
		for (k = 0; k < K; k++)
			f->Dk[k] = 0;
		for (l = 0; l < f->L; l++)
			for (m = f->off[l]; m < f->off[l + 1]; m++)
				for (i = 0; i < f->I; i++) {
					int c = f->cnt[(size_t)i * T + m];
					if (c == 0)
						continue;
					tmp = 0;
					for (k = 0; k < K; k++) {
						t[k] = eta[k] * p[(size_t)k * T + m];
						tmp += t[k];
					}
					for (k = 0; k < K; k++)
						f->Dk[k] += c * t[k] / tmp;
				}
	}
	return ll;
}

/* em_alg.c:592-754, working from the sums left by the E-step (or by the
 * hard-assignment initialiser) */
static void m_step_admixture(orc_fit *f)
{
	const int K = f->K, T = f->T;
	double *eta = f->eta[f->tindex], *p = f->p[f->tindex];
	double temp;
	int i, k, l, m;

	if (f->o.eta_constrained) {
		temp = 0.0;
		for (k = 0; k < K; k++) {
			eta[k] = f->Dk[k];
			temp += eta[k];
		}
		for (k = 0; k < K; k++)
			eta[k] /= temp;
		if (f->o.do_projection)
			orc_project(eta, K, f->eta_lb);
	} else {
		for (i = 0; i < f->I; i++) {
			double *e = eta + (size_t)i * K;
			temp = 0.0;
			for (k = 0; k < K; k++) {
				e[k] = f->D[(size_t)i * K + k];
				temp += e[k];
			}
			for (k = 0; k < K; k++)
				e[k] /= temp;
			if (f->o.do_projection)
				orc_project(e, K, f->eta_lb);
		}
	}
	for (k = 0; k < K; k++)
		for (l = 0; l < f->L; l++) {
			double *row = p + (size_t)k * T + f->off[l];
			const double *nrow = f->N + (size_t)k * T + f->off[l];
			temp = 0.0;
			for (m = 0; m < f->J[l]; m++) {
				row[m] = nrow[m];
				temp += row[m];
			}
			for (m = 0; m < f->J[l]; m++)
				row[m] /= temp;
			if (f->o.do_projection)
				orc_project(row, f->J[l], f->p_lb);
		}
}

/* --------------------------------------------------- mixture E and M steps */

/* em_alg.c:763-897 */
static double e_step_mixture(orc_fit *f)
{
	const int K = f->K, T = f->T;
	const double *eta = f->eta[f->findex], *p = f->p[f->findex];
	double log_eta[K], temp, max_ll, ll = 0;
	int i, k, m;

	for (k = 0; k < K; k++)
		log_eta[k] = log(eta[k]);
	for (i = 0; i < f->I; i++) {
		double *v = f->vik + (size_t)i * K;
		const int32_t *c = f->cnt + (size_t)i * T;
		max_ll = -INFINITY;
		for (k = 0; k < K; k++) {
			v[k] = log_eta[k];
			for (m = 0; m < T; m++) {	/* l then allele order */
				if (c[m] == 0 || p[(size_t)k * T + m] == 0.0)
					continue;
				v[k] += c[m] * log(p[(size_t)k * T + m]);
			}
			if (v[k] > max_ll)
				max_ll = v[k];
		}
		temp = 0;
		for (k = 0; k < K; k++) {
			v[k] = exp(v[k] - max_ll);
			temp += v[k];
		}
		for (k = 0; k < K; k++)
			v[k] /= temp;
		ll += log(temp) + max_ll;
	}
	return ll;
}

/* em_alg.c:907-1011 */
static void m_step_mixture(orc_fit *f)
{
	const int K = f->K, T = f->T;
	double *eta = f->eta[f->tindex], *p = f->p[f->tindex];
	double temp = 0.0;
	int i, k, l, m;

	for (k = 0; k < K; k++) {
		eta[k] = 0;
		for (i = 0; i < f->I; i++)
			eta[k] += f->vik[(size_t)i * K + k];
		temp += eta[k];
	}
	for (k = 0; k < K; k++)
		eta[k] /= temp;
	if (f->o.do_projection)
		orc_project(eta, K, f->eta_lb);

	for (k = 0; k < K; k++)
		for (l = 0; l < f->L; l++) {
			double *row = p + (size_t)k * T + f->off[l];
			temp = 0.0;
			for (m = 0; m < f->J[l]; m++) {
				row[m] = f->p_lb;	/* pseudo-count, em_alg.c:972 */
				for (i = 0; i < f->I; i++) {
					int c = f->cnt[(size_t)i * T + f->off[l] + m];
					if (c)
						row[m] += f->vik[(size_t)i * K + k] * c;
				}
				temp += row[m];
			}
			for (m = 0; m < f->J[l]; m++)
				row[m] /= temp;
			if (f->o.do_projection)
				orc_project(row, f->J[l], f->p_lb);
		}
}

/* ------------------------------------------------ log_likelihood.c:96-232 */

static double logL_admixture(orc_fit *f, int which)
{
	const int K = f->K, T = f->T;
	const double *eta = f->eta[which], *p = f->p[which];
	double temp, ll = 0.0;
	int i, k, m;

	for (i = 0; i < f->I; i++) {
		const double *e = f->per_indiv ? eta + (size_t)i * K : eta;
		const int32_t *c = f->cnt + (size_t)i * T;
		for (m = 0; m < T; m++) {
			if (c[m] == 0)
				continue;
			temp = 0.0;
			for (k = 0; k < K; k++)
				temp += e[k] * p[(size_t)k * T + m];
			ll += c[m] * log(temp);
		}
	}
	return ll;
}

static double logL_mixture(orc_fit *f, int which)
{
	const int K = f->K, T = f->T;
	const double *eta = f->eta[which], *p = f->p[which];
	double v[K], log_eta[K], max_exp, temp_exp, scale_exp, ll = 0.0;
	int i, k, m, out_of_range;

	for (k = 0; k < K; k++)
		log_eta[k] = log(eta[k]);
	for (i = 0; i < f->I; i++) {
		const int32_t *c = f->cnt + (size_t)i * T;
		max_exp = -INFINITY;
		for (k = 0; k < K; k++) {
			v[k] = 0.0;
			for (m = 0; m < T; m++) {
				if (c[m] == 0)
					continue;
				v[k] += c[m] * log(p[(size_t)k * T + m]);
			}
			v[k] += log_eta[k];
			if (v[k] > max_exp)
				max_exp = v[k];
		}
		temp_exp = exp(max_exp);
		scale_exp = 0.0;
		out_of_range = 0;
		if (temp_exp == 0.0 || temp_exp == HUGE_VAL) {
			out_of_range = 1;
			scale_exp = (temp_exp == HUGE_VAL) ? max_exp : -max_exp;
			do {
				scale_exp *= 0.5;
				temp_exp = exp(scale_exp);
			} while (temp_exp == HUGE_VAL);
			scale_exp = max_exp - scale_exp;
		}
		if (out_of_range)
			for (k = 0; k < K; k++)
				v[k] -= scale_exp;
		temp_exp = 0.0;
		for (k = 0; k < K; k++)
			temp_exp = temp_exp + exp(v[k]);
		ll = ll + log(temp_exp) + scale_exp;
	}
	return ll;
}

double orc_log_likelihood(orc_fit *f, int which)
{
	return f->o.admixture ? logL_admixture(f, which) : logL_mixture(f, which);
}

double orc_e_step(orc_fit *f)
{
	return f->o.admixture ? e_step_admixture(f) : e_step_mixture(f);
}

void orc_m_step(orc_fit *f)
{
	if (f->o.admixture)
		m_step_admixture(f);
	else
		m_step_mixture(f);
}

/* ---- sharded fits (test support): the sums a rank contributes ---- */

void orc_get_sums(const orc_fit *f, double *N, double *S)
{
	const int K = f->K, T = f->T;
	int i, k, m;

	if (f->o.admixture) {
		memcpy(N, f->N, sizeof(double) * (size_t)K * T);
		for (k = 0; k < K; k++) {
			S[k] = 0;
			for (i = 0; i < f->I; i++)
				S[k] += f->D[(size_t)i * K + k];
		}
		return;
	}
	for (k = 0; k < K; k++) {
		S[k] = 0;
		for (i = 0; i < f->I; i++)
			S[k] += f->vik[(size_t)i * K + k];
		for (m = 0; m < T; m++) {
			double s = 0;
			for (i = 0; i < f->I; i++) {
				int c = f->cnt[(size_t)i * T + m];
				if (c)
					s += f->vik[(size_t)i * K + k] * c;
			}
			N[(size_t)k * T + m] = s;
		}
	}
}

void orc_m_step_from_sums(orc_fit *f, const double *N, const double *S)
{
	const int K = f->K, T = f->T;
	double *eta = f->eta[f->tindex], *p = f->p[f->tindex];
	double temp;
	int k, l, m;

	if (f->o.admixture && !f->o.eta_constrained) {
		/* eta rows are local to the shard: the regular M-step forms them;
		 * only p needs the global sums */
		memcpy(f->N, N, sizeof(double) * (size_t)K * T);
		m_step_admixture(f);
		return;
	}
	if (f->o.admixture) {
		memcpy(f->N, N, sizeof(double) * (size_t)K * T);
		memcpy(f->Dk, S, sizeof(double) * (size_t)K);
		m_step_admixture(f);
		return;
	}
	temp = 0;
	for (k = 0; k < K; k++) {
		eta[k] = S[k];
		temp += eta[k];
	}
	for (k = 0; k < K; k++)
		eta[k] /= temp;
	if (f->o.do_projection)
		orc_project(eta, K, f->eta_lb);
	for (k = 0; k < K; k++)
		for (l = 0; l < f->L; l++) {
			double *row = p + (size_t)k * T + f->off[l];
			temp = 0.0;
			for (m = 0; m < f->J[l]; m++) {
				row[m] = f->p_lb + N[(size_t)k * T + f->off[l] + m];
				temp += row[m];
			}
			for (m = 0; m < f->J[l]; m++)
				row[m] /= temp;
			if (f->o.do_projection)
				orc_project(row, f->J[l], f->p_lb);
		}
}

/* ------------------------------------------------------ em_alg.c:101-207 */

static int converged(orc_fit *f, double ll)
{
	int stop = 1;
	double abs_diff = 0, rel_diff = 0;

	if (f->o.abs_error)
		abs_diff = fabs(ll - f->logL);
	if (f->o.rel_error)
		rel_diff = abs_diff / fabs(f->logL);
	if (f->o.abs_error && abs_diff > f->o.abs_error)
		stop &= 0;
	if (f->o.rel_error && rel_diff > f->o.rel_error)
		stop &= 0;
	if (stop)
		f->converged = 1;
	return stop;
}

static int stop_condition(orc_fit *f, double ll)
{
	if (f->o.max_iter && f->n_iter > f->o.max_iter) {
		f->iter_stop = 1;
		return 1;
	}
	return converged(f, ll);
}

static int stop(orc_fit *f, double ll)
{
	f->n_iter++;
	if (f->ntrace == f->ctrace) {
		f->ctrace = f->ctrace ? 2 * f->ctrace : 256;
		f->trace = realloc(f->trace, sizeof(double) * (size_t)f->ctrace);
	}
	f->trace[f->ntrace++] = ll;
	if (isnan(ll)) {		/* reference: exit(0) */
		f->aborted = 1;
		f->stopped = 1;
		return 1;
	}
	f->stopped = stop_condition(f, ll);
	if (ll < f->logL && !f->stopped) {	/* reference: exit(0) */
		f->aborted = 2;
		f->stopped = 1;
		return 1;
	}
	f->accel_step = 0;
	f->logL = ll;
	return f->stopped;
}

int orc_em_step(orc_fit *f)
{
	double ll = orc_e_step(f);
	orc_m_step(f);
	return stop(f, ll);
}

/* --------------------------------------------------- em_alg.c:1072-1211 */

int orc_em_2_steps(orc_fit *f)
{
	const size_t np = (size_t)f->K * f->T;
	size_t x;
	int j;

	f->findex = f->pindex;
	f->tindex = (f->findex + 1) % 3;
	for (j = 0; j < 2; j++) {
		double *dp = j ? f->vp[f->delta_index] : f->up[f->delta_index];
		double *de = j ? f->ve[f->delta_index] : f->ue[f->delta_index];
		if (orc_em_step(f))
			return 1;
		for (x = 0; x < np; x++)
			dp[x] = f->p[f->tindex][x] - f->p[f->findex][x];
		for (x = 0; x < (size_t)f->neta; x++)
			de[x] = f->eta[f->tindex][x] - f->eta[f->findex][x];
		f->findex = f->tindex;
		f->tindex = (f->findex + 1) % 3;
		if (f->tindex == f->pindex)
			f->tindex = (f->tindex + 1) % 3;
	}
	f->delta_index = (f->delta_index + 1) % f->q;
	return 0;
}

/* ------------------------------------------------------ accel_em.c:130-243 */

static double step_size(orc_fit *f)
{
	const double *ue = f->ue[f->delta_index], *ve = f->ve[f->delta_index];
	const double *up = f->up[f->delta_index], *vp = f->vp[f->delta_index];
	double utu = 0, utvu = 0, vutvu = 0, s;
	size_t x, np = (size_t)f->K * f->T;

	for (x = 0; x < (size_t)f->neta; x++) {
		utu += ue[x] * ue[x];
		utvu += ue[x] * (ve[x] - ue[x]);
		vutvu += (ve[x] - ue[x]) * (ve[x] - ue[x]);
	}
	for (x = 0; x < np; x++) {
		utu += up[x] * up[x];
		utvu += up[x] * (vp[x] - up[x]);
		vutvu += (vp[x] - up[x]) * (vp[x] - up[x]);
	}
	switch (f->o.accel_scheme) {
	case 1: s = utu / utvu; break;
	case 2: s = utvu / vutvu; break;
	case 3:
		if (sqrt(utu) < 1e-8)
			return NAN;
		s = -sqrt(utu / vutvu);
		break;
	case 4: s = -utu / utvu; break;
	default: s = -1;
	}
	if (f->o.accel_scheme < 4 && s > -1)
		s = -1;
	return s;
}

/* ------------------------------------------------------ accel_em.c:422-551 */

static double accelerated_update(orc_fit *f, double s)
{
	const int K = f->K, T = f->T;
	const int qn = f->o.accel_scheme == 4;
	double *pt = f->p[f->tindex], *et = f->eta[f->tindex];
	const double *pp = f->p[f->pindex], *ep = f->eta[f->pindex];
	const double *up, *vp, *ue, *ve;
	double ll;
	int i, k, l, m;

	f->delta_index = f->delta_index ? f->delta_index - 1 : f->q - 1;
	up = f->up[f->delta_index]; vp = f->vp[f->delta_index];
	ue = f->ue[f->delta_index]; ve = f->ve[f->delta_index];

	for (l = 0; l < f->L; l++)
		for (k = 0; k < K; k++) {
			size_t b = (size_t)k * T + f->off[l];
			for (m = 0; m < f->J[l]; m++) {
				if (qn)
					pt[b + m] = pp[b + m] + up[b + m]
						+ s * vp[b + m];
				else
					pt[b + m] = pp[b + m]
						- 2 * s * up[b + m]
						+ s * s * (vp[b + m] - up[b + m]);
			}
			if (f->o.do_projection)
				orc_project(pt + b, f->J[l], f->p_lb);
		}
	if (f->per_indiv) {
		for (i = 0; i < f->I; i++) {
			size_t b = (size_t)i * K;
			for (k = 0; k < K; k++)
				if (qn)
					et[b + k] = ep[b + k] + ue[b + k]
						+ s * ve[b + k];
				else
					et[b + k] = ep[b + k]
						- 2 * s * ue[b + k]
						+ s * s * (ve[b + k] - ue[b + k]);
			if (f->o.do_projection)
				orc_project(et + b, K, f->eta_lb);
		}
	} else {
		for (k = 0; k < K; k++)
			if (qn)
				et[k] = ep[k] + ue[k] + s * ve[k];
			else
				et[k] = ep[k] - 2 * s * ue[k]
					+ s * s * (ve[k] - ue[k]);
		if (f->o.do_projection)
			orc_project(et, K, f->eta_lb);
	}
	ll = orc_log_likelihood(f, f->tindex);
	f->delta_index = (f->delta_index + 1) % f->q;
	return ll;
}

/* ------------------------------------------------------ accel_em.c:262-419 */

static double qn_accelerated_update(orc_fit *f)
{
	const int K = f->K, T = f->T, q = f->q, I = f->I;
	int vindex = f->delta_index ? f->delta_index - 1 : q - 1;
	int uindex = vindex ? vindex - 1 : q - 1;
	double *pt = f->p[f->tindex], *et = f->eta[f->tindex];
	const double *pp = f->p[f->pindex], *ep = f->eta[f->pindex];
	double utu, utv, det, *A = f->A, *Ai = f->Ainv;
	int q1, q2, i, j, k, n, m;

	q1 = f->delta_index;
	j = 0;
	do {
		q2 = f->delta_index;
		n = 0;
		do {
			utu = 0;
			utv = 0;
			for (k = 0; k < K; k++) {
				if (f->per_indiv) {
					for (i = 0; i < I; i++) {
						size_t x = (size_t)i * K + k;
						utu += f->ue[q1][x] * f->ue[q2][x];
						utv += f->ue[q1][x] * f->ve[q2][x];
					}
				} else {
					utu += f->ue[q1][k] * f->ue[q2][k];
					utv += f->ue[q1][k] * f->ve[q2][k];
				}
				for (m = 0; m < T; m++) {
					size_t x = (size_t)k * T + m;
					utu += f->up[q1][x] * f->up[q2][x];
					utv += f->up[q1][x] * f->vp[q2][x];
				}
			}
			f->cutu[n] = utu;
			A[j * q + n] = utu - utv;
			n++;
			q2 = (q2 + 1) % q;
		} while (q2 != f->delta_index);
		q1 = (q1 + 1) % q;
		j++;
	} while (q1 != f->delta_index);

	if (q == 1) {
		Ai[0] = 1 / A[0];
	} else if (q == 2) {
		det = A[0] * A[3] - A[1] * A[2];
		Ai[0] = A[3] / det;
		Ai[3] = A[0] / det;
		Ai[1] = -A[1] / det;
		Ai[2] = -A[2] / det;
	} else if (q == 3) {
		det = A[0] * (A[4] * A[8] - A[5] * A[7])
			- A[1] * (A[8] * A[3] - A[5] * A[6])
			+ A[2] * (A[3] * A[7] - A[4] * A[6]);
		Ai[0] = (A[4] * A[8] - A[5] * A[7]) / det;
		Ai[1] = (A[2] * A[7] - A[1] * A[8]) / det;
		Ai[2] = (A[1] * A[5] - A[2] * A[4]) / det;
		Ai[3] = (A[5] * A[6] - A[3] * A[8]) / det;
		Ai[4] = (A[0] * A[8] - A[2] * A[6]) / det;
		Ai[5] = (A[2] * A[3] - A[0] * A[5]) / det;
		Ai[6] = (A[3] * A[7] - A[4] * A[6]) / det;
		Ai[7] = (A[1] * A[6] - A[0] * A[7]) / det;
		Ai[8] = (A[0] * A[4] - A[1] * A[3]) / det;
	}

	for (i = 0; i < f->neta; i++)
		et[i] = ep[i] + f->ue[uindex][i];
	for (size_t x = 0; x < (size_t)K * T; x++)
		pt[x] = pp[x] + f->up[uindex][x];
	q1 = f->delta_index;
	j = 0;
	do {
		n = 0;
		q2 = f->delta_index;
		do {
			for (i = 0; i < f->neta; i++)
				et[i] += f->ve[q1][i] * Ai[j * q + n] * f->cutu[n];
			for (size_t x = 0; x < (size_t)K * T; x++)
				pt[x] += f->vp[q1][x] * Ai[j * q + n] * f->cutu[n];
			q2 = (q2 + 1) % q;
			n++;
		} while (q2 != f->delta_index);
		q1 = (q1 + 1) % q;
		j++;
	} while (q1 != f->delta_index);

	if (f->o.do_projection) {
		if (f->per_indiv)
			for (i = 0; i < I; i++)
				orc_project(et + (size_t)i * K, K, f->eta_lb);
		else
			orc_project(et, K, f->eta_lb);
		for (k = 0; k < K; k++)
			for (int l = 0; l < f->L; l++)
				orc_project(pt + (size_t)k * T + f->off[l],
					f->J[l], f->p_lb);
	}
	return orc_log_likelihood(f, f->tindex);
}

/* ------------------------------------------------------- accel_em.c:35-114 */

int orc_accelerated_em_step(orc_fit *f)
{
	int n_adjust = 0;
	double emll, ll = 0, s = 0;

	orc_em_2_steps(f);
	if (f->stopped)
		return 1;
	emll = orc_log_likelihood(f, f->findex);
	if (f->o.accel_scheme <= 4) {
		s = step_size(f);
		if (isnan(s) || isinf(s))
			goto EM_EXIT;
	}
	do {
		if (f->o.accel_scheme <= 4)
			ll = accelerated_update(f, s);
		else
			ll = qn_accelerated_update(f);
		if (f->o.adjust_step && ll < emll)
			s = (s - 1) / 2;
	} while (n_adjust++ < f->o.adjust_step && ll < emll && s < -1);

	if (ll > emll) {
		f->pindex = f->tindex;
		f->accel_step = 1;
		return 0;
	}
EM_EXIT:
	f->pindex = f->findex;
	return 0;
}

/* ---------------------------------------------------------- em_alg.c:44-90 */

void orc_em(orc_fit *f)
{
	int i, stop_ = 0;

	if (f->K == 1) {
		orc_em_step(f);
		f->logL = orc_log_likelihood(f, f->tindex);
		return;
	}
	while (f->n_iter < f->o.n_init_iter && !stop_)
		stop_ = orc_em_step(f);
	for (i = 1; i < f->q; i++) {
		orc_em_2_steps(f);
		f->pindex = f->findex;
	}
	if (f->converged || f->aborted)
		return;
	do {
		if (!f->o.accel_scheme)
			stop_ = orc_em_step(f);
		else
			stop_ = orc_accelerated_em_step(f);
	} while (!stop_);
}

/* ------------------------------------------------- rnd_init.c: initialisers */

void orc_seed(long seed)
{
	if (seed >= 0)
		srand((unsigned int)seed);
}

/* random_allele_partition + m_step_admixture, rnd_init.c:349-357,456-482.
 * d_iklm is SET to 1 (not incremented), so two copies of one allele drawn to
 * the same k count once. */
static void init_admixture(orc_fit *f)
{
	const int K = f->K, T = f->T, P = f->P;
	unsigned char *hit = calloc((size_t)K * (T ? T : 1), 1);
	size_t *touched = malloc(sizeof(size_t) * (size_t)f->L * P + 1);
	size_t nt, x;
	int i, k, l, a;

	memset(f->D, 0, sizeof(double) * (size_t)f->I * K);
	memset(f->N, 0, sizeof(double) * (size_t)K * T);
	for (k = 0; k < K; k++)
		f->Dk[k] = 0;
	for (i = 0; i < f->I; i++) {
		nt = 0;
		for (l = 0; l < f->L; l++)
			for (a = 0; a < P; a++) {
				int c = f->codes[((size_t)i * f->L + l) * P + a];
				k = (int)rand() % K;	/* drawn for missing too */
				if (c == MISSING_CODE)
					continue;
				x = (size_t)k * T + f->off[l] + c;
				if (!hit[x]) {
					hit[x] = 1;
					touched[nt++] = x;
				}
			}
		for (x = 0; x < nt; x++) {
			k = (int)(touched[x] / (size_t)T);
			f->D[(size_t)i * K + k] += 1;
			f->N[touched[x]] += 1;
			f->Dk[k] += 1;
			hit[touched[x]] = 0;
		}
	}
	free(hit);
	free(touched);
	m_step_admixture(f);
}

/* random_individual_center + initialize_parameters_mixture,
 * rnd_init.c:192-339 */
static void init_mixture(orc_fit *f)
{
	const int K = f->K, T = f->T, I = f->I;
	int *I_K = malloc(sizeof(int) * (size_t)I);
	int center[K];
	double *eta = f->eta[f->tindex], *p = f->p[f->tindex];
	double count_diff, min_count_diff, temp;
	int i, j, k, l, m, flag;

	if (K == 1) {
		for (i = 0; i < I; i++)
			I_K[i] = 0;
	} else {
		for (k = 0; k < K; k++) {
			center[k] = (int)(rand() % I);
			do {
				flag = 0;
				for (j = 0; j < k; j++)
					if (center[k] == center[j]) {
						center[k] = (int)(rand() % I);
						flag = 1;
						break;
					}
			} while (flag == 1);
		}
		for (i = 0; i < I; i++) {
			I_K[i] = 0;
			if (i == center[0])
				continue;
			min_count_diff = INFINITY;
			for (k = 0; k < K; k++) {
				if (i == center[k]) {
					I_K[i] = k;
					break;
				}
				count_diff = 0;
				for (m = 0; m < T; m++)
					count_diff += abs(f->cnt[(size_t)i * T + m]
						- f->cnt[(size_t)center[k] * T + m]);
				if (count_diff < min_count_diff) {
					I_K[i] = k;
					min_count_diff = count_diff;
				}
			}
		}
	}

	for (k = 0; k < K; k++)
		eta[k] = 1;
	for (i = 0; i < I; i++)
		eta[I_K[i]]++;
	for (k = 0; k < K; k++)
		eta[k] /= I + K;

	/* the accumulation sits inside the k loop (rnd_init.c:296-318): row k
	 * is reset at iteration k and then keeps receiving the counts of its
	 * members on every later pass */
	for (k = 0; k < K; k++)
		for (l = 0; l < f->L; l++)
			for (m = f->off[l]; m < f->off[l + 1]; m++) {
				p[(size_t)k * T + m] = 1.0;
				for (i = 0; i < I; i++) {
					int c = f->cnt[(size_t)i * T + m];
					if (c == 0)
						continue;
					p[(size_t)I_K[i] * T + m] += c;
				}
			}
	for (k = 0; k < K; k++)
		for (l = 0; l < f->L; l++) {
			temp = 0.0;
			for (m = f->off[l]; m < f->off[l + 1]; m++)
				temp += p[(size_t)k * T + m];
			for (m = f->off[l]; m < f->off[l + 1]; m++)
				p[(size_t)k * T + m] /= temp;
		}
	free(I_K);
}

/* initialize_model, rnd_init.c:54-89 */
void orc_initialize(orc_fit *f)
{
	f->n_iter = 0;
	f->logL = -INFINITY;
	f->converged = 0;
	f->stopped = 0;
	f->iter_stop = 0;
	f->aborted = 0;
	f->accel_step = 0;
	if (f->o.accel_scheme)
		f->pindex = f->tindex = f->findex = 0;
	if (f->o.admixture)
		init_admixture(f);
	else
		init_mixture(f);
}
