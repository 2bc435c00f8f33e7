/*
 * mc_oracle.h -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the
 * reference's EM hot path on the flat layout used by the CUDA implementation.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libmc_cuda.so and
 * the multiclust host binary) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py on vectors written by the reference itself checks this restatement
 * bit-for-bit (log likelihood, parameters, posterior sums) against the
 * unmodified reference objects driven by oracle/ref_harness.c, and against the
 * golden vectors that harness wrote into tests/golden/.
 *
 * Flat layout: off[l] = sum_{l'<l} J_l', T = off[L]; p[k*T + off[l] + j];
 * eta[i*K + k] (admixture) or eta[k] (mixture / -c); codes[(i*L + l)*P + a]
 * in 0..nreal_l-1 or 255 (missing); cnt[i*T + off[l] + j] = ILM.
 */
#ifndef MC_ORACLE_H
#define MC_ORACLE_H

#include <stdint.h>

typedef struct {
	int admixture;		/* -a */
	int eta_constrained;	/* -c */
	int accel_scheme;	/* -s 0..6 */
	int do_projection;	/* 0 after --projection */
	int n_init_iter;	/* -i */
	int max_iter;		/* -T / -C */
	int adjust_step;	/* -g */
	double abs_error;	/* -E */
	double rel_error;	/* -e */
	double lower_bound;	/* --bound */
} orc_options;

typedef struct orc_fit orc_fit;

/* data + options; J[l] includes the phantom slot of loci with missing data */
orc_fit *orc_create(int I, int L, int P, const int32_t *J,
	const uint8_t *codes, const orc_options *opt);
void orc_destroy(orc_fit *f);

int orc_T(const orc_fit *f);
double orc_lower_bound(const orc_fit *f);	/* after synchronize() */
int orc_q(const orc_fit *f);

/* allocate_model_for_k (multiclust.c:1181-1279) */
int orc_alloc_model(orc_fit *f, int K);
int orc_n_parameters(const orc_fit *f);

/* initialize_model (rnd_init.c:54-89); consumes libc rand() exactly as the
 * reference does.  seed < 0: do not call srand (reference default). */
void orc_seed(long seed);
void orc_initialize(orc_fit *f);

/* parameter slots 0..2 */
void orc_get_params(const orc_fit *f, int slot, double *eta, double *p);
void orc_set_params(orc_fit *f, int slot, const double *eta, const double *p);
int orc_eta_len(const orc_fit *f);

/* pieces of the path */
double orc_e_step(orc_fit *f);			/* reads findex, fills d / vik */
void orc_m_step(orc_fit *f);			/* writes tindex */
int orc_em_step(orc_fit *f);			/* em_step: E, M, stop() */
double orc_log_likelihood(orc_fit *f, int slot);
int orc_em_2_steps(orc_fit *f);
int orc_accelerated_em_step(orc_fit *f);
void orc_project(double *x, int n, double floor_);	/* michelot_project */
void orc_em(orc_fit *f);			/* em() */

/* sufficient statistics left by orc_e_step, for tests of individual-sharded
 * fits: [K*T allele-count sums | K pooled-eta sums] (admixture: N and Dk;
 * mixture: sum_i v_ik c_ilj without the pseudo-count and sum_i v_ik) */
void orc_get_sums(const orc_fit *f, double *N, double *S);
/* M-step from externally summed statistics (the exchange step of a sharded
 * fit): overwrites the allele-count / pooled sums, keeps the local D_ik */
void orc_m_step_from_sums(orc_fit *f, const double *N, const double *S);

/* state */
void orc_set_indices(orc_fit *f, int pindex, int findex, int tindex);
void orc_get_state(const orc_fit *f, double *logL, int *n_iter, int *converged,
	int *stopped, int *iter_stop, int *pindex, int *aborted);
void orc_get_posterior(const orc_fit *f, double *out);	/* D_ik or v_ik, [I][K] */
int orc_trace_len(const orc_fit *f);
void orc_get_trace(const orc_fit *f, double *ll);	/* every ll given to stop() */
void orc_reset_trace(orc_fit *f);
double orc_aic(const orc_fit *f, double max_logL);
double orc_bic(const orc_fit *f, double max_logL);

#endif
