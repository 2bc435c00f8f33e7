/*
 * init_model.c -- random initialisation of the parameters (reference
 * rnd_init.c:54-110, 192-357, 456-482).
 *
 * The reference draws from glibc rand() -- one shared stream over all
 * initialisations and all K (multiclust.c:516-531, never reseeded unless -r) --
 * and parity needs the same stream, so the draws are made here on the host in
 * the reference's order, from the bit-identical generator of mc_rand.h.  What is done with them runs on the device:
 *   admixture: one cluster per allele copy, drawn on the device from the
 *              stream positions the host computes (mc_init_admixture_rand),
 *              -> hard-assignment counts and the M-step;
 *   mixture:   K random centre individuals, nearest-centre partition, smoothed
 *              counts; the O(I*L) counting is done here on the 8-bit codes and
 *              the resulting parameters are uploaded (mc_set_params).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "multiclust.h"

#define GPU(call) gpu_check(mod, (call), #call)

/* draws per device thread: a multiple of the kernel's 496-draw round */
#define RAND_BLOCK (496LL * 64)

/* rnd_init.c:456-482: k = rand() % K for every copy, missing ones included.
 * The draws are made on the device (mc_init_admixture_rand): the stream is a
 * linear recurrence, so the host only computes where every block of RAND_BLOCK
 * draws starts -- one 31 x 31 matrix-vector product per block -- and advances
 * its own generator by the same I*L*P draws.  With --gpus N device r draws for
 * its rows of individuals, i.e. from draw row_first[r]*L*P on. */
static uint32_t g_jump[MCR_LAG * MCR_LAG];	/* advance by RAND_BLOCK draws */
static pthread_once_t g_jump_once = PTHREAD_ONCE_INIT;

static void make_jump(void)
{
	mcr_jump_matrix(RAND_BLOCK, g_jump);
}

static int random_initialize_admixture(options *opt, data *dat, model *mod)
{
	const uint32_t *jump = g_jump;
	uint32_t h[MCR_LAG];
	const long long per_row = (long long)dat->L * dat->ploidy;

	(void)opt;
	pthread_once(&g_jump_once, make_jump);	/* fits may run on several host threads */
	mcr_history(mod->rng, h);
	for (int r = 0; r < mod->n_gpus; r++) {
		const long long n = (long long)(mod->row_first[r + 1] - mod->row_first[r]) * per_row;
		const long long nb = n ? (n + RAND_BLOCK - 1) / RAND_BLOCK : 1;
		uint32_t *hist = malloc(sizeof *hist * MCR_LAG * (size_t)nb);

		if (!hist)
			return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "generator states\n");
		for (long long b = 0; b < nb; b++) {
			memcpy(hist + b * MCR_LAG, h, sizeof h);
			if (b + 1 < nb)
				mcr_apply(jump, h);
			else
				mcr_step_history(h, n - b * RAND_BLOCK);
		}
		if (mod->n_gpus == 1)
			GPU(mc_init_admixture_rand(mod->gpus[r], mod->tindex, hist, nb, RAND_BLOCK));
		else
			GPU(mc_init_admixture_rand_local(mod->gpus[r], mod->tindex, hist, nb,
				RAND_BLOCK));
		free(hist);
	}
	mcr_from_history(mod->rng, h);
	if (mod->n_gpus > 1) {
		/* the allele counts are summed over devices before p is normalised */
		if (mc_comm_exchange(mod->comm) != MC_OK)
			return mmessage(ERROR_MSG, GPU_ERROR, "%s\n",
				mc_comm_last_error(mod->comm));
		for (int r = 0; r < mod->n_gpus; r++)
			GPU(mc_em_step_finish(mod->gpus[r], mod->tindex, NULL));
	}
	return NO_ERROR;
}

/* rnd_init.c:192-339.  The host draws the K distinct centre individuals from
 * the rand() stream; the nearest-centre assignment (I * K * L distance work),
 * the cluster counts and the starting parameters are made on the device
 * (mc_init_mixture). */
static int random_initialize_mixture(options *opt, data *dat, model *mod)
{
	const int K = mod->K, I = dat->I, L = dat->L, P = dat->ploidy;
	const size_t row = (size_t)L * P;
	int *center = malloc(sizeof *center * (size_t)K);
	int32_t *local = malloc(sizeof *local * (size_t)K);
	uint8_t *rows = malloc(row * (size_t)K > 0 ? row * (size_t)K : 1);

	(void)opt;
	if (!center || !local || !rows)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "initial parameters\n");
	if (K > 1) {
		/* K distinct random centres, re-drawing on collision (205-217) */
		for (int k = 0; k < K; k++) {
			int again;
			center[k] = mcr_next(mod->rng) % I;
			do {
				again = 0;
				for (int j = 0; j < k; j++)
					if (center[k] == center[j]) {
						center[k] = mcr_next(mod->rng) % I;
						again = 1;
						break;
					}
			} while (again);
			memcpy(rows + (size_t)k * row, dat->codes + (size_t)center[k] * row, row);
		}
	}
	for (int r = 0; r < mod->n_gpus; r++) {
		for (int k = 0; k < K; k++) {
			const int64_t i = K > 1 ? center[k] - mod->row_first[r] : -1;
			local[k] = (K > 1 && i >= 0 && center[k] < mod->row_first[r + 1]) ? (int32_t)i : -1;
		}
		if (mod->n_gpus == 1)
			GPU(mc_init_mixture(mod->gpus[r], mod->tindex, local, rows));
		else
			GPU(mc_init_mixture_local(mod->gpus[r], local, rows));
	}
	if (mod->n_gpus > 1) {
		/* counts and cluster sizes are summed over devices first */
		if (mc_comm_exchange(mod->comm) != MC_OK)
			return mmessage(ERROR_MSG, GPU_ERROR, "%s\n",
				mc_comm_last_error(mod->comm));
		for (int r = 0; r < mod->n_gpus; r++)
			GPU(mc_init_mixture_finish(mod->gpus[r], mod->tindex, I));
	}
	free(center);
	free(local);
	free(rows);
	return NO_ERROR;
}

/* bootstrap.c:31-52, 77-175: one parametric bootstrap sample under the saved H0
 * estimates.  As for the admixture initialiser the draws are made on the device
 * from the stream positions computed here (mc_bootstrap_data); the host
 * advances its generator by the 2 I L P (admixture) or I (1 + L P) (mixture)
 * draws of the sample (split by rows of individuals over the devices).  The mixture initialiser reads the centres' genotype
 * rows on the host (dat->codes): they are fetched back from the device. */
int parametric_bootstrap(options *opt, data *dat, model *mod)
{
	const long long per = (long long)dat->L * dat->ploidy;
	const long long per_i = opt->admixture ? 2 * per : 1 + per;
	uint32_t h[MCR_LAG];

	pthread_once(&g_jump_once, make_jump);
	mcr_history(mod->rng, h);
	/* with --gpus N device r draws the sample of its rows of individuals, i.e.
	 * from draw row_first[r] * per_i of the sample on */
	for (int r = 0; r < mod->n_gpus; r++) {
		const long long n = (long long)(mod->row_first[r + 1] - mod->row_first[r]) * per_i;
		const long long nb = n ? (n + RAND_BLOCK - 1) / RAND_BLOCK : 1;
		uint32_t *hist = malloc(sizeof *hist * MCR_LAG * (size_t)nb);

		if (!hist)
			return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "generator states\n");
		for (long long b = 0; b < nb; b++) {
			memcpy(hist + b * MCR_LAG, h, sizeof h);
			if (b + 1 < nb)
				mcr_apply(g_jump, h);
			else
				mcr_step_history(h, n - b * RAND_BLOCK);
		}
		GPU(mc_bootstrap_data(mod->gpus[r], hist, nb, RAND_BLOCK));
		free(hist);
	}
	mcr_from_history(mod->rng, h);
	if (!dat->codes_orig) {
		dat->codes_orig = dat->codes;
		dat->codes = malloc((size_t)dat->I * (size_t)per > 0 ? (size_t)dat->I * (size_t)per : 1);
		if (!dat->codes) {
			dat->codes = dat->codes_orig;
			dat->codes_orig = NULL;
			return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "bootstrap data\n");
		}
	}
	for (int r = 0; r < mod->n_gpus; r++)
		GPU(mc_get_codes(mod->gpus[r], dat->codes + (size_t)mod->row_first[r] * (size_t)per));
	return NO_ERROR;
}

/* bootstrap.c:60-66 */
int cleanup_parametric_bootstrap(data *dat, model *mod)
{
	if (dat->codes_orig) {
		free(dat->codes);
		dat->codes = dat->codes_orig;
		dat->codes_orig = NULL;
	}
	for (int r = 0; r < mod->n_gpus; r++)
		GPU(mc_restore_data(mod->gpus[r]));
	return NO_ERROR;
}

/* reference rnd_init.c:54-89 */
int initialize_model(options *opt, data *dat, model *mod)
{
	mod->n_iter = 0;
	mod->logL = -INFINITY;
	mod->converged = 0;
	if (opt->accel_scheme)
		mod->pindex = mod->tindex = mod->findex = 0;
	if (opt->admixture)
		return random_initialize_admixture(opt, dat, mod);
	return random_initialize_mixture(opt, dat, mod);
}
