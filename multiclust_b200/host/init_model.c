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

/* L1 distance between the allele-count vectors of two individuals at all loci
 * (rnd_init.c:238-247), computed from the codes: at one locus it is
 * sum_j |n_a(j) - n_b(j)| over the alleles either of them carries */
static double count_distance(const data *dat, int a, int b)
{
	const int P = dat->ploidy, L = dat->L;
	const uint8_t *ca = dat->codes + (size_t)a * L * P;
	const uint8_t *cb = dat->codes + (size_t)b * L * P;
	double d = 0;

	for (int l = 0; l < L; l++, ca += P, cb += P) {
		for (int x = 0; x < P; x++) {
			int first = 1, na = 0, nb = 0;
			if (ca[x] == MC_CODE_MISSING)
				continue;
			for (int y = 0; y < x; y++)
				if (ca[y] == ca[x])
					first = 0;
			if (!first)
				continue;
			for (int y = 0; y < P; y++) {
				na += ca[y] == ca[x];
				nb += cb[y] == ca[x];
			}
			d += abs(na - nb);
		}
		for (int x = 0; x < P; x++) {	/* alleles only b carries */
			int first = 1, nb = 0, in_a = 0;
			if (cb[x] == MC_CODE_MISSING)
				continue;
			for (int y = 0; y < x; y++)
				if (cb[y] == cb[x])
					first = 0;
			for (int y = 0; y < P; y++) {
				in_a |= ca[y] == cb[x];
				nb += cb[y] == cb[x];
			}
			if (first && !in_a)
				d += nb;
		}
	}
	return d;
}

/* rnd_init.c:192-339 */
static int random_initialize_mixture(options *opt, data *dat, model *mod)
{
	const int K = mod->K, I = dat->I, L = dat->L, P = dat->ploidy;
	const int64_t T = mod->T;
	int *center = malloc(sizeof *center * (size_t)K);
	double *eta = calloc((size_t)K, sizeof *eta);
	double *p = calloc((size_t)K * (T ? T : 1), sizeof *p);
	int *part = malloc(sizeof *part * (size_t)I);

	(void)opt;
	if (!center || !eta || !p || !part)
		return mmessage(ERROR_MSG, MEMORY_ALLOCATION, "initial parameters\n");

	if (K == 1) {
		for (int i = 0; i < I; i++)
			part[i] = 0;
	} else {
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
		}
		/* nearest centre, strictly smaller distance wins (220-258) */
		for (int i = 0; i < I; i++) {
			double best = INFINITY;
			part[i] = 0;
			if (i == center[0])
				continue;
			for (int k = 0; k < K; k++) {
				double d;
				if (i == center[k]) {
					part[i] = k;
					break;
				}
				d = count_distance(dat, i, center[k]);
				if (d < best) {
					part[i] = k;
					best = d;
				}
			}
		}
	}

	/* eta_k = (1 + n_k) / (I + K)  (274-293) */
	for (int k = 0; k < K; k++)
		eta[k] = 1;
	for (int i = 0; i < I; i++)
		eta[part[i]]++;
	for (int k = 0; k < K; k++)
		eta[k] /= I + K;

	/* p_klj proportional to 1 + (K - k) S_klj with S_klj the allele count of
	 * cluster k: the reference's accumulation sits inside its k loop
	 * (296-318), so row k is reset at pass k and then receives its members'
	 * counts on passes k..K-1.  Counts are integers: the closed form is
	 * bit-identical to the repeated additions. */
	for (int i = 0; i < I; i++) {
		const uint8_t *c = dat->codes + (size_t)i * L * P;
		double *row = p + (size_t)part[i] * T;
		for (int l = 0; l < L; l++)
			for (int a = 0; a < P; a++)
				if (c[l * P + a] != MC_CODE_MISSING)
					row[dat->allele_off[l] + c[l * P + a]] += 1;
	}
	for (int k = 0; k < K; k++)
		for (int l = 0; l < L; l++) {
			double *row = p + (size_t)k * T + dat->allele_off[l];
			double sum = 0.0;
			for (int m = 0; m < dat->uniquealleles[l]; m++) {
				row[m] = 1.0 + (K - k) * row[m];
				sum += row[m];
			}
			for (int m = 0; m < dat->uniquealleles[l]; m++)
				row[m] /= sum;
		}
	for (int r = 0; r < mod->n_gpus; r++)	/* eta_k and p are replicated */
		GPU(mc_set_params(mod->gpus[r], mod->tindex, eta, p));
	free(center);
	free(eta);
	free(p);
	free(part);
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
