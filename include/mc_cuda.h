/*
 * mc_cuda.h -- C ABI of the B200 (sm_100a) EM hot path of MULTICLUST.
 *
 * This is the drop-in boundary of SURVEY.md section 8(b): a thin layer around
 * the reference's L2 "EM hot path" (the functions declared at
 * multiclust.h:371-388).  The host program keeps every policy decision the
 * reference makes in C -- slot rotation, accept/reject of accelerated steps,
 * stop()/converged(), the exit(0) rules (em_alg.c:44-207, accel_em.c:35-114)
 * -- and calls down here only for the data-parallel work.  Plain pointers and
 * sizes, no C++ or torch types.  One host thread per context; every device
 * buffer is owned by the context.  One piece of process-wide state: device
 * memory comes from the device's default stream-ordered pool (cudaMallocAsync),
 * and mc_create lifts that pool's release threshold so that what one context
 * frees is handed out again to the next (a K sweep re-plans per K, and a fresh
 * cudaMalloc of the multi-GB layouts costs ~1 s); a host that shares the
 * process with other pool users can lower it again with cudaMemPoolSetAttribute.
 *
 * Flat layouts (all parameters IEEE double, as in the reference):
 *   J[l]    = allele slots of locus l, INCLUDING the phantom slot the
 *             reference creates for loci with missing data
 *             (read_file.c:527-530); off[l] = sum_{l'<l} J[l'], T = off[L]
 *   codes   = uint8 [I][L][P], 0..J[l]-1, 255 = missing copy      (dat->IL)
 *   p       = double [K][T], p[k*T + off[l] + j]                  (mod->vpklm[s])
 *   eta     = double [I][K] (admixture) or [K] (mixture, or -c)   (mod->vetaik[s] / vetak[s])
 *   slots   = 0..2, the reference's three rotating parameter copies
 *             (multiclust.h:265-271)
 *
 * Every function returns MC_OK or an MC_ERR_* code; mc_last_error() gives the
 * message.  There is no CPU fallback: without a CUDA device mc_create fails.
 */
#ifndef MC_CUDA_H
#define MC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#include "mc_synth.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mc_ctx mc_ctx;

enum {
	MC_OK = 0,
	MC_ERR_CUDA = 1,	/* a CUDA runtime call failed */
	MC_ERR_ARG = 2,		/* invalid argument */
	MC_ERR_STATE = 3,	/* call made in the wrong order (no data / no model) */
	MC_ERR_NOMEM = 4,
	MC_ERR_UNSUPPORTED = 5	/* e.g. ploidy > 16 or > 254 alleles at a locus */
};

#define MC_ABI_VERSION 3

/* message of the last error raised through `ctx` (or creation, if ctx NULL) */
const char *mc_last_error(const mc_ctx *ctx);
int mc_abi_version(void);

/* ---- context ---------------------------------------------------------- */

/* One context per GPU.  `device` is a CUDA ordinal. */
int mc_create(mc_ctx **ctx, int device);
void mc_destroy(mc_ctx *ctx);
/* Run every later launch on an existing cudaStream_t (e.g. the caller's
 * torch stream); NULL restores the context's own stream. */
int mc_set_stream(mc_ctx *ctx, void *cuda_stream);
int mc_sync(mc_ctx *ctx);
/* the CUDA ordinal and the cudaStream_t launches currently go to */
int mc_ctx_device(const mc_ctx *ctx);
void *mc_ctx_stream(const mc_ctx *ctx);

/* Plan options, to be set before mc_alloc_model (they replace environment
 * variables: a stray variable must not switch kernels).
 *   MC_OPT_KERNEL  which genotype-streaming kernel the planner may pick:
 *                  MC_KERNEL_AUTO    mixture model, K <= 16, ploidy <= 15: the
 *                                    digit-sliced integer kernels (mc_digit.cuh;
 *                                    the count layouts must fit the device).
 *                                    Admixture model with no locus of more than
 *                                    two observed alleles and K <= 16: the dense
 *                                    DMMA kernels (mc_dense.cuh).  Else the
 *                                    two-pass gather kernel (mc_admix3.cuh;
 *                                    K <= 16, ploidy <= 8), else the one-pass
 *                                    tile kernel
 *                  MC_KERNEL_TILE    the one-pass tile kernel only
 *                  MC_KERNEL_ADMIX3  never the dense or the digit-sliced kernels
 *                  MC_KERNEL_DENSE   the dense DMMA kernels where they apply (for
 *                                    the mixture model too), else the one-pass
 *                                    tile kernel
 *                  MC_KERNEL_DIGIT   same plans as AUTO (named for tests)
 *                  The kernels sum in different orders, so results agree to
 *                  rounding (1e-13 relative observed), not bit for bit.
 *   MC_OPT_TIMING  non-zero: the planner prints its phases to stderr
 *   MC_OPT_GRAPH   (default 1) mc_em_step and mc_loglik replay their launch
 *                  sequence as a CUDA graph from the third call of a slot
 *                  (pair) on: a small fit is launch-bound.  Same kernels, same
 *                  results.  Off while mc_profile_enable is on. */
enum { MC_OPT_KERNEL = 1, MC_OPT_TIMING = 2, MC_OPT_GRAPH = 3 };
enum { MC_KERNEL_AUTO = 0, MC_KERNEL_TILE = 1, MC_KERNEL_ADMIX3 = 2, MC_KERNEL_DENSE = 3,
	MC_KERNEL_DIGIT = 4 };
int mc_set_option(mc_ctx *ctx, int option, int value);

/* ---- data: replaces dat->IL / ILM / uniquealleles as the path reads them
 *      (multiclust.h:223-250, read_file.c:633-663) ------------------------ */

/* Upload recoded genotypes from HOST memory and build the device layout. */
int mc_set_data(mc_ctx *ctx, int64_t I, int32_t L, int32_t P,
	const int32_t *J, const uint8_t *codes);
/* Generate the synthetic workload of include/mc_synth.h directly in HBM for
 * individuals [i_first, i_first + I) (bench sizes never touch the host).
 * Recodes like the reference parser would; J is available afterwards. */
int mc_set_data_synth(mc_ctx *ctx, int64_t I, int32_t L, const mcs_params *g,
	int64_t i_first);
int mc_get_dims(const mc_ctx *ctx, int64_t *I, int32_t *L, int32_t *P,
	int64_t *T);
int mc_get_J(const mc_ctx *ctx, int32_t *J);
/* download natural-layout codes [I][L][P] (tests, round trips) */
int mc_get_codes(mc_ctx *ctx, uint8_t *codes);

/* ---- model: replaces allocate_model_for_k (multiclust.c:1181-1279) ------ */

/* q = number of secant pairs kept (0 without acceleration, else opt->q);
 * eta_lb / p_lb are the bounds of synchronize() (multiclust.c:812-815). */
int mc_alloc_model(mc_ctx *ctx, int32_t K, int admixture, int eta_constrained,
	int q, double eta_lb, double p_lb, int do_projection);
int mc_eta_len(const mc_ctx *ctx, int64_t *n);
int mc_set_params(mc_ctx *ctx, int slot, const double *eta, const double *p);
int mc_get_params(mc_ctx *ctx, int slot, double *eta, double *p);

/* random_initialize_admixture (rnd_init.c:349-357, 456-482): `z` is a HOST
 * array [I][L][P] holding the cluster drawn for every allele copy (missing
 * copies included, the reference draws for them too) in the reference's
 * i, l, a order -- the host owns the rand() stream.  The device forms the hard
 * assignment d_iklj = 1 for every (k, allele) pair that occurs in individual i
 * at locus l (set, not incremented: two copies of one allele drawn to the same
 * k count once) and runs the M-step (with projections) into `slot`. */
int mc_init_admixture(mc_ctx *ctx, int slot, const uint8_t *z);
/* the same split for an individual-sharded fit: counts of this context's
 * individuals (eta rows of `slot` are final, the allele counts are left in
 * the exchange buffer); sum over ranks, then mc_em_step_finish(slot) */
int mc_init_admixture_local(mc_ctx *ctx, int slot, const uint8_t *z);
/* The same initialiser with the draws made on the device (SURVEY.md 8f rank 1:
 * at I*L*P = 2e9 copies the host loop over rand() costs more than the fit).
 * The reference's generator is glibc's TYPE_3 rand(): x[n] = x[n-31] + x[n-3]
 * mod 2^32, rand() = x[n] >> 1 (rnd_init.c:460-481 draws k = rand() % K per
 * copy).  The host cuts this context's I*L*P draws into n_blocks blocks of
 * block_draws (a multiple of 16; the last block may be short) and hands over,
 * for every block, the 31 words x[n-31 .. n-1] in front of its first draw
 * (hist[n_blocks][31]); it owns the stream and advances it by the same number
 * of draws (host/mc_rand.h: mcr_jump_matrix / mcr_apply).  Results are
 * bit-identical to mc_init_admixture with the z the host loop would draw. */
int mc_init_admixture_rand(mc_ctx *ctx, int slot, const uint32_t *hist,
	int64_t n_blocks, int64_t block_draws);
int mc_init_admixture_rand_local(mc_ctx *ctx, int slot, const uint32_t *hist,
	int64_t n_blocks, int64_t block_draws);

/* random_initialize_mixture (rnd_init.c:192-339) with the I*K*L distance work
 * on the device.  The host draws the K distinct centre
 * individuals from its rand() stream (rnd_init.c:205-217; K draws plus
 * re-draws on collision) and hands over their genotype rows: center_codes is
 * a HOST array [K][L][P]; center_idx[k] is the centre's row among THIS
 * context's individuals, or -1 when it lives in another shard.  Every
 * individual joins the nearest centre by the L1 distance of the allele counts
 * (integers: the strict "smaller wins, first on ties" rule is exact), a centre
 * keeps its own cluster; then eta_k = (1 + n_k) / (I + K) and p_klj
 * proportional to 1 + (K - k) S_klj, row-normalised, go to `slot` (no
 * projection, like the reference).  The _local / _finish pair splits the call
 * for individual-sharded fits exactly like mc_em_step: the counts and cluster
 * sizes are summed over ranks through the exchange buffer in between;
 * I_total is the number of individuals over all shards.  K == 1: everybody in
 * cluster 0, the centre arguments may be NULL. */
int mc_init_mixture(mc_ctx *ctx, int slot, const int32_t *center_idx,
	const uint8_t *center_codes);
int mc_init_mixture_local(mc_ctx *ctx, const int32_t *center_idx, const uint8_t *center_codes);
int mc_init_mixture_finish(mc_ctx *ctx, int slot, int64_t I_total);

/* ---- parametric bootstrap (bootstrap.c:31-175, multiclust.c:562-581, 675-708;
 *      SURVEY.md 8f rank 4) ------------------------------------------------- */

/* Keep the parameters of `slot` as the estimates under H0 that the bootstrap
 * samples are drawn from (mod->mle_pKLM and mle_etak / mle_etaik,
 * multiclust.c:562-581).  They survive mc_alloc_model. */
int mc_save_mle(mc_ctx *ctx, int slot);
/* Replace this context's genotype data by one parametric bootstrap sample
 * drawn from the saved estimates, with the reference's draws: glibc TYPE_3
 * rand() as in mc_init_admixture_rand -- hist[n_blocks][31] are the generator
 * words in front of every block of block_draws draws (a multiple of 16) --
 * r = rand() / RAND_MAX, and the inverse-CDF walks of bootstrap.c:92-117 /
 * 136-169 in the same left-to-right FP64 sums.  Draws, in stream order:
 *   admixture: for every individual, locus and allele copy one draw for the
 *              source cluster (eta_i., or the pooled eta with -c) and one for
 *              the allele (p_kl.): 2 I L P draws;
 *   mixture:   one draw per individual for its cluster, then one per locus and
 *              copy for the allele: I (1 + L P) draws.
 * As in the reference's default parse mode every copy is drawn (missing data
 * are filled in) and an allele slot is chosen among all J_l, phantom slot
 * included; a locus without any allele (J_l = 0) stays missing.  The original
 * data are kept: the admixture initialisers go on reading them, exactly like
 * the reference's random_allele_partition reads dat->IL, which bootstrap.c never
 * rewrites (rnd_init.c:460-481).  Frees the model; layouts are rebuilt by the
 * next mc_alloc_model. */
int mc_bootstrap_data(mc_ctx *ctx, const uint32_t *hist, int64_t n_blocks, int64_t block_draws);
/* back to the original data (cleanup_parametric_bootstrap, bootstrap.c:60-66) */
int mc_restore_data(mc_ctx *ctx);

/* ---- the hot path ------------------------------------------------------ */

/* E-step on slot `from`, M-step (+ simplex projection) into slot `to`;
 * returns the log likelihood of slot `from`, one step late like the
 * reference.  Replaces em_step() minus stop() (em_alg.c:195-207 ->
 * e_step_admixture_orig 291-486 + m_step_admixture_orig 592-754, or
 * e_step_mixture 763-897 + m_step_mixture 907-1011).  from == to is the
 * reference's in-place EM.  Leaves D_ik / v_ik of this E-step on the device. */
int mc_em_step(mc_ctx *ctx, int from, int to, double *ll);

/* log_likelihood() (log_likelihood.c:56-62, 96-232); touches no posterior. */
int mc_loglik(mc_ctx *ctx, int slot, double *ll);

/* Log likelihood left on the device by the last mc_em_step_finish / mc_loglik
 * call that was given ll == NULL (lets one host thread keep several devices
 * busy: launch everywhere first, read afterwards).  Synchronises the stream. */
int mc_read_ll(mc_ctx *ctx, double *ll);

/* Posterior sums of the last E-step: D_ik = sum_{l,j} d_iklj (admixture, what
 * write_file.c:359-381,446-459,525-543 reduce diklm to) or v_ik (mixture). */
int mc_get_posterior(mc_ctx *ctx, double *out);
/* Sums of the posterior rows (D_ik or v_ik) over the individuals of every
 * sampling locale, the numerators of the popq tables (write_file.c:446-459,
 * 658-666): `locale` is a HOST array [I] of 0..n_locales-1 for this context's
 * individuals, `out` a HOST array [n_locales][K].  Fixed summation order. */
int mc_locale_sums(mc_ctx *ctx, const int32_t *locale, int32_t n_locales, double *out);
/* argmax partition on device (write_file.c:350-382, 582-600) */
int mc_partition(mc_ctx *ctx, int32_t *I_K, int32_t *count_K);

/* secant pair `pair` (0..q-1): which = 0 -> u, 1 -> v;
 * delta = x[slot_t] - x[slot_f]  (em_2_steps, em_alg.c:1104-1161) */
int mc_delta(mc_ctx *ctx, int which, int pair, int slot_t, int slot_f);

/* step_size() sums (accel_em.c:142-184): out = {utu, utvu, vutvu} of pair.
 * eta_part / p_part are returned separately so that an individual-sharded
 * run can add the eta parts of all ranks; either pointer may be NULL. */
int mc_step_dots(mc_ctx *ctx, int pair, double eta_part[3], double p_part[3]);
/* qn_accelerated_update() sums (accel_em.c:291-310): {u[q1].u[q2], u[q1].v[q2]} */
int mc_qn_dots(mc_ctx *ctx, int q1, int q2, double eta_part[2], double p_part[2]);

/* accelerated_update() (accel_em.c:440-541) without its log_likelihood call:
 * qn1 == 0: x[t] = x[p] - 2 s u + s^2 (v - u)   (SQUAREM, -s 1..3)
 * qn1 != 0: x[t] = x[p] + u + s v               (QN q=1, -s 4)
 * followed by the projections when enabled. */
int mc_accel_update(mc_ctx *ctx, int qn1, int slot_t, int slot_p, int pair,
	double s);
/* qn_accelerated_update() (accel_em.c:364-415) without the dot products, the
 * inverse and the log_likelihood call: x[t] = x[p] + u[uindex], then for row
 * j and column n in the reference's cyclic order starting at delta_index:
 * x[t] += v[pair(j)] * Ainv[j*q+n] * cutu[n]; then the projections. */
int mc_qn_update(mc_ctx *ctx, int slot_t, int slot_p, int uindex,
	int delta_index, const double *Ainv, const double *cutu);

/* simplex_project_eta / simplex_project_pklm on a whole slot (simplex.c) */
int mc_project(mc_ctx *ctx, int slot);
int mc_copy_slot(mc_ctx *ctx, int dst, int src);

/* ---- individual-sharded fits: one context per GPU holds a
 *      slice of individuals; p is replicated --------------------------------
 * mc_em_step_local runs the E-step and everything that needs only this
 * rank's individuals (eta rows of slot `to`, D_ik / v_ik), and leaves the
 * partial sums that must be added over ranks in the exchange buffer:
 * [K*T allele-count sums | ll | K pooled-eta sums], all double.  The caller
 * sums the buffer over ranks (NCCL all-reduce on `*dev_ptr`, or any other
 * mechanism; nothing to do for a single rank) and then calls
 * mc_em_step_finish, which normalises and projects p (and the pooled eta of
 * the mixture / -c models) into slot `to` on every rank identically.
 * mc_em_step == mc_em_step_local + mc_em_step_finish.
 * mc_em_step_local leaves the eta side of an admixture step (which needs no
 * remote data) running on a second stream beside the caller's exchange;
 * mc_em_step_finish joins it, so the pair must be called together. */
int mc_em_step_local(mc_ctx *ctx, int from, int to);
int mc_exchange_buffer(mc_ctx *ctx, void **dev_ptr, size_t *n_doubles);
int mc_em_step_finish(mc_ctx *ctx, int to, double *ll);
/* Deterministic alternative to an all-reduce: `gathered` is a DEVICE buffer
 * holding the exchange buffers of all ranks back to back (the result of an
 * all-gather, n_ranks * n_doubles); they are added in rank order into this
 * context's exchange buffer, so every rank computes bit-identical sums
 * whatever algorithm the collective library picked. */
int mc_exchange_sum(mc_ctx *ctx, const void *gathered, int n_ranks);
/* The same sum as a reduce-scatter + all-gather, whose traffic per rank does
 * not grow with the number of ranks.  The exchange buffer has room for 64
 * doubles beyond n_doubles (zero), so it can be cut into n_ranks <= 64 equal
 * slices of m = ceil(n_doubles / n_ranks): every rank sends slice j to rank j
 * (all-to-all), rank r adds the n_ranks copies of slice r it received (`parts`,
 * DEVICE memory, [n_ranks][count], rank order) into its own buffer at `first`
 * = r * m with this call, and the slices are all-gathered back in place. */
int mc_exchange_sum_slice(mc_ctx *ctx, const void *parts, int n_ranks,
	int64_t first, int64_t count);

/* ---- introspection for bench / profiles --------------------------------- */

typedef struct {
	int32_t K, k_split, k_per_lane;		/* lanes per locus, k values per lane */
	int32_t loci_per_warp, warps, groups;	/* tile = warps*groups*loci_per_warp loci */
	int32_t n_tiles, n_chunks, n_units, grid, block;
	int32_t indiv_per_block, ploidy_padded;
	int32_t two_pass;	/* 3: the dense DMMA kernels (mc_dense.cuh), 2: the two-pass
				 * admixture kernel (mc_admix3.cuh), 0: the one-pass tile kernel */
	int32_t reserved;
	int64_t smem_bytes;
	int64_t algorithmic_bytes_em;	/* I*L*P + 16*I*K + 16*K*T */
	int64_t algorithmic_bytes_ll;	/* I*L*P +  8*I*K +  8*K*T */
} mc_plan_info;
int mc_get_plan(const mc_ctx *ctx, mc_plan_info *out);

/* kernels launched through this context so far */
int64_t mc_launch_count(const mc_ctx *ctx);
/* When enabled, the dominant (genotype-streaming) kernel of every
 * mc_em_step / mc_loglik is bracketed by CUDA events on the launch stream. */
int mc_profile_enable(mc_ctx *ctx, int on);
int mc_profile_read(mc_ctx *ctx, int64_t *n_launches, double *total_ms);

#ifdef __cplusplus
}
#endif

#endif /* MC_CUDA_H */
