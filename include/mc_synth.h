/*
 * mc_synth.h -- counter-based synthetic genotype generator (host C and CUDA).
 *
 * The reference's own simulator (multiclust.c:167-186, read_file.c:302-372) is
 * biallelic-only, ignores Q and cannot produce the multi-allelic / missing /
 * polyploid workloads of BASELINE.json (SURVEY.md finding 4, section 8d), so
 * the bench and the parity tests share this generator instead.  Every draw is
 * a pure function of (seed, indices) in 64-bit integer arithmetic, so the host
 * C build, the oracle and the CUDA fill kernel produce identical bytes.
 *
 * Model: locus l has n_l = mcs_nalleles() alleles; population k has allele
 * weights w_klj = u^2, u in [1, 2^20]; individual i has ancestry weights
 * q_ik = v^3, v in [1, 2^16]; each allele copy draws a population from q_i
 * and then an allele from w_k,l by inverse CDF; a copy is missing with
 * probability miss_bp / 10000.  Codes are 0..n_l-1, 255 = missing.
 */
#ifndef MC_SYNTH_H
#define MC_SYNTH_H

#include <stdint.h>

#ifdef __CUDACC__
#define MCS_FN __host__ __device__ static inline
#else
#define MCS_FN static inline
#endif

#define MC_MISSING_CODE 255

typedef struct {
	uint64_t seed;
	int32_t K;        /* number of source populations */
	int32_t jmax;     /* max alleles per locus (2 => biallelic everywhere) */
	int32_t miss_bp;  /* missing rate in basis points (500 = 5 %) */
	int32_t ploidy;
} mcs_params;

MCS_FN uint64_t mcs_mix(uint64_t x)
{
	x += 0x9e3779b97f4a7c15ULL;
	x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
	x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
	return x ^ (x >> 31);
}

MCS_FN uint64_t mcs_hash(uint64_t seed, uint64_t tag, uint64_t a, uint64_t b)
{
	uint64_t h = mcs_mix(seed ^ (tag * 0xd6e8feb86659fd93ULL));
	h = mcs_mix(h ^ a);
	h = mcs_mix(h ^ (b + 0x632be59bd9b4e019ULL));
	return h;
}

/* number of (nominal) alleles at locus l: 2 .. jmax */
MCS_FN int mcs_nalleles(const mcs_params *g, int64_t l)
{
	if (g->jmax <= 2)
		return 2;
	return 2 + (int)(mcs_hash(g->seed, 1, (uint64_t)l, 0) % (uint64_t)(g->jmax - 1));
}

/* allele weight of allele j at locus l in population k (40-bit integer) */
MCS_FN uint64_t mcs_allele_weight(const mcs_params *g, int k, int64_t l, int j)
{
	uint64_t u = 1 + (mcs_hash(g->seed, 2, (uint64_t)l * 4096u + (uint64_t)k, (uint64_t)j) >> 44);
	return u * u;
}

/* ancestry weight of population k in individual i (48-bit integer) */
MCS_FN uint64_t mcs_ancestry_weight(const mcs_params *g, int64_t i, int k)
{
	uint64_t v = 1 + (mcs_hash(g->seed, 3, (uint64_t)i, (uint64_t)k) >> 48);
	return v * v * v;
}

/* allele code of copy a of individual i at locus l */
MCS_FN uint8_t mcs_code(const mcs_params *g, int64_t i, int64_t l, int a)
{
	uint64_t cell = (uint64_t)i * 0x100000001b3ULL + (uint64_t)l;
	uint64_t r, tot, acc;
	int k, j, nl, kk = 0, jj = 0;

	if (g->miss_bp > 0 &&
	    (int)(mcs_hash(g->seed, 4, cell, (uint64_t)a) % 10000u) < g->miss_bp)
		return MC_MISSING_CODE;

	tot = 0;
	for (k = 0; k < g->K; k++)
		tot += mcs_ancestry_weight(g, i, k);
	r = mcs_hash(g->seed, 5, cell, (uint64_t)a) % tot;
	acc = 0;
	for (k = 0; k < g->K; k++) {
		acc += mcs_ancestry_weight(g, i, k);
		if (r < acc) {
			kk = k;
			break;
		}
	}

	nl = mcs_nalleles(g, l);
	tot = 0;
	for (j = 0; j < nl; j++)
		tot += mcs_allele_weight(g, kk, l, j);
	r = mcs_hash(g->seed, 6, cell, (uint64_t)a) % tot;
	acc = 0;
	for (j = 0; j < nl; j++) {
		acc += mcs_allele_weight(g, kk, l, j);
		if (r < acc) {
			jj = j;
			break;
		}
	}
	return (uint8_t)jj;
}

#endif /* MC_SYNTH_H */
