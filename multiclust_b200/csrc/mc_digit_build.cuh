/*
 * mc_digit_build.cuh -- layout and table builders of the digit-sliced mixture
 * path (mc_digit.cuh).  Included by mc_cuda.cu only.
 */
#pragma once

#include "mc_digit.cuh"

/* c0 | c1 << 4 of one (individual, locus) from the natural [I][L][P] codes */
__device__ __forceinline__ unsigned dg_count_byte(const unsigned char *nat, long long i,
	int l, long long I, int L, int P)
{
	if (i >= I || l >= L)
		return 0u;
	const unsigned char *c = nat + ((size_t)i * L + l) * P;
	unsigned c0 = 0, c1 = 0;
	for (int a = 0; a < P; a++) {
		c0 += c[a] == 0;
		c1 += c[a] == 1;
	}
	return c0 | c1 << 4;
}

/* counts of the allele columns 2 B and 2 B + 1 of individual i (general form):
 * column t belongs to locus col_locus[t] and stands for allele code t - off_l */
__device__ __forceinline__ unsigned dg_count_cols(const unsigned char *nat, long long i,
	long long B, long long I, int L, int P, long long T, const int *col_locus, const int *off)
{
	unsigned byte = 0u;
	if (i >= I)
		return 0u;
	for (int h = 0; h < 2; h++) {
		const long long col = 2 * B + h;
		if (col >= T)
			break;
		const int l = col_locus[col];
		const unsigned code = (unsigned)(col - off[l]);
		const unsigned char *c = nat + ((size_t)i * L + l) * P;
		unsigned n = 0;
		for (int a = 0; a < P; a++)
			n += c[a] == code;
		byte |= n << (4 * h);
	}
	return byte;
}

/* one thread per uint4 of the fragment-ordered count array (mc_digit.cuh).
 * mode DG_MIX_E: m-tile = 16 individuals, block = 64 loci;
 * mode DG_MIX_M: m-tile = 8 loci, block = 128 individuals.
 * general: "locus" reads "byte of two adjacent allele columns". */
__global__ void k_digit_counts(const unsigned char *nat, uint4 *cnt, long long I, int L,
	int P, int n_mtiles, int n_blocks, int mode, int general, long long T,
	const int *col_locus, const int *off)
{
	const long long n = (long long)n_mtiles * n_blocks * 64;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int lane = (int)(x & 31), h = (int)((x >> 5) & 1);
		const int g = lane >> 2, t = lane & 3;
		const long long tb = x >> 6;
		const int b = (int)(tb % n_blocks);
		const long long mt = tb / n_blocks;
		unsigned w[4];
		for (int s = 0; s < 4; s++) {
			w[s] = 0u;
			for (int jj = 0; jj < 4; jj++) {
				long long i;
				long long l;
				if (mode == DG_MIX_E) {
					i = mt * 16 + 8 * h + g;
					l = (long long)b * 64 + 16 * t + 4 * s + jj;
				} else {
					l = mt * 8 + g;
					i = (long long)b * 128 + 32 * t + 16 * h + 4 * s + jj;
				}
				/* general form: `l` is the byte (column pair) index */
				if (general)
					w[s] |= dg_count_cols(nat, i, l, I, L, P, T, col_locus, off) << (8 * jj);
				else if (l < L)
					w[s] |= dg_count_byte(nat, i, (int)l, I, L, P) << (8 * jj);
			}
		}
		cnt[x] = make_uint4(w[0], w[1], w[2], w[3]);
	}
}

/* The B fragments of the digit table.  One warp per (block, step, class): the
 * fragment of lane (g, t) holds digit g of the eight k slots 4 t + jj and
 * 16 + 4 t + jj (jj = 0..3) of the step.  Lane (g, t) computes the fixed-point
 * value of one of the eight slots of thread column t -- (half = g / 4,
 * jj = g % 4) -- then the eight lanes of a column exchange their values and
 * every lane keeps byte g of each.
 *
 * E table (take_log 1 or 2, k_dense_p's rule): slot = (locus 64 b + 16 t + 4 s +
 * jj, allele half) -- column 2 (64 b + 16 t + 4 s + jj) + half in the general
 * form -- X = |log p_kla| 2^54; p == 0 with take_log 1 contributes
 * nothing; any value outside [0, 1024) -- log 0 of the log-likelihood pass, a
 * NaN -- raises *flag and the pass falls back to the FP64 kernels.
 * M table (take_log 0): slot = individual 128 b + 32 t + 16 half + 4 s + jj,
 * X = v_ik 2^(64 - e_k), vscale[k] = 2^(64 - e_k) with max_i v_ik < 2^e_k
 * (k_mix_final): 64 bits below the largest posterior of the class, so a nearly
 * empty class keeps its relative precision. */
__global__ void k_digit_table(const double *src, const double *vscale, uint2 *tab, int *flag,
	int take_log, int K, long long I, int L, long long T, const int *off, const int *J,
	int n_blocks, int general)
{
	const int lane = threadIdx.x & 31;
	const int g = lane >> 2, t = lane & 3;
	const int half = g >> 2, jj = g & 3;
	const long long n = (long long)n_blocks * 4 * K;
	const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	bool bad = false;
	for (long long x = warp0; x < n; x += nwarps) {
		const int k = (int)(x % K);
		const int s = (int)((x / K) & 3);
		const long long b = x / (4 * K);
		unsigned long long X = 0ull;
		if (take_log) {
			const long long l = b * 64 + 16 * t + 4 * s + jj;
			/* general form: column 2 l + half of the [K][T] table */
			const bool valid = general ? 2 * l + half < T : l < L && half < J[l];
			if (valid) {
				const double p = src[(size_t)k * T + (general ? 2 * l : off[l]) + half];
				if (!(take_log == 1 && p == 0.0)) {
					const double lp = -log(p);
					if (lp >= 0.0 && lp < 1024.0)
						X = __double2ull_rn(lp * 18014398509481984.0);
					else if (lp < 0.0 && lp > -1e-14)
						X = 0ull;	/* p a rounding error above 1 */
					else
						bad = true;
				}
			}
		} else {
			const long long i = b * 128 + 32 * t + 16 * half + 4 * s + jj;
			if (i < I) {
				/* v <= max_i v_ik < 2^e_k: the product is below 2^64 */
				const double v = fmax(src[(size_t)i * K + k], 0.0) * vscale[k];
				X = v < 18446744073709551616.0 ? __double2ull_rn(v)
					: 0xffffffffffffffffull;
			}
		}
		unsigned b0 = 0u, b1 = 0u;
#pragma unroll
		for (int r = 0; r < 8; r++) {
			/* value of slot (half = r / 4, jj = r % 4) of column t */
			const int from = (r << 2) | t;
			const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)X, from);
			const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(X >> 32), from);
			const unsigned long long Y = ((unsigned long long)hi << 32) | lo;
			const unsigned d = (unsigned)(Y >> (8 * g)) & 0xffu;
			if (r < 4)
				b0 |= d << (8 * r);
			else
				b1 |= d << (8 * (r - 4));
		}
		tab[x * 32 + lane] = make_uint2(b0, b1);
	}
	if (bad)
		*flag = 1;
}
