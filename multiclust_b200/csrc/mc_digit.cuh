/*
 * mc_digit.cuh -- the mixture model's two contractions as exact integer GEMMs
 * on the integer tensor path (IMMA, mma.sync m16n8k32 u8 x u8 -> s32).
 *
 * Reference: em_alg.c:782-826 (E-step, a_ik = sum_l sum_a c_ila log p_kla),
 * log_likelihood.c:186-201 (the same sum for logL_mixture) and em_alg.c:964-990
 * (M-step, N_kla = sum_i v_ik c_ila).  One factor of either contraction is a
 * matrix of allele counts c in 0..P <= 15, the other an FP64 table in a known
 * range (|log p| < 1024 for every positive double; 0 <= v <= 1).  Written as a
 * 64-bit fixed-point number the table splits into eight 8-bit digits,
 *	x = sum_d digit_d(x) 256^d,
 * and count x digit products accumulate exactly in 32-bit integers.  One kernel
 * serves both passes:
 *	D[row][(k, d)] = sum_kk C[row][kk] digit_d(X[kk][k])
 *	E pass: row = individual, kk = (locus, allele), X = |log p| 2^54
 *	M pass: row = (locus, allele), kk = individual, X = v 2^(64 - e_k),
 *	        max_i v_ik < 2^e_k
 * with n-tile j of the MMA = class k and column d of the tile = digit d, so a
 * thread quad holds the eight digits of one (row, k) and recombines them with
 * three shuffles per four values.  Sums are exact up to the final rounding to FP64, i.e. at
 * least as accurate as any FP64 summation order (DESIGN.md has the bound), and
 * the integer tensor path runs 32 MACs for each FP64 FMA of the DMMA path
 * (tools/imma_probe.cu, profiles/r02_imma_probe.txt).
 *
 * Data with more than two alleles at a locus use the same kernel on COLUMN
 * pairs instead of (locus, allele) pairs (`general`): the count matrix is
 * [I][T] over the allele columns off_l + j, two adjacent columns per byte, and
 * "locus l, allele h" below reads "column 2 B + h of byte B".  Biallelic data keep
 * the locus form: a phantom slot (J_l = 3) costs no column there.
 *
 * Layout (built once per data set by mc_digit_build.cuh):
 *	cnt [m-tile][block][2][32] uint4 -- 16 rows x 64 count bytes (c0 | c1 << 4)
 *	    in fragment order: a warp's two 512-byte loads are contiguous.
 *	    E: m-tile = 16 individuals, block = 64 loci; lane (g, t), half h holds
 *	       individual 16 mt + 8 h + g, loci 64 b + 16 t + 4 s + jj in byte jj of
 *	       word s.
 *	    M: m-tile = 8 loci (row g = allele 0, row g + 8 = allele 1), block = 128
 *	       individuals; lane (g, t), half h holds locus 8 mt + g, individuals
 *	       128 b + 32 t + 16 h + 4 s + jj.
 *	tab [block][4 steps][K][32] uint2 -- the B fragments of the digit table,
 *	    rebuilt from p (E) or the posteriors (M) before every pass.
 * A CTA is four warps on one chunk of blocks; each warp owns R m-tiles and keeps
 * R x K accumulator tiles in registers.  cp.async stages a block three blocks
 * ahead in shared memory: its B fragments, shared by the warps, and each warp's
 * own counts (at 0.7 us of MMAs per block one block of look-ahead left the HBM
 * latency exposed, profiles/r02_ncu_digit_kernels.txt).
 */
#pragma once

#include <cstddef>
#include <cstdint>

#include "mc_device.cuh"

#define DG_THREADS 128
#define DG_WARPS (DG_THREADS / 32)
enum { DG_MIX_E = 0, DG_MIX_M = 1 };

/* m-tiles per warp: accumulators are R x K x 4 registers */
__host__ __device__ constexpr int dg_R(int K)
{
	return K <= 5 ? 4 : K <= 8 ? 3 : K <= 12 ? 2 : 1;
}

struct DigitArgs {
	int K, L;
	int general;			/* 0: biallelic (locus, allele) pairs; 1: column pairs */
	int n_mtiles, n_blocks, n_chunks, n_ctarows;
	long long I, Ipad, T;
	const int *off, *J;		/* [L + 1], [L] */
	const uint4 *cnt;
	const uint2 *tab;
	const int *skip_if;		/* nullable: nonzero = table not representable */
	int *n_chunks_dev;		/* nullable: the chunk count for k_mix_tail */
	const double *unscale;		/* M: [K] 2^(e_k - 64), the table's column scaling */
	double *out;			/* E: Apart [n_chunks][Ipad][K]; M: Npart [n_chunks][K T] */
};

__device__ __forceinline__ void dg_imma(int *c, unsigned a0, unsigned a1, unsigned a2,
	unsigned a3, uint2 b)
{
	asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
		: "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
		: "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b.x), "r"(b.y));
}

/* A thread's two digits (2t, 2t + 1) of one (row, class) as a double: exact,
 * lo + 256 hi < 2^40 and the power of two only shifts the exponent */
__device__ __forceinline__ double dg_pair(int lo, int hi, double scale_t)
{
	return fma((double)hi, 256.0, (double)lo) * scale_t;
}

/* Four values, each spread over the four lanes t of a quad: two exchange rounds
 * leave lane t with the quad's sum of value 2 (t & 1) + (t >> 1).  Three
 * shuffles for four sums; the order of the additions is fixed. */
__device__ __forceinline__ double dg_quad_sum4(double x0, double x1, double x2, double x3, int t)
{
	const bool odd = t & 1, up = t & 2;
	const double ya = (odd ? x2 : x0) + shfl_xor_f64(odd ? x0 : x2, 1);
	const double yb = (odd ? x3 : x1) + shfl_xor_f64(odd ? x1 : x3, 1);
	return (up ? yb : ya) + shfl_xor_f64(up ? ya : yb, 2);
}

__device__ __forceinline__ void dg_cp_async16(void *smem_dst, const void *gsrc)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gsrc));
}

/* shared memory: DG_STAGES blocks, each the B fragments (K KB, shared by the
 * warps) and every warp's own R m-tiles of counts (R KB) */
__host__ __device__ constexpr int dg_stage_u4(int K)
{
	return 4 * K * 16 + DG_WARPS * dg_R(K) * 64;		/* in uint4 */
}
/* two CTAs per SM (up to 255 registers a thread); a third one -- 168 registers,
 * three stages -- measured no faster at K = 5 */
#define DG_STAGES 4
static inline size_t dg_smem_bytes(int K)
{
	return (size_t)DG_STAGES * dg_stage_u4(K) * sizeof(uint4);
}

template <int K, int MODE>
__global__ void __launch_bounds__(DG_THREADS, 2) digit_kernel(const DigitArgs a)
{
	constexpr int R = dg_R(K);
	constexpr int TAB_U4 = 4 * K * 16;		/* uint4 per block of the table */
	constexpr int STAGE_U4 = dg_stage_u4(K);
	extern __shared__ __align__(16) uint4 dg_sm[];	/* [DG_STAGES][table | counts] */
	if (a.skip_if && *a.skip_if)
		return;
	if (a.n_chunks_dev && blockIdx.x == 0 && threadIdx.x == 0)
		*a.n_chunks_dev = a.n_chunks;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int g = lane >> 2, t = lane & 3;
	const int c = blockIdx.x / a.n_ctarows, ctarow = blockIdx.x - c * a.n_ctarows;
	const int mt0 = (ctarow * DG_WARPS + w) * R;
	/* a warp past the last m-tile still stages the table and meets the barriers */
	const bool active = mt0 < a.n_mtiles;
	const int b0 = (int)((long long)a.n_blocks * c / a.n_chunks);
	const int b1 = (int)((long long)a.n_blocks * (c + 1) / a.n_chunks);

	int acc[R][K][4];
#pragma unroll
	for (int r = 0; r < R; r++)
#pragma unroll
		for (int j = 0; j < K; j++)
#pragma unroll
			for (int e = 0; e < 4; e++)
				acc[r][j][e] = 0;

	/* block b sits in stage b % DG_STAGES, copied DG_STAGES - 1 blocks ahead:
	 * the table by the whole CTA, a warp's counts by the lanes that read them
	 * (an m-tile past the end repeats the last one; its sums are not written) */
	const uint4 *cnt_w[R];
#pragma unroll
	for (int r = 0; r < R; r++) {
		const int mt = min(mt0 + r, a.n_mtiles - 1);
		cnt_w[r] = a.cnt + (size_t)mt * a.n_blocks * 64 + lane;
	}
	auto stage = [&](int b) {
		if (b < b1) {
			uint4 *dst = dg_sm + (size_t)(b % DG_STAGES) * STAGE_U4;
			const uint4 *src = reinterpret_cast<const uint4 *>(a.tab) + (size_t)b * TAB_U4;
			for (int x = threadIdx.x; x < TAB_U4; x += DG_THREADS)
				dg_cp_async16(dst + x, src + x);
			if (active) {
				uint4 *dw = dst + TAB_U4 + w * R * 64 + lane;
#pragma unroll
				for (int r = 0; r < R; r++) {
					dg_cp_async16(dw + r * 64, cnt_w[r] + (size_t)b * 64);
					dg_cp_async16(dw + r * 64 + 32, cnt_w[r] + (size_t)b * 64 + 32);
				}
			}
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
#pragma unroll
	for (int x = 0; x < DG_STAGES - 1; x++)
		stage(b0 + x);
	for (int b = b0; b < b1; b++) {
		asm volatile("cp.async.wait_group %0;" :: "n"(DG_STAGES - 2) : "memory");
		__syncthreads();
		stage(b + DG_STAGES - 1);
		if (!active)
			continue;
		const uint4 *st = dg_sm + (size_t)(b % DG_STAGES) * STAGE_U4;
		const uint2 *tb = reinterpret_cast<const uint2 *>(st) + lane;
		const uint4 *aw = st + TAB_U4 + w * R * 64 + lane;
		uint4 cur[R][2];
#pragma unroll
		for (int r = 0; r < R; r++) {
			cur[r][0] = aw[r * 64];
			cur[r][1] = aw[r * 64 + 32];
		}
#pragma unroll
		for (int s = 0; s < 4; s++) {
			uint2 B[K];
#pragma unroll
			for (int j = 0; j < K; j++)
				B[j] = tb[(s * K + j) * 32];
#pragma unroll
			for (int r = 0; r < R; r++) {
				const unsigned w0 = s == 0 ? cur[r][0].x : s == 1 ? cur[r][0].y
					: s == 2 ? cur[r][0].z : cur[r][0].w;
				const unsigned w1 = s == 0 ? cur[r][1].x : s == 1 ? cur[r][1].y
					: s == 2 ? cur[r][1].z : cur[r][1].w;
				const unsigned m = 0x0f0f0f0fu;
				unsigned a0, a1, a2, a3;
				if (MODE == DG_MIX_E) {
					/* k slots 0..15: allele 0 of 16 loci, 16..31: allele 1;
					 * w0 = row g, w1 = row g + 8 */
					a0 = w0 & m; a1 = w1 & m;
					a2 = (w0 >> 4) & m; a3 = (w1 >> 4) & m;
				} else {
					/* k slots = 32 individuals (w0: first 16, w1: second
					 * 16); row g = allele 0, row g + 8 = allele 1 */
					a0 = w0 & m; a1 = (w0 >> 4) & m;
					a2 = w1 & m; a3 = (w1 >> 4) & m;
				}
#pragma unroll
				for (int j = 0; j < K; j++)
					dg_imma(acc[r][j], a0, a1, a2, a3, B[j]);
			}
		}
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");
	if (!active)
		return;

	/* ---- recombine the digits and write the chunk's partial sums ----
	 * value n = 2 j + h of an m-tile: class j, row g + 8 h; lane t of the quad
	 * ends up with value 4 q + 2 (t & 1) + (t >> 1) of every group q of four */
	const double scale_t = __longlong_as_double((long long)(1023 + 16 * t) << 52);
	const int own = 2 * (t & 1) + (t >> 1);
#pragma unroll
	for (int r = 0; r < R; r++) {
		const int mt = mt0 + r;
		if (mt >= a.n_mtiles)
			continue;
		int Jl = 0;
		long long ol = 0;
		if (MODE == DG_MIX_M) {
			if (a.general) {	/* row g + 8 h = allele column 2 (8 mt + g) + h */
				ol = 2LL * (mt * 8 + g);
				Jl = ol + 1 < a.T ? 2 : ol < a.T ? 1 : 0;
			} else {
				const int l = mt * 8 + g;
				if (l < a.L) {
					Jl = a.J[l];
					ol = a.off[l];
				}
			}
		}
#pragma unroll
		for (int q = 0; q < (2 * K + 3) / 4; q++) {
			double x[4];
#pragma unroll
			for (int e = 0; e < 4; e++) {
				const int n = 4 * q + e;
				x[e] = n < 2 * K ? dg_pair(acc[r][n < 2 * K ? n >> 1 : 0][2 * (n & 1)],
					acc[r][n < 2 * K ? n >> 1 : 0][2 * (n & 1) + 1], scale_t) : 0.0;
			}
			const double v = dg_quad_sum4(x[0], x[1], x[2], x[3], t);
			const int n = 4 * q + own, j = n >> 1, h = n & 1;
			if (n >= 2 * K)
				continue;
			if (MODE == DG_MIX_E) {
				/* a_ik = -(sum) 2^-54 */
				const long long i = (long long)mt * 16 + g + 8 * h;
				if (i < a.I)
					a.out[((size_t)c * a.Ipad + i) * K + j] = v * -5.5511151231257827e-17;
			} else {
				/* N_kla = (sum) 2^(e_k - 64); row g + 8 h = allele h */
				double *o = a.out + (size_t)c * K * a.T + (size_t)j * a.T + ol;
				if (h < Jl)
					o[h] = v * __ldg(a.unscale + j);
				/* the phantom slot: no copy carries it */
				if (h == 0 && Jl > 2)
					o[2] = 0.0;
			}
		}
	}
}
