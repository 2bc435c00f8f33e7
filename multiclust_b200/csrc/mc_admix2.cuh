/*
 * mc_admix2.cuh -- two-pass admixture kernel: both sufficient statistics are
 * accumulated in REGISTERS and shared memory only carries read-only operands.
 *
 * The fused E+M step of the admixture model (em_alg.c:325-433, 604-725) is,
 * per allele copy (i, l, a) with allele j:
 *     tmp = sum_k eta_ik p_klj,  w = c / tmp,  ll += c log tmp,
 *     A_ik += p_klj w   (-> D_ik = eta_ik A_ik,  the eta update)
 *     G_klj += eta_ik w (-> N_klj = p_klj G_klj, the p update)
 * One side of every copy is an individual, the other an allele column; a
 * thread can keep only one of them in registers.  Instead of making the other
 * side a read-modify-write of shared memory (3K 8-byte words per copy, the
 * round-1 kernel), a CTA walks (IT individuals x LT loci) tiles twice:
 *   pass 1  thread <-> (individual, a share of the tile's loci): eta_i and A_i
 *           in registers; the p rows are read from a shared read-only tile
 *           (lanes are different individuals at the SAME locus, so a warp's
 *           addresses fall on at most J_l rows); w is parked in shared memory;
 *   pass 2  thread <-> (allele column, segment): G_lj in registers; the
 *           column's copies come from an allele-sorted entry list built once
 *           at upload (`k_build_csc`), w from shared memory, eta_i from a
 *           shared read-only tile.
 * Shared memory traffic per copy drops from 3K read-modify-write words to
 * 2K + 2 read-only words, no table is thread-private (512 threads per CTA
 * whatever the number of alleles), A_i never leaves registers while the CTA
 * sweeps its chunk of loci, and G is flushed once per tile into a per-CTA
 * shared accumulator by the column's owner: no atomics, fixed order.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define A2_IT 256		/* individuals per tile (8-bit index in the lists) */
#ifndef A2_THREADS
#define A2_THREADS 256
#endif
#define A2_H (A2_THREADS / A2_IT)	/* threads sharing an individual in pass 1 */
#ifndef A2_CTAS_PER_SM
#define A2_CTAS_PER_SM 2	/* two CTAs per SM: one's pass 1 (FP64) overlaps the other's pass 2 (LDS) */
#endif

struct Admix2Args {
	int K, KR;			/* KR = K rounded up to even */
	int LT, LH;			/* loci per tile, per pass-1 thread */
	int n_itiles, n_ltiles, n_lchunks, n_ichunks, n_units;
	long long I, Ipad, T;
	int L, ncolmax, max_chunk_rows, max_tile_rows;
	/* per locus tile (static) */
	const int *lt_first;		/* [n_ltiles] first locus */
	const int *lt_ncol;		/* [n_ltiles] real allele columns */
	const int *lt_S;		/* [n_ltiles] segments per column (power of 2) */
	const unsigned short *colinfo;	/* [n_ltiles][ncolmax] locus_in_tile << 8 | allele */
	/* per locus chunk */
	const int *lc_first;		/* [n_lchunks + 1] first ltile of each chunk */
	const int *off;			/* [L + 1] prefix sums of J */
	/* data */
	const unsigned char *csr;	/* [n_itiles][n_ltiles][A2_THREADS][8] */
	const unsigned short *csc;	/* [n_itiles][n_ltiles][cap] sorted entries */
	const unsigned short *colstart;	/* [n_itiles][n_ltiles][ncolmax + 1] */
	int cap;			/* entries per tile: A2_IT * LT * PP */
	/* parameters */
	const double *p, *eta;
	long long eta_stride;
	/* outputs */
	double *Apart;			/* [n_lchunks * A2_H][Ipad][K] */
	double *Npart;			/* [n_ichunks][K*T] */
	double *llpart;			/* [n_units] */
};

/* ---------------------------------------------------------------------- */
/* one-time layout builders                                                 */

/* natural [I][L][P] codes -> pass-1 units: 8 bytes = LH loci x PP copies of
 * one individual, thread-major inside a tile */
__global__ void k_build_csr(const unsigned char *nat, unsigned char *csr,
	long long I, int L, int P, int PP, int LT, int LH, int n_itiles, int n_ltiles)
{
	const long long n = (long long)n_itiles * n_ltiles * A2_THREADS;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int t = (int)(x % A2_THREADS);
		const long long r = x / A2_THREADS;
		const int lt = (int)(r % n_ltiles);
		const long long it = r / n_ltiles;
		const long long i = it * A2_IT + t % A2_IT;
		const int h = t / A2_IT;
		unsigned char b[8];
		for (int q = 0; q < 8; q++) {
			const int l = lt * LT + h * LH + q / PP, a = q % PP;
			b[q] = (i < I && l < L && a < P && q / PP < LH)
				? nat[((size_t)i * L + l) * P + a] : 255;
		}
		*reinterpret_cast<uint2 *>(csr + (size_t)x * 8) = *reinterpret_cast<uint2 *>(b);
	}
}

/* how often each allele slot occurs (orders the columns of a locus tile) */
__global__ void k_allele_hist(const unsigned char *nat, long long I, int L, int P,
	const int *off, unsigned *hist)
{
	const long long n = I * (long long)L;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int l = (int)(x % L);
		for (int ap = 0; ap < P; ap++) {
			const unsigned char c = nat[(size_t)x * P + ap];
			if (c != 255)
				atomicAdd(&hist[off[l] + c], 1u);
		}
	}
}

/* allele-sorted entry lists of one (itile, ltile): for every real allele
 * column in `colinfo` order, the individuals carrying it in ascending order,
 * one entry per (individual, allele): i | first copy << 8 | (count-1) << 12 */
__global__ void k_build_csc(const unsigned char *nat, long long I, int L, int P,
	int PP, int LT, int n_ltiles, int ncolmax, int cap, const int *lt_ncol,
	const unsigned short *colinfo, unsigned short *csc, unsigned short *colstart)
{
	extern __shared__ unsigned char sm[];
	unsigned char *codes = sm;				/* [A2_IT][LT][PP] */
	int *cnt = reinterpret_cast<int *>(sm + ((size_t)A2_IT * LT * PP + 15) / 16 * 16);
	const int lt = blockIdx.x % n_ltiles;
	const long long it = blockIdx.x / n_ltiles;
	const int ncol = lt_ncol[lt];
	const unsigned short *ci = colinfo + (size_t)lt * ncolmax;
	unsigned short *out = csc + (size_t)blockIdx.x * cap;
	unsigned short *cs = colstart + (size_t)blockIdx.x * (ncolmax + 1);

	for (int x = threadIdx.x; x < A2_IT * LT * PP; x += blockDim.x) {
		const int a = x % PP, ll = (x / PP) % LT, ii = x / (PP * LT);
		const long long i = it * A2_IT + ii;
		const int l = lt * LT + ll;
		codes[x] = (i < I && l < L && a < P) ? nat[((size_t)i * L + l) * P + a] : 255;
	}
	__syncthreads();
	for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
		const int ll = ci[c] >> 8, j = ci[c] & 0xff;
		int n = 0;
		for (int ii = 0; ii < A2_IT; ii++) {
			bool has = false;
			for (int a = 0; a < PP; a++)
				has |= codes[(ii * LT + ll) * PP + a] == j;
			n += has;
		}
		cnt[c] = n;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		int acc = 0;
		for (int c = 0; c < ncol; c++) {
			const int n = cnt[c];
			cnt[c] = acc;
			cs[c] = (unsigned short)acc;
			acc += n;
		}
		for (int c = ncol; c <= ncolmax; c++)
			cs[c] = (unsigned short)acc;
	}
	__syncthreads();
	for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
		const int ll = ci[c] >> 8, j = ci[c] & 0xff;
		int pos = cnt[c];
		for (int ii = 0; ii < A2_IT; ii++) {
			int n = 0, first = 0;
			for (int a = PP - 1; a >= 0; a--)
				if (codes[(ii * LT + ll) * PP + a] == j) {
					n++;
					first = a;
				}
			if (n)
				out[pos++] = (unsigned short)(ii | first << 8 | (n - 1) << 12);
		}
	}
}

/* ---------------------------------------------------------------------- */

__device__ __forceinline__ void a2_cp_async16(void *smem_dst, const void *gsrc)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gsrc));
}
__device__ __forceinline__ void a2_cp_async_wait()
{
	asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

/* MODE 0: E+M step, MODE 1: log likelihood only */
template <int KP, int PP, int MODE>
__global__ void __launch_bounds__(A2_THREADS, A2_CTAS_PER_SM) admix2_kernel(const Admix2Args a)
{
	constexpr int KR = 2 * KP;
	constexpr bool EM = (MODE == 0);
	extern __shared__ double smem[];
	const int t = threadIdx.x, lane = t & 31;
	/* pass 1: a warp is 32 consecutive individuals at the SAME loci */
	const int ii = t % A2_IT, h = t / A2_IT;
	const int LT = a.LT, LH = a.LH;

	/* shared memory carve-up (doubles first, then 16-bit tables) */
	double *B_s = smem;						/* [max_chunk_rows][KR] */
	double *p_s = B_s + (EM ? (size_t)a.max_chunk_rows * KR : 0);	/* [max_tile_rows][KR] */
	double *eta_s = p_s + (size_t)a.max_tile_rows * KR;		/* [A2_IT][KR] */
	double *w_s = eta_s + (EM ? (size_t)A2_IT * KR : 0);		/* [LT*PP][A2_IT] */
	double *red = w_s + (EM ? (size_t)a.cap : 0);			/* [A2_THREADS/32] */
	unsigned short *csc_s = reinterpret_cast<unsigned short *>(red + A2_THREADS / 32);
	unsigned short *cst_s = csc_s + (EM ? a.cap : 0);		/* [ncolmax + 1] */
	int *rb_s = reinterpret_cast<int *>(cst_s + ((a.ncolmax + 1 + 7) / 8) * 8);	/* [LT] */

	for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
		const int c = u % a.n_lchunks, r = u / a.n_lchunks;
		const int lt0 = a.lc_first[c], lt1 = a.lc_first[c + 1];
		const long long it0 = (long long)a.n_itiles * r / a.n_ichunks;
		const long long it1 = (long long)a.n_itiles * (r + 1) / a.n_ichunks;
		const int l0 = a.lt_first[lt0];
		const int lend = lt1 < a.n_ltiles ? a.lt_first[lt1] : a.L;
		const int row0 = a.off[l0];
		const int chunk_rows = a.off[lend] - row0;
		double prod = 1.0, ll_slow = 0.0;
		long long esum = 0;

		__syncthreads();
		if (EM)
			for (int x = t; x < chunk_rows * KR; x += A2_THREADS)
				B_s[x] = 0.0;

		for (long long it = it0; it < it1; it++) {
			const long long i = it * A2_IT + ii;
			const long long ic = i < a.I ? i : a.I - 1;
			double e[KR], A[EM ? KR : 1];

			/* eta rows of this individual tile: registers for pass 1,
			 * shared tile for pass 2 */
#pragma unroll
			for (int k = 0; k < KR; k++)
				e[k] = k < a.K ? __ldg(a.eta + (size_t)ic * a.eta_stride + k) : 0.0;
			if (EM) {
#pragma unroll
				for (int k = 0; k < KR; k++)
					A[k] = 0.0;
				__syncthreads();	/* pass 2 of the previous tile is done */
				if (h == 0) {
#pragma unroll
					for (int kp = 0; kp < KP; kp++)
						*reinterpret_cast<double2 *>(eta_s + (size_t)ii * KR + 2 * kp)
							= make_double2(e[2 * kp], e[2 * kp + 1]);
				}
			}

			for (int lt = lt0; lt < lt1; lt++) {
				const int lf = a.lt_first[lt];
				const int lnext = lt + 1 < a.n_ltiles ? a.lt_first[lt + 1] : a.L;
				const int trow0 = a.off[lf];
				const int tile_rows = a.off[lnext] - trow0;
				const size_t tix = (size_t)it * a.n_ltiles + lt;

				__syncthreads();	/* w_s, p_s, csc_s free again */
				/* stage the tile's p rows: p_s[row][k] */
				for (int x = t; x < tile_rows * KR; x += A2_THREADS) {
					const int k = x / tile_rows, row = x % tile_rows;
					p_s[(size_t)row * KR + k] = k < a.K
						? __ldg(a.p + (size_t)k * a.T + trow0 + row) : 0.0;
				}
				if (t < LT)	/* first row of each locus inside the tile */
					rb_s[t] = lf + t < a.L ? a.off[lf + t] - trow0 : 0;
				if (EM) {
					/* sorted entry list + column starts -> shared */
					const unsigned short *cg = a.csc + tix * a.cap;
					const unsigned short *sg = a.colstart + tix * (a.ncolmax + 1);
					const int nent = sg[a.lt_ncol[lt]];
					for (int x = t * 8; x < nent; x += A2_THREADS * 8)
						a2_cp_async16(csc_s + x, cg + x);
					for (int x = t; x <= a.ncolmax; x += A2_THREADS)
						cst_s[x] = sg[x];
				}
				const uint2 cw = __ldg(reinterpret_cast<const uint2 *>(a.csr)
					+ tix * A2_THREADS + t);
				__syncthreads();

				/* ---- pass 1: tmp, w, A, log likelihood ---- */
#pragma unroll
				for (int q0 = 0; q0 < 8; q0 += 2) {
					double pr[2][KR], tmp[2];
					bool valid[2];
#pragma unroll
					for (int z = 0; z < 2; z++) {
						const int q = q0 + z;
						const unsigned code = ((q < 4 ? cw.x : cw.y) >> ((q & 3) * 8)) & 0xffu;
						valid[z] = code != 255u;
						const int row = valid[z] ? rb_s[h * LH + q / PP] + (int)code : 0;
						const double2 *src = reinterpret_cast<const double2 *>(p_s + (size_t)row * KR);
#pragma unroll
						for (int kp = 0; kp < KP; kp++) {
							const double2 v = src[kp];
							pr[z][2 * kp] = v.x;
							pr[z][2 * kp + 1] = v.y;
						}
					}
#pragma unroll
					for (int z = 0; z < 2; z++) {
						double s0 = 0.0, s1 = 0.0;
#pragma unroll
						for (int kp = 0; kp < KP; kp++) {
							s0 = fma(e[2 * kp], pr[z][2 * kp], s0);
							s1 = fma(e[2 * kp + 1], pr[z][2 * kp + 1], s1);
						}
						tmp[z] = valid[z] ? s0 + s1 : 1.0;
					}
					unsigned bad = 0;
#pragma unroll
					for (int z = 0; z < 2; z++)
						bad |= (unsigned)(__double2hiint(tmp[z]) - 0x00100000) >= 0x7fe00000u;
					if (!bad) {
						int es = 0;
#pragma unroll
						for (int z = 0; z < 2; z++) {
							const int hi = __double2hiint(tmp[z]);
							es += hi >> 20;
							prod *= __hiloint2double((hi & 0x000fffff) | 0x3ff00000,
								__double2loint(tmp[z]));
						}
						const int hi = __double2hiint(prod);
						es += (hi >> 20) - 3 * 1023;
						prod = __hiloint2double((hi & 0x000fffff) | 0x3ff00000,
							__double2loint(prod));
						esum += es;
					} else {
						ll_slow += log(tmp[0]) + log(tmp[1]);
					}
					if (EM) {
#pragma unroll
						for (int z = 0; z < 2; z++) {
							const int q = q0 + z;
							const double wgt = valid[z] ? mc_rcp(tmp[z]) : 0.0;
#pragma unroll
							for (int k = 0; k < KR; k++)
								A[k] = fma(pr[z][k], wgt, A[k]);
							w_s[(size_t)((h * LH + q / PP) * PP + q % PP) * A2_IT + ii] = wgt;
						}
					}
				}
				if (!EM)
					continue;
				a2_cp_async_wait();
				__syncthreads();

				/* ---- pass 2: G_lj += eta_i w over the column's entries ---- */
				const int S = a.lt_S[lt], ncol = a.lt_ncol[lt];
				const int col = t / S, seg = t % S;
				double g[KR];
				int ll2 = 0, j2 = 0;
#pragma unroll
				for (int k = 0; k < KR; k++)
					g[k] = 0.0;
				if (col < ncol) {
					const unsigned info = a.colinfo[(size_t)lt * a.ncolmax + col];
					const int ce = cst_s[col + 1];
					ll2 = info >> 8;
					j2 = info & 0xff;
					int x = cst_s[col] + seg;
					/* two entries per trip: their loads are issued together */
					for (; x + S < ce; x += 2 * S) {
						const unsigned e0 = csc_s[x], e1 = csc_s[x + S];
						const int i0 = e0 & 0xff, i1 = e1 & 0xff;
						const double w0 = w_s[(size_t)(ll2 * PP + ((e0 >> 8) & 0xf)) * A2_IT + i0]
							* (double)((e0 >> 12) + 1);
						const double w1 = w_s[(size_t)(ll2 * PP + ((e1 >> 8) & 0xf)) * A2_IT + i1]
							* (double)((e1 >> 12) + 1);
						const double2 *r0 = reinterpret_cast<const double2 *>(eta_s + (size_t)i0 * KR);
						const double2 *r1 = reinterpret_cast<const double2 *>(eta_s + (size_t)i1 * KR);
						double2 v0[KP], v1[KP];
#pragma unroll
						for (int kp = 0; kp < KP; kp++) {
							v0[kp] = r0[kp];
							v1[kp] = r1[kp];
						}
#pragma unroll
						for (int kp = 0; kp < KP; kp++) {
							g[2 * kp] = fma(v0[kp].x, w0, g[2 * kp]);
							g[2 * kp + 1] = fma(v0[kp].y, w0, g[2 * kp + 1]);
						}
#pragma unroll
						for (int kp = 0; kp < KP; kp++) {
							g[2 * kp] = fma(v1[kp].x, w1, g[2 * kp]);
							g[2 * kp + 1] = fma(v1[kp].y, w1, g[2 * kp + 1]);
						}
					}
					if (x < ce) {
						const unsigned ent = csc_s[x];
						const int ei = ent & 0xff, ea = (ent >> 8) & 0xf;
						const double wv = w_s[(size_t)(ll2 * PP + ea) * A2_IT + ei]
							* (double)((ent >> 12) + 1);
						const double2 *er = reinterpret_cast<const double2 *>(eta_s + (size_t)ei * KR);
#pragma unroll
						for (int kp = 0; kp < KP; kp++) {
							const double2 v = er[kp];
							g[2 * kp] = fma(v.x, wv, g[2 * kp]);
							g[2 * kp + 1] = fma(v.y, wv, g[2 * kp + 1]);
						}
					}
				}
				/* fold the S segments of a column (adjacent lanes; every lane
				 * of the warp takes part, idle ones carry zeros) */
#pragma unroll
				for (int m = 1; m < 32; m <<= 1)
					if (m < S) {
#pragma unroll
						for (int k = 0; k < KR; k++)
							g[k] += shfl_xor_f64(g[k], m);
					}
				if (col < ncol && seg == 0) {	/* the column's owner flushes */
					double2 *dst = reinterpret_cast<double2 *>(B_s
						+ (size_t)(trow0 - row0 + rb_s[ll2] + j2) * KR);
#pragma unroll
					for (int kp = 0; kp < KP; kp++) {
						double2 v = dst[kp];
						v.x += g[2 * kp];
						v.y += g[2 * kp + 1];
						dst[kp] = v;
					}
				}
			}
			if (EM) {
				/* A_i of this chunk of loci; the A2_H pass-1 threads of an
				 * individual own separate slots, k_admix_eta adds them */
				double *dst = a.Apart + ((size_t)(c * A2_H + h) * a.Ipad + i) * a.K;
#pragma unroll
				for (int k = 0; k < KR; k++)
					if (k < a.K)
						dst[k] = A[k];
			}
		}

		/* ---- flush the chunk's allele sums: N_klj = p_klj G_klj ---- */
		if (EM) {
			__syncthreads();
			double *Np = a.Npart + (size_t)r * a.K * a.T;
			for (int x = t; x < chunk_rows * a.K; x += A2_THREADS) {
				const int k = x / chunk_rows, row = x % chunk_rows;
				const size_t gx = (size_t)k * a.T + row0 + row;
				Np[gx] = B_s[(size_t)row * KR + k] * __ldg(a.p + gx);
			}
		}
		/* ---- log likelihood of the unit ---- */
		double ll = log(prod) + (double)esum * 0.693147180559945309417232121458 + ll_slow;
#pragma unroll
		for (int m = 16; m >= 1; m >>= 1)
			ll += shfl_xor_f64(ll, m);
		__syncthreads();
		if (lane == 0)
			red[t >> 5] = ll;
		__syncthreads();
		if (t == 0) {
			double sum = 0.0;
			for (int wv = 0; wv < A2_THREADS / 32; wv++)
				sum += red[wv];
			a.llpart[u] = sum;
		}
	}
}
