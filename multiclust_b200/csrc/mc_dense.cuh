/*
 * mc_dense.cuh -- dense FP64 tensor-core (DMMA) kernels for biallelic data.
 *
 * When no locus has more than two observed alleles (BASELINE configs 2 and 5)
 * every (individual, locus) pair touches both allele columns, so the E- and
 * M-step sums are dense K-inner-dimension contractions and the gather design
 * of mc_admix3.cuh buys nothing.  With c_ila the number of copies of allele a:
 *   admixture (em_alg.c:325-433, 604-725; log_likelihood.c:128-144)
 *     tmp_ila = sum_k eta_ik p_kla            tmp = eta . p        [I x K][K x 2L]
 *     w_ila   = c_ila / tmp_ila,  ll += c_ila log tmp_ila
 *     A_ik   += sum_{l,a} w_ila p_kla         A   = W . p^T        [I x 2L][2L x K]
 *     G_kla  += sum_i eta_ik w_ila            G   = eta^T . W      [K x I][I x 2L]
 *   mixture (em_alg.c:793-827, 965-986; log_likelihood.c:189-203)
 *     a_ik    = sum_{l,a} c_ila log p_kla     a   = C . (log p)^T  (the A product with W = C)
 *     N_kla   = sum_i v_ik c_ila              N   = v^T . C        (the G product with W = C)
 * All three products run on mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  Measured on
 * B200 (tools/dmma_probe.cu, profiles/r02_dmma_probe.txt): DMMA issues 64
 * FMA/clk/SM, exactly the DFMA rate, and the two share one pipe (4 DMMA + 8 DFMA
 * interleaved: 60 FMA/clk), so the tensor path does not raise the FP64 ceiling --
 * what it buys is operand sharing across lanes: one LDS.64 feeds 256 FMAs where
 * the SIMT form needs a shared-memory broadcast per 32, and an eighth of the
 * issue slots.  The kernel is bound by the FP64 pipe.
 *
 * Mapping.  A CTA tile is 256 individuals x 16 loci; warp w owns the tile's
 * individuals 32w..32w+31 as four groups of 8.  One step handles (8 individuals)
 * x (4 loci = 8 allele columns); with r = lane / 4, q = lane % 4:
 *   tmp   C[i = r][col = 2q + a]  = sum_kb  A{eta[i=r][k=4kb+q]} . B{p[k=4kb+q][col=r]}
 *         -> the lane holds tmp of individual r at locus q for both alleles
 *   A     C[i = r][k = 2q + e]   += A{w_a[i=r][locus q]} . B{p_a[locus q][k=r]},  a = 0, 1
 *         (the contracted index is the locus, so the tmp fragment is used as it is)
 *   G     C[k = r][col = 2q + a] += A{eta[k=r][i=4h+q]} . B{w[i=4h+q][col=r]},  h = 0, 1
 *         (the contracted index is the individual: the 8 x 8 block of w goes
 *          through a 512-byte per-warp scratch, swizzled so that the 128-bit
 *          stores and the 64-bit fragment loads are both conflict free)
 * A_ik stays in the accumulator fragments while the CTA sweeps its locus chunk;
 * the G fragments of the 8 warps are added in warp order into the CTA's
 * shared accumulator once per tile: no atomics, fixed order.
 *
 * Genotypes are stored as one byte per (individual, locus): c_0 | c_1 << 4, in
 * tile-major 16-byte rows -- a quarter of the natural codes at ploidy 4.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_device.cuh"

#define DN_THREADS 256
#define DN_IT 256		/* individuals per tile */
#define DN_TL 16		/* loci per tile */
#define DN_CTAS_PER_SM 2

enum { DN_ADMIX_EM = 0, DN_ADMIX_LL = 1, DN_MIX_E = 2, DN_MIX_M = 3 };

/* doubles per locus of the dense p table: [k < 8 NB][a < 2] + 8 of padding
 * (pitch = 8 mod 16: the B fragments of a half warp fall into 16 bank pairs) */
static inline __host__ __device__ int dn_pl(int NB) { return 16 * NB + 8; }

struct DenseArgs {
	int K, L;
	int n_itiles, n_ltiles, n_lchunks, n_ichunks, n_units;
	int max_chunk_tiles;
	long long I, Ipad, T;
	const int *lc_first;		/* [n_lchunks + 1] first locus tile of each chunk */
	const int *off, *J;		/* [L + 1], [L] */
	const unsigned char *cnt;	/* [n_itiles][n_ltiles][256][16] c0 | c1 << 4 */
	const double *pd;		/* [n_ltiles * 16][PL] dense p or log p */
	const double *p;		/* [K][T]: N_kla = p_kla G_kla at the flush */
	const double *eta;		/* eta rows (admixture) / v rows (mixture M-step) */
	long long eta_stride;
	double *Apart;			/* [n_lchunks][Ipad][K] */
	double *Npart;			/* [n_ichunks][K * T] */
	double *llpart;			/* [n_units] */
	/* fall-back of the digit-sliced mixture pass (mc_digit.cuh): run only when
	 * *run_if is set, and leave the chunk count for k_mix_tail */
	const int *run_if;
	int *n_chunks_dev;
};

/* ---------------------------------------------------------------------- */

__device__ __forceinline__ void dn_mma(double &c0, double &c1, double a, double b)
{
	/* not volatile: the compiler may interleave independent steps */
	asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
		: "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void dn_cp_async16(void *smem_dst, const void *gsrc)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gsrc));
}

/* log-likelihood terms the mantissa product cannot take (never seen in a
 * healthy fit); out of line: sixteen inlined copies of log() per tile would
 * push the tile loop out of the instruction cache */
static __device__ __noinline__ double dn_slow_ll(double t0, double t1, unsigned c0, unsigned c1)
{
	double s = 0.0;
	if (c0)
		s += (double)c0 * log(t0);
	if (c1)
		s += (double)c1 * log(t1);
	return s;
}

/* m^c for a mantissa m in [1, 2) and a count c <= PMAX (the ploidy) */
template <int PMAX> __device__ __forceinline__ double dn_pow(double m, unsigned c)
{
	double r = (c & 1u) ? m : 1.0;
	if (PMAX >= 2) {
		const double m2 = m * m;
		r = (c & 2u) ? r * m2 : r;
		if (PMAX >= 4) {
			const double m4 = m2 * m2;
			if (PMAX == 4) {	/* c = 4 is the only count with bit 2 set */
				r = (c & 4u) ? m4 : r;
			} else {
				r = (c & 4u) ? r * m4 : r;
				if (PMAX >= 8) {
					const double m8 = m4 * m4;
					r = (c & 8u) ? r * m8 : r;
				}
			}
		}
	}
	return r;
}

/* a count 0..15 as a double.  The conversion instruction runs on the
 * conversion pipe, which is idle here; the integer-bits-plus-DADD form would
 * put it on the FP64 pipe, the kernel's bottleneck (measured: +9 % time). */
__device__ __forceinline__ double dn_count(unsigned c)
{
	return (double)c;
}

/* bytes of dynamic shared memory; the host planner uses the same formula */
static inline size_t dn_smem_bytes(int NB, int mode, int max_chunk_tiles)
{
	const bool has_g = mode == DN_ADMIX_EM || mode == DN_MIX_M;
	const bool has_p = mode != DN_MIX_M;
	size_t d = 16;
	if (has_g)
		d += (size_t)max_chunk_tiles * NB * 256 + (size_t)8 * NB * 256;
	if (mode == DN_ADMIX_EM)
		d += 8 * 4 * 64;
	if (has_p)
		d += (size_t)2 * DN_TL * dn_pl(NB);
	return d * sizeof(double) + (size_t)2 * DN_IT * 16;
}

template <int NB, int PMAX, int MODE>
__global__ void __launch_bounds__(DN_THREADS, DN_CTAS_PER_SM) dense_kernel(const DenseArgs a)
{
	constexpr int KB = 2 * NB;		/* inner chunks of 4 clusters */
	constexpr int PL = 16 * NB + 8;
	constexpr bool HAS_G = (MODE == DN_ADMIX_EM || MODE == DN_MIX_M);
	constexpr bool HAS_A = (MODE == DN_ADMIX_EM || MODE == DN_MIX_E);
	constexpr bool HAS_P = (MODE != DN_MIX_M);
	constexpr bool HAS_TMP = (MODE == DN_ADMIX_EM || MODE == DN_ADMIX_LL);
	extern __shared__ __align__(128) double dsm[];
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	const int r = lane >> 2, q = lane & 3;
	if (MODE == DN_MIX_E) {
		if (a.run_if && !*a.run_if)
			return;
		if (a.n_chunks_dev && blockIdx.x == 0 && t == 0)
			*a.n_chunks_dev = a.n_lchunks;
	}

	double *B_s = dsm;								/* [chunk tiles][NB][256] */
	double *scr = B_s + (HAS_G ? (size_t)a.max_chunk_tiles * NB * 256 : 0);	/* [8][NB][4][32][2] */
	double *ws = scr + (HAS_G ? 8 * NB * 256 : 0);					/* [8][4][64] */
	double *pt = ws + (MODE == DN_ADMIX_EM ? 8 * 4 * 64 : 0);			/* [2][16][PL] */
	double *red = pt + (HAS_P ? 2 * DN_TL * PL : 0);				/* [16] */
	unsigned char *ct = reinterpret_cast<unsigned char *>(red + 16);		/* [2][256][16] */
	double *ws_w = ws + w * 256;

	for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
		const int c = u % a.n_lchunks, rr = u / a.n_lchunks;
		const int lt0 = a.lc_first[c], lt1 = a.lc_first[c + 1];
		const long long it0 = (long long)a.n_itiles * rr / a.n_ichunks;
		const long long it1 = (long long)a.n_itiles * (rr + 1) / a.n_ichunks;
		double prod[4] = { 1.0, 1.0, 1.0, 1.0 }, ll_slow = 0.0;	/* one chain per group */
		long long esum = 0;

		__syncthreads();	/* the previous unit has left shared memory */
		if (HAS_G)
			for (int x = t; x < (lt1 - lt0) * NB * 256; x += DN_THREADS)
				B_s[x] = 0.0;

		auto stage = [&](long long it, int lt, int buf) {
			const size_t tix = (size_t)it * a.n_ltiles + lt;
			dn_cp_async16(ct + ((size_t)buf * DN_IT + t) * 16, a.cnt + (tix * DN_IT + t) * 16);
			if (HAS_P)
				for (int x = t; x < DN_TL * PL / 2; x += DN_THREADS)
					dn_cp_async16(pt + (size_t)buf * DN_TL * PL + 2 * x,
						a.pd + (size_t)lt * DN_TL * PL + 2 * x);
			asm volatile("cp.async.commit_group;" ::: "memory");
		};
		if (it0 < it1 && lt0 < lt1)
			stage(it0, lt0, 0);
		int buf = 0;

		for (long long it = it0; it < it1; it++) {
			const long long ib = it * DN_IT + w * 32;
			/* eta (or v) fragments of the warp's 32 individuals */
			double Ae[HAS_TMP ? 4 : 1][KB], AeT[HAS_G ? 4 : 1][2][NB];
			if (HAS_TMP) {
#pragma unroll
				for (int g = 0; g < 4; g++)
#pragma unroll
					for (int kb = 0; kb < KB; kb++) {
						const long long i = ib + 8 * g + r;
						const int k = 4 * kb + q;
						Ae[g][kb] = (i < a.I && k < a.K)
							? __ldg(a.eta + (size_t)i * a.eta_stride + k) : 0.0;
					}
			}
			if (HAS_G) {
#pragma unroll
				for (int g = 0; g < 4; g++)
#pragma unroll
					for (int h = 0; h < 2; h++)
#pragma unroll
						for (int nb = 0; nb < NB; nb++) {
							const long long i = ib + 8 * g + 4 * h + q;
							const int k = 8 * nb + r;
							AeT[g][h][nb] = (i < a.I && k < a.K)
								? __ldg(a.eta + (size_t)i * a.eta_stride + k) : 0.0;
						}
			}
			double CA[HAS_A ? 4 : 1][NB][2];
			if (HAS_A) {
#pragma unroll
				for (int g = 0; g < 4; g++)
#pragma unroll
					for (int nb = 0; nb < NB; nb++)
						CA[g][nb][0] = CA[g][nb][1] = 0.0;
			}

			for (int lt = lt0; lt < lt1; lt++, buf ^= 1) {
				const bool last = lt + 1 == lt1;
				const long long itn = last ? it + 1 : it;
				const int ltn = last ? lt0 : lt + 1;
				/* this tile has landed; behind the barrier everybody is past the
				 * previous tile's reads of buf ^ 1, which the next tile may now
				 * overwrite while this one is computed */
				asm volatile("cp.async.wait_group 0;" ::: "memory");
				__syncthreads();
				if (itn < it1)
					stage(itn, ltn, buf ^ 1);

				const unsigned char *cts = ct + (size_t)buf * DN_IT * 16;
				const double *pts = pt + (size_t)buf * DN_TL * PL;
				/* One locus quad at a time; inside it every phase runs over the
				 * four groups of 8 individuals, so four independent dependency
				 * chains are in flight (the warp issues in order). */
#pragma unroll
				for (int lq = 0; lq < 4; lq++) {
					double G[HAS_G ? 2 : 1][NB][2];	/* two chains: h = 0, 1 */
					if (HAS_G) {
#pragma unroll
						for (int h = 0; h < 2; h++)
#pragma unroll
							for (int nb = 0; nb < NB; nb++)
								G[h][nb][0] = G[h][nb][1] = 0.0;
					}
					if (MODE == DN_MIX_M) {
						/* the B fragment straight from the counts: individual
						 * 4h + q of the group, column r = locus r / 2, allele r % 2 */
#pragma unroll
						for (int g = 0; g < 4; g++)
#pragma unroll
							for (int h = 0; h < 2; h++) {
								const unsigned cb = cts[(w * 32 + 8 * g + 4 * h + q) * 16
									+ 4 * lq + (r >> 1)];
								const double wv = dn_count((cb >> ((r & 1) * 4)) & 15u);
#pragma unroll
								for (int nb = 0; nb < NB; nb++)
									dn_mma(G[h][nb][0], G[h][nb][1], AeT[g][h][nb], wv);
							}
					} else {
						/* p fragments of the locus quad, shared by the four groups */
						double Bp[HAS_TMP ? KB : 1], BpT[HAS_A ? 2 : 1][NB];
						if (HAS_TMP) {
#pragma unroll
							for (int kb = 0; kb < KB; kb++)
								Bp[kb] = pts[(4 * lq + (r >> 1)) * PL + 2 * (4 * kb + q) + (r & 1)];
						}
						if (HAS_A) {
#pragma unroll
							for (int al = 0; al < 2; al++)
#pragma unroll
								for (int nb = 0; nb < NB; nb++)
									BpT[al][nb] = pts[(4 * lq + q) * PL + 2 * (8 * nb + r) + al];
						}
						/* group after group: the DMMAs of one group (16 clocks of the
						 * pipe each) are spaced by its own element-wise work, which
						 * keeps the in-order warp issuing while the pipe is busy */
#pragma unroll
						for (int g = 0; g < 4; g++) {
							const unsigned cb = cts[(w * 32 + 8 * g + r) * 16 + 4 * lq + q];
							const unsigned c0 = cb & 15u, c1 = cb >> 4;
							double w0, w1;
							if (MODE == DN_MIX_E) {
								w0 = dn_count(c0);
								w1 = dn_count(c1);
							} else {
								double t0 = 0.0, t1 = 0.0;
#pragma unroll
								for (int kb = 0; kb < KB; kb++)
									dn_mma(t0, t1, Ae[g][kb], Bp[kb]);
								/* log likelihood: mantissas multiplied up, exponents
								 * summed as integers, one log per unit */
								const int h0 = __double2hiint(t0), h1 = __double2hiint(t1);
								const bool ok0 = (unsigned)(h0 - 0x00100000) < 0x7fe00000u;
								const bool ok1 = (unsigned)(h1 - 0x00100000) < 0x7fe00000u;
								const unsigned f0 = ok0 ? c0 : 0u, f1 = ok1 ? c1 : 0u;
								const double m0 = __hiloint2double((h0 & 0x000fffff) | 0x3ff00000,
									__double2loint(t0));
								const double m1 = __hiloint2double((h1 & 0x000fffff) | 0x3ff00000,
									__double2loint(t1));
								prod[g] *= dn_pow<PMAX>(m0, f0) * dn_pow<PMAX>(m1, f1);
								const int hp = __double2hiint(prod[g]);
								esum += (int)f0 * ((h0 >> 20) - 1023) + (int)f1 * ((h1 >> 20) - 1023)
									+ ((hp >> 20) - 1023);
								prod[g] = __hiloint2double((hp & 0x000fffff) | 0x3ff00000,
									__double2loint(prod[g]));
								if ((c0 && !ok0) || (c1 && !ok1))	/* zero, subnormal, inf, nan */
									ll_slow += dn_slow_ll(t0, t1, ok0 ? 0u : c0, ok1 ? 0u : c1);
								if (MODE == DN_ADMIX_LL)
									continue;
								w0 = c0 ? dn_count(c0) * mc_rcp(t0) : 0.0;
								w1 = c1 ? dn_count(c1) * mc_rcp(t1) : 0.0;
							}
							/* A_ik += sum over the quad's loci of w_a p_a */
#pragma unroll
							for (int nb = 0; nb < NB; nb++) {
								dn_mma(CA[g][nb][0], CA[g][nb][1], w0, BpT[0][nb]);
								dn_mma(CA[g][nb][0], CA[g][nb][1], w1, BpT[1][nb]);
							}
							if (MODE == DN_ADMIX_EM)
								/* the 8 x 8 block of w is transposed through the warp's
								 * scratch: element (row, col) at row * 8 + (col ^ (row & 2 ? 4 : 0)) */
								*reinterpret_cast<double2 *>(ws_w + g * 64 + r * 8
									+ ((2 * q) ^ ((r & 2) << 1))) = make_double2(w0, w1);
						}
						if (MODE == DN_ADMIX_EM) {
							__syncwarp();
#pragma unroll
							for (int g = 0; g < 4; g++)
#pragma unroll
								for (int h = 0; h < 2; h++) {
									const int row = 4 * h + q;
									const double wv = ws_w[g * 64 + row * 8 + (r ^ ((row & 2) << 1))];
#pragma unroll
									for (int nb = 0; nb < NB; nb++)
										dn_mma(G[h][nb][0], G[h][nb][1], AeT[g][h][nb], wv);
								}
							__syncwarp();	/* the next quad overwrites the scratch */
						}
					}
					/* the warp's G fragment of this quad goes to the fold scratch */
					if (HAS_G) {
#pragma unroll
						for (int nb = 0; nb < NB; nb++)
							*reinterpret_cast<double2 *>(scr + (((size_t)w * NB + nb) * 4 + lq) * 64
								+ lane * 2) = make_double2(G[0][nb][0] + G[1][nb][0],
								G[0][nb][1] + G[1][nb][1]);
					}
				}
				if (!HAS_G)
					continue;
				/* ---- the warps' G fragments, added in warp order ---- */
				__syncthreads();
#pragma unroll
				for (int nb = 0; nb < NB; nb++) {
					double s = 0.0;
#pragma unroll
					for (int ww = 0; ww < 8; ww++)
						s += scr[((size_t)ww * NB + nb) * 256 + t];
					B_s[((size_t)(lt - lt0) * NB + nb) * 256 + t] += s;
				}
			}
			if (HAS_A) {
#pragma unroll
				for (int g = 0; g < 4; g++)
#pragma unroll
					for (int nb = 0; nb < NB; nb++)
#pragma unroll
						for (int e = 0; e < 2; e++) {
							const long long i = ib + 8 * g + r;
							const int k = 8 * nb + 2 * q + e;
							if (k < a.K)
								a.Apart[((size_t)c * a.Ipad + i) * a.K + k] = CA[g][nb][e];
						}
			}
		}
		asm volatile("cp.async.wait_group 0;" ::: "memory");

		/* ---- flush the chunk's allele sums ---- */
		if (HAS_G) {
			__syncthreads();
			double *Np = a.Npart + (size_t)rr * a.K * a.T;
			for (int x = t; x < (lt1 - lt0) * NB * 256; x += DN_THREADS) {
				const int tau = x / (NB * 256), rem = x - tau * NB * 256;
				const int nb = rem >> 8, tt = rem & 255;
				const int lq = tt >> 6, ln = (tt & 63) >> 1, e = tt & 1;
				const int l = (lt0 + tau) * DN_TL + 4 * lq + (ln & 3);
				const int k = 8 * nb + (ln >> 2);
				if (l < a.L && k < a.K && e < a.J[l]) {
					const size_t gx = (size_t)k * a.T + a.off[l] + e;
					double v = B_s[x];
					if (MODE == DN_ADMIX_EM)
						v *= __ldg(a.p + gx);
					Np[gx] = v;
					/* the phantom slot: no copy carries it (the mixture's
					 * finish step reuses Npart, so it is rewritten every time) */
					if (e == 0 && a.J[l] > 2)
						Np[gx + 2] = 0.0;
				}
			}
		}
		/* ---- log likelihood of the unit ---- */
		if (HAS_TMP) {
			double ll = log(prod[0]) + log(prod[1]) + log(prod[2]) + log(prod[3])
				+ (double)esum * 0.693147180559945309417232121458 + ll_slow;
#pragma unroll
			for (int mm = 16; mm >= 1; mm >>= 1)
				ll += shfl_xor_f64(ll, mm);
			__syncthreads();
			if (lane == 0)
				red[w] = ll;
			__syncthreads();
			if (t == 0) {
				double sum = 0.0;
				for (int wv = 0; wv < DN_THREADS / 32; wv++)
					sum += red[wv];
				a.llpart[u] = sum;
			}
		}
	}
}

