/*
 * mc_admix3.cuh -- two-pass admixture kernel, third mapping: A3_IT individuals
 * per tile (one per thread), gathers laid out for the minimum number of
 * shared-memory wavefronts.
 *
 * The fused E+M step of the admixture model (em_alg.c:325-433, 604-725) is,
 * per allele copy (i, l, a) with allele j:
 *     tmp = sum_k eta_ik p_klj,  w = c / tmp,  ll += c log tmp,
 *     A_ik += p_klj w   (-> D_ik = eta_ik A_ik,  the eta update)
 *     G_klj += eta_ik w (-> N_klj = p_klj G_klj, the p update)
 * Measured on B200 (tools/lds_probe3.cu, tools/shfl_probe.cu): the SM delivers
 * 128 B/clk from shared memory whatever the access width, multi-lane broadcast
 * does not make a 64-bit load cheaper, SHFL shares the same data path, and
 * FP64 issues 64 FMAs/clk.  Every copy needs one K-vector gathered per pass
 * (2 x 8K bytes against 3K FMAs), so the path is bound by the shared-memory
 * pipe and the kernel is built around wavefront counts:
 *   pass 1  thread <-> individual (eta_i, A_i in registers for the whole sweep
 *           over the CTA's loci); p is staged k-major, p_s[k][row], so the 32
 *           lanes of a warp -- 32 individuals at the SAME locus -- read one
 *           8-byte word each from <= J_l consecutive words: 2 wavefronts per
 *           LDS.64, the minimum.  w goes to shared memory, w_s[copy][i].
 *   pass 2  thread <-> (allele column, segment); a column gets a number of
 *           logical lanes proportional to its allele count, spread over the
 *           quarter warps (a3_thread_lane).  The eta rows live in shared
 *           memory with a pitch of an odd number of 16-byte pieces, so the rows
 *           of individuals with different i % 8 start in different bank groups
 *           and a quarter warp's LDS.128 costs as many wavefronts as the most
 *           frequent residue among its 8 entries.  Which entry a lane reads in
 *           which step is scheduled by the list builder (k3_build_csc, once per
 *           data set) so that the 8 lanes of a quarter warp mostly touch 8
 *           different bank groups.  The lists are stored step-major, two
 *           entries per 32-bit word and thread.
 *           (A first version padded the rows to 8 pieces and rotated the
 *           piece order per lane -- conflict-free for any assignment but 8
 *           loads for 5 useful pieces at K = 10.)
 *   fold    the lanes' partial sums go through a shared scratch; thread
 *           (column, piece) adds the column's partials in lane order and
 *           read-modify-writes the column's running sum in L2 (Gacc): no
 *           atomics, fixed order.
 * Two barriers per (A3_IT individuals x A3_NC copies) tile; two CTAs per SM so
 * that one's pass 1 (FP64) overlaps the other's pass 2 (shared-memory pipe).
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_device.cuh"

#ifndef A3_THREADS
#define A3_THREADS 256		/* measured at C3: 256 x 2 CTAs/SM 26.7 ms, 512 x 1 30.3 ms */
#endif
#define A3_IT A3_THREADS	/* individuals per tile = pass-1 threads */
#ifndef A3_NC
#define A3_NC 16		/* allele copies per individual and tile (8 or 16): pass 1
				 * runs A3_NC / 8 rounds of 8 copies, pass 2 walks lists of
				 * A3_IT * A3_NC entries, so the per-tile costs (two barriers,
				 * the fold, the staging) are paid once per A3_NC copies */
#endif
#define A3_WP (A3_IT + 8)	/* pitch of a copy's row of w in doubles: rows of copies of
				 * different parity start 8 bank pairs apart, so the two
				 * lanes of a half warp that share an eta bank group (i % 8)
				 * collide on w only when their copies have the same parity */
#define A3_PR 192		/* pitch of the k-major p tile in doubles: a compile-time
				 * constant, so the K loads of a copy share one address
				 * register (a tile has at most A3_PR allele rows) */
#ifndef A3_CTAS_PER_SM
#define A3_CTAS_PER_SM (512 / A3_THREADS)	/* CTAs sharing an SM (and its shared memory) */
#endif
#define A3_IDLE 1023u		/* lane map: no column */

struct Admix3Args {
	int K;
	int n_itiles, n_ltiles, n_lchunks, n_ichunks, n_units;
	long long I, Ipad, T;
	int L, ncolmax, max_chunk_rows;
	/* per locus tile (static) */
	const int *lt_ncol;		/* [n_ltiles] real allele columns */
	const unsigned short *colinfo;	/* [n_ltiles][ncolmax] locus_in_tile << 8 | allele */
	const int *lc_first;		/* [n_lchunks + 1] first locus tile of each chunk */
	const int *off;			/* [L + 1] prefix sums of J */
	const int *nat_of;		/* [T] row inside the kernel -> allele slot, or null: the rows
					 * of a locus with more than 16 slots are reordered (mc_cuda.cu);
					 * `p` then is the parameter slot in kernel row order
					 * (k3_permute_rows) and `p_nat` the slot itself */
	/* data */
	const unsigned char *codes;	/* [n_itiles][n_ltiles][A3_THREADS][A3_NC] */
	const unsigned short *csc;	/* [n_itiles][n_ltiles][cap / (2 A3_THREADS)][A3_THREADS][2]:
					 * entries 2j and 2j + 1 of thread t, i | first copy << 9 |
					 * (count - 1) << 12 */
	const unsigned short *colstart;	/* [n_itiles][n_ltiles][3 csw + A3_THREADS], csw =
					 * ncolmax + 1 rounded up to a multiple of 8: first entry
					 * (builder only), first logical lane, locus_in_tile << 8 | row
					 * of every column, then column | entries << 8 of every
					 * logical pass-2 lane (column 255: idle) */
	int cap;			/* 16-bit entries per tile: the longest lane list of any tile,
					 * rounded up to an even number, x A3_THREADS */
	/* parameters */
	const double *p, *eta, *p_nat;
	long long eta_stride;
	/* outputs */
	double *Apart;			/* [n_lchunks][Ipad][K] */
	double *Npart;			/* [n_ichunks][K*T] */
	double *Gacc;			/* [n_ichunks][T][2 KP]: the allele sums G_lj of every unit,
					 * read-modified-written in L2 by the fold (each row belongs
					 * to one CTA) */
	double *llpart;			/* [n_units] */
	/* MIX_E as the fall-back of the digit-sliced pass (mc_digit.cuh): run only
	 * when *run_if is set, and leave the chunk count for k_mix_tail */
	const int *run_if;
	int *n_chunks_dev;
};

/* ---------------------------------------------------------------------- */

__device__ __forceinline__ void a3_cp_async16(void *smem_dst, const void *gsrc)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gsrc));
}
__device__ __forceinline__ void a3_cp_async8(void *smem_dst, const void *gsrc)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(s), "l"(gsrc));
}
__device__ __forceinline__ void a3_cp_async_commit()
{
	asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void a3_cp_async_wait()
{
	asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

/* shared-memory loads from 32-bit addresses */
__device__ __forceinline__ double2 a3_lds_f64x2(unsigned addr)
{
	double2 v;
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
	return v;
}
__device__ __forceinline__ double a3_lds_f64(unsigned addr)
{
	double v;
	asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ unsigned a3_lds_u32(unsigned addr)
{
	unsigned v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

/* log-likelihood terms of sums that are zero, subnormal or not finite (never
 * seen in a healthy fit); out of line to keep pass 1 short */
static __device__ __noinline__ double a3_slow_ll(double t0, double t1)
{
	return log(t0) + log(t1);
}

/* Pass-2 lanes.  A column owns consecutive LOGICAL lanes; logical lane ln runs on thread
 * (ln % A3_NQ) * 8 + ln / A3_NQ, i.e. in quarter warp ln % A3_NQ: consecutive logical
 * lanes sit in different quarter warps, so the 8 lanes of a quarter warp read the lists
 * of 8 different columns (k3_build_csc, mc_admix3_build.cuh) */
#define A3_NQ (A3_THREADS / 8)
__host__ __device__ __forceinline__ int a3_lane_thread(int ln)
{
	return (ln % A3_NQ) * 8 + ln / A3_NQ;
}
__host__ __device__ __forceinline__ int a3_thread_lane(int t)
{
	return (t & 7) * A3_NQ + (t >> 3);
}
/* row of logical lane ln in the partial-sum scratch: consecutive lanes in consecutive
 * rows (the fold walks a column's lanes), one row skipped per A3_NQ lanes so that the 8
 * lanes of a quarter warp (ln = qw, A3_NQ + qw, ...) store into 8 different bank groups */
__host__ __device__ __forceinline__ int a3_part_row(int ln)
{
	return ln + ln / A3_NQ;
}
#define A3_PART_ROWS (A3_THREADS + 8)

/* 16-byte pieces per eta row in shared memory: an odd number, so that the rows
 * of 8 individuals with different i % 8 start in 8 different bank groups */
template <int KP> struct A3Row { static constexpr int NP = KP | 1; };

/* kernel modes */
enum { A3_ADMIX_EM = 0, A3_ADMIX_LL = 1, A3_MIX_E = 2, A3_MIX_M = 3 };

/* bytes of dynamic shared memory; the host planner uses the same formula */
static inline size_t a3_smem_bytes(int KP, int mode, int ncolmax, int cap)
{
	const int KR = 2 * KP, NP = KP | 1;
	const bool p1 = mode != A3_MIX_M, p2 = mode == A3_ADMIX_EM || mode == A3_MIX_M;
	const size_t csw = ((size_t)ncolmax + 1 + 7) / 8 * 8;
	size_t d = 16;
	if (p1)
		d += (size_t)KR * A3_PR;
	if (p2)
		d += (size_t)A3_IT * NP * 2 + (size_t)A3_PART_ROWS * KR;
	if (mode == A3_ADMIX_EM)
		d += (size_t)A3_NC * A3_WP;
	return d * sizeof(double)
		+ ((p2 ? (size_t)cap : 0) + 2 * (3 * csw + A3_THREADS)) * sizeof(unsigned short)
		+ 2 * 16 * sizeof(int);
}

/* L2-only accesses of the allele sums: a row is read-modified-written by one
 * CTA for a whole launch, but by different threads from tile to tile.  The rows
 * (K x T doubles per individual chunk, tens of MB) are marked evict-last and
 * the genotype codes / entry lists that stream through evict-first, so the
 * sums stay in the 126 MB L2 between two visits instead of going to HBM and
 * back once per tile of individuals. */
__device__ __forceinline__ unsigned long long a3_policy_keep()
{
	unsigned long long pol;
	asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ unsigned long long a3_policy_stream()
{
	unsigned long long pol;
	asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ double2 a3_ldcg2(const double *p, unsigned long long pol)
{
	double2 v;
	asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
		: "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
	return v;
}
__device__ __forceinline__ void a3_stcg2(double *p, double2 v, unsigned long long pol)
{
	asm volatile("st.global.cg.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;"
		:: "l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
/* (The entry lists are staged with plain cp.async: `cp.async ... L2::cache_hint` came
 * out of ptxas 12.9 with a clobbered descriptor register in some instantiations --
 * UMOV into the descriptor pair between its set-up and the LDGSTS -- and faulted as
 * an illegal instruction on B200.) */

/* MODE A3_ADMIX_EM: admixture E+M step (both passes); A3_ADMIX_LL: admixture log
 * likelihood (pass 1 only); A3_MIX_E: mixture E pass, a_ik = sum of log p over the
 * observed copies (pass 1 only: `p` is the log p table, no eta; em_alg.c:793-827,
 * log_likelihood.c:189-203); A3_MIX_M: mixture M pass, N_klj = sum_i v_ik c_ilj (pass 2
 * only: `eta` holds the v_ik rows, the weight of an entry is its copy count;
 * em_alg.c:965-986) */
template <int KP, int PP, int MODE>
__global__ void __launch_bounds__(A3_THREADS, A3_CTAS_PER_SM) admix3_kernel(const Admix3Args a)
{
	if (MODE == A3_MIX_E) {
		if (a.run_if && !*a.run_if)
			return;
		if (a.n_chunks_dev && blockIdx.x == 0 && threadIdx.x == 0)
			*a.n_chunks_dev = a.n_lchunks;
	}
	constexpr bool P1 = (MODE != A3_MIX_M);			/* pass 1 runs */
	constexpr bool P2 = (MODE == A3_ADMIX_EM || MODE == A3_MIX_M);	/* pass 2 + fold run */
	constexpr bool HAS_W = (MODE == A3_ADMIX_EM);
	constexpr bool HAS_A = (MODE == A3_ADMIX_EM || MODE == A3_MIX_E);
	constexpr bool HAS_TMP = (MODE == A3_ADMIX_EM || MODE == A3_ADMIX_LL);
	constexpr bool HAS_E = (MODE != A3_MIX_E);
	constexpr int KR = 2 * KP;
	constexpr int NP = A3Row<KP>::NP;
	constexpr int LT = A3_NC / PP;		/* loci per tile */
	constexpr int LH = 8 / PP;		/* loci per round of 8 copies */
	constexpr int NH = A3_NC / 8;		/* rounds of pass 1 */
	constexpr int PR = A3_PR;
	extern __shared__ __align__(128) double smem3d[];
	const int t = threadIdx.x, lane = t & 31;
	const int tl = a3_thread_lane(t);	/* logical pass-2 lane of this thread */
	const int csw = ((a.ncolmax + 1 + 7) / 8) * 8;	/* colstart row, 16-byte multiple */
	const int cstn = 3 * csw + A3_THREADS;	/* shorts of one tile's column tables */

	/* eta rows first: the xor rotation needs them aligned to their size */
	double *eta_s = smem3d;						/* [A3_IT][2 NP] */
	double *w_s = eta_s + (P2 ? (size_t)A3_IT * NP * 2 : 0);	/* [A3_NC][A3_WP] */
	double *part_s = w_s + (HAS_W ? (size_t)A3_NC * A3_WP : 0);	/* [A3_PART_ROWS][KR] */
	double *p_s = part_s + (P2 ? (size_t)A3_PART_ROWS * KR : 0);	/* [KR][PR] */
	double *red = p_s + (P1 ? (size_t)KR * PR : 0);		/* [16] */
	unsigned short *csc_s = reinterpret_cast<unsigned short *>(red + 16);	/* [cap] */
	unsigned short *cst2_s = csc_s + (P2 ? a.cap : 0);		/* [2][cstn]: first entry, first
									 * lane, locus/allele per column,
									 * column per lane; two tiles (the
									 * fold still reads one while the
									 * next one lands) */
	int *rb_s = reinterpret_cast<int *>(cst2_s + 2 * cstn);		/* [2][16] row bases */
	const unsigned eta_sa = (unsigned)__cvta_generic_to_shared(eta_s);
	const unsigned w_sa = (unsigned)__cvta_generic_to_shared(w_s);
	const unsigned csc_sa = (unsigned)__cvta_generic_to_shared(csc_s);
	const unsigned long long pol_keep = a3_policy_keep(), pol_stream = a3_policy_stream();

	for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
		const int c = u % a.n_lchunks, r = u / a.n_lchunks;
		const int lt0 = a.lc_first[c], lt1 = a.lc_first[c + 1];
		const long long it0 = (long long)a.n_itiles * r / a.n_ichunks;
		const long long it1 = (long long)a.n_itiles * (r + 1) / a.n_ichunks;
		const int l0 = lt0 * LT;
		const int lend = lt1 * LT < a.L ? lt1 * LT : a.L;
		const int row0 = a.off[l0];
		const int chunk_rows = a.off[lend] - row0;
		double *G_u = P2 ? a.Gacc + ((size_t)r * a.T + row0) * KR : nullptr;
		double prod = 1.0, ll_slow = 0.0;
		long long esum = 0;
		double e[KR], A[HAS_A ? KR : 1];

		__syncthreads();
		if (P2)
			for (int x = t; x < chunk_rows * KP; x += A3_THREADS)
				a3_stcg2(G_u + 2 * (size_t)x, make_double2(0.0, 0.0), pol_keep);

		/* what the NEXT tile needs, fetched one tile ahead: the thread's
		 * allele codes in registers, the p rows and the entry lists by cp.async */
		uint2 cw_n[NH];
		int tro_n = 0;	/* first allele row of the tile inside the chunk */
#pragma unroll
		for (int h = 0; h < NH; h++)
			cw_n[h] = make_uint2(0xffffffffu, 0xffffffffu);
		auto fetch_regs = [&](long long it, int lt) {
			const size_t tix = (size_t)it * a.n_ltiles + lt;
			const uint2 *src = reinterpret_cast<const uint2 *>(a.codes)
				+ (tix * A3_THREADS + t) * NH;
			if (!P1) {
				/* no pass 1: the codes are not read */
			} else if (NH == 2) {
				uint4 v;
				asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
					: "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src), "l"(pol_stream));
				cw_n[0] = make_uint2(v.x, v.y);
				cw_n[NH - 1] = make_uint2(v.z, v.w);
			} else {
				cw_n[0] = __ldg(src);
			}
			tro_n = __ldg(a.off + lt * LT) - row0;
		};
		auto stage_p = [&](int lt, int buf) {
			const int lf = lt * LT;
			const int lnext = lf + LT < a.L ? lf + LT : a.L;
			const int trow0 = a.off[lf];
			const int tile_rows = a.off[lnext] - trow0;
			for (int x = t; x < tile_rows * KR; x += A3_THREADS) {
				const int k = x / tile_rows, row = x - k * tile_rows;
				if (k < a.K)
					a3_cp_async8(p_s + (size_t)k * PR + row,
						a.p + (size_t)k * a.T + trow0 + row);
				else
					p_s[(size_t)k * PR + row] = 0.0;
			}
			if (t < LT)
				rb_s[buf * 16 + t] = lf + t < a.L ? a.off[lf + t] - trow0 : 0;
			a3_cp_async_commit();
		};
		auto stage_rb = [&](int lt, int buf) {	/* row bases when there is no p tile */
			const int lf = lt * LT;
			if (t < LT)
				rb_s[buf * 16 + t] = lf + t < a.L ? a.off[lf + t] - a.off[lf] : 0;
		};
		auto stage_lists = [&](long long it, int lt, int buf) {
			const size_t tix = (size_t)it * a.n_ltiles + lt;
			unsigned short *cst_s = cst2_s + buf * cstn;
			for (int x = t * 8; x < a.cap; x += A3_THREADS * 8)
				a3_cp_async16(csc_s + x, a.csc + tix * a.cap + x);
			if (t * 8 < cstn)
				a3_cp_async16(cst_s + t * 8, a.colstart + tix * cstn + t * 8);
			a3_cp_async_commit();
		};

		fetch_regs(it0, lt0);
		if (P1)
			stage_p(lt0, 0);
		else
			stage_rb(lt0, 0);
		if (P2) {
			stage_lists(it0, lt0, 0);
			/* the two-pass sweeps keep two barriers per tile (after pass 1, after
			 * pass 2); the first tile's p rows are waited for here */
			a3_cp_async_wait<0>();
			__syncthreads();
		}
		int buf = 0;

		for (long long it = it0; it < it1; it++) {
			const long long i = it * A3_IT + t;
			const long long ic = i < a.I ? i : a.I - 1;

			if (HAS_E) {
#pragma unroll
				for (int k = 0; k < KR; k++)
					e[k] = k < a.K ? __ldg(a.eta + (size_t)ic * a.eta_stride + k) : 0.0;
			}
			if (HAS_A) {
#pragma unroll
				for (int k = 0; k < KR; k++)
					A[k] = 0.0;
			}
			if (P2) {
				/* every reader of the previous tile's eta rows is past the
				 * barrier that follows pass 2 */
				double2 *dst = reinterpret_cast<double2 *>(eta_s + (size_t)t * NP * 2);
#pragma unroll
				for (int pc = 0; pc < KP; pc++)
					dst[pc] = make_double2(e[2 * pc], e[2 * pc + 1]);
			}

			for (int lt = lt0; lt < lt1; lt++, buf ^= 1) {
				uint2 cwa[NH];
#pragma unroll
				for (int h = 0; h < NH; h++)
					cwa[h] = cw_n[h];
				const int tro = tro_n;
				const bool last = lt + 1 == lt1;
				const long long itn = last ? it + 1 : it;
				const int ltn = last ? lt0 : lt + 1;
				const bool more = itn < it1;
				const int *rb = rb_s + buf * 16;
				const unsigned short *cst_s = cst2_s + buf * cstn;

				if (more)
					fetch_regs(itn, ltn);
				if (!P2) {
					a3_cp_async_wait<0>();
					__syncthreads();
				}
				/* E+M: this tile's p rows landed before the barrier that closed
				 * the previous tile's pass 2; its entry lists are only needed
				 * after pass 1 */

				/* ---- pass 1: tmp, w, A, log likelihood; rounds of 8 copies ---- */
#pragma unroll 1
				for (int hh = 0; P1 && hh < NH; hh++) {
				const uint2 cw = (NH == 2 && hh) ? cwa[NH - 1] : cwa[0];
				const int *rbh = rb + hh * LH;
				double pr[2][KR];
				bool valid[2];
				unsigned codes_q[2];
				auto load_row = [&](int q, int z) {
					const unsigned code = ((q < 4 ? cw.x : cw.y) >> ((q & 3) * 8)) & 0xffu;
					codes_q[z] = code;
					valid[z] = code != 255u;
					/* a missing copy reads the locus's first row: same
					 * bank window as the lanes that carry an allele */
					const int row = rbh[q / PP] + (valid[z] ? (int)code : 0);
#pragma unroll
					for (int k = 0; k < KR; k++)
						pr[z][k] = p_s[k * PR + row];
				};
				load_row(0, 0);
				double tprev = 1.0;
#pragma unroll
				for (int q = 0; q < 8; q++) {
					const int z = q & 1;
					if (q + 1 < 8)
						load_row(q + 1, z ^ 1);
					if (MODE == A3_MIX_E) {	/* a_ik += log p_klj of an observed copy */
#pragma unroll
						for (int k = 0; k < KR; k++)
							A[k] += valid[z] ? pr[z][k] : 0.0;
						continue;
					}
					double s0 = 0.0, s1 = 0.0;
#pragma unroll
					for (int kp = 0; kp < KP; kp++) {
						s0 = fma(e[2 * kp], pr[z][2 * kp], s0);
						s1 = fma(e[2 * kp + 1], pr[z][2 * kp + 1], s1);
					}
					const double tmp = valid[z] ? s0 + s1 : 1.0;
					if (MODE == A3_ADMIX_EM) {
						const double wgt = valid[z] ? mc_rcp(tmp) : 0.0;
#pragma unroll
						for (int k = 0; k < KR; k++)
							A[k] = fma(pr[z][k], wgt, A[k]);
						/* pass 2 meets an (individual, allele) pair once, at its
						 * first copy: that slot carries w times the number of
						 * copies of the allele at this locus */
						double wst = wgt;
						if (PP == 2) {
							if ((q & 1) == 0) {
								const unsigned nxt = ((q + 1 < 4 ? cw.x : cw.y)
									>> (((q + 1) & 3) * 8)) & 0xffu;
								wst = nxt == codes_q[z] ? wgt + wgt : wgt;
							}
						} else if (PP > 2) {
							int n = 1;
#pragma unroll
							for (int q2 = q + 1; q2 < 8; q2++)
								if (q2 / PP == q / PP)
									n += (((q2 < 4 ? cw.x : cw.y) >> ((q2 & 3) * 8)) & 0xffu)
										== codes_q[z];
							wst = wgt * (double)n;
						}
						w_s[(size_t)(hh * 8 + q) * A3_WP + t] = wst;
					}
					if (z == 0) {
						tprev = tmp;
					} else {
						/* log of a product of mantissas, exponents summed as
						 * integers: one log per unit instead of per copy */
						const int h0 = __double2hiint(tprev), h1 = __double2hiint(tmp);
						const unsigned bad = ((unsigned)(h0 - 0x00100000) >= 0x7fe00000u)
							| ((unsigned)(h1 - 0x00100000) >= 0x7fe00000u);
						if (!bad) {
							prod *= __hiloint2double((h0 & 0x000fffff) | 0x3ff00000,
								__double2loint(tprev));
							prod *= __hiloint2double((h1 & 0x000fffff) | 0x3ff00000,
								__double2loint(tmp));
							const int hp = __double2hiint(prod);
							esum += (h0 >> 20) + (h1 >> 20) + (hp >> 20) - 3 * 1023;
							prod = __hiloint2double((hp & 0x000fffff) | 0x3ff00000,
								__double2loint(prod));
						} else {
							ll_slow += a3_slow_ll(tprev, tmp);
						}
					}
				}
				}
				if (P2)
					a3_cp_async_wait<0>();
				__syncthreads();
				/* p_s is free: the next tile's p rows travel during pass 2 */
				if (more && P1)
					stage_p(ltn, buf ^ 1);
				else
					a3_cp_async_commit();	/* keeps the group count in step */
				if (!P2)
					continue;

				/* ---- pass 2: G_lj += eta_i w over the lane's entries ---- */
				double g[KR];
#pragma unroll
				for (int k = 0; k < KR; k++)
					g[k] = 0.0;
				/* the lane's column, from the tile's lane table.  A column's lanes
				 * are consecutive LOGICAL lanes; thread t is logical lane tl, so the
				 * 8 lanes of a quarter warp belong to 8 different columns (the list
				 * builder schedules their entries into different bank groups) */
				const unsigned lane_info = cst_s[3 * csw + tl];	/* column | entries << 8 */
				const int col = (int)(lane_info & 0xffu);
				if (col != 255) {
					const int nq = (int)(lane_info >> 8);	/* entries of this lane (>= 1) */
					const unsigned wb = w_sa + (unsigned)(cst_s[2 * csw + col] >> 8)
						* (PP * A3_WP * 8);
					/* entries 2j and 2j + 1 of thread t are the 32-bit word
					 * j * A3_THREADS + t of the tile's list: a warp reads 128
					 * consecutive bytes every other entry */
					unsigned x = csc_sa + 4u * (unsigned)t;
					auto w_addr = [&](unsigned en) {
						return wb + ((en >> 9) & 7) * (A3_WP * 8) + (en & (A3_IT - 1)) * 8;
					};
					/* mixture: the weight is the number of copies of the allele,
					 * made a double as 2^52 + n - 2^52 (an I2F.F64 here came out
					 * of ptxas 12.9 with its source inside the destination pair
					 * for even KP and faulted as an illegal instruction on B200) */
					auto copies = [](unsigned en) {
						return __hiloint2double(0x43300000, (int)((en >> 12) + 1))
							- 4503599627370496.0;
					};
					auto load_row = [&](unsigned en, double2 (&v)[KP]) {
						const unsigned rm = eta_sa + (en & (A3_IT - 1)) * (NP * 16);
#pragma unroll
						for (int s = 0; s < KP; s++)
							v[s] = a3_lds_f64x2(rm + (s << 4));
					};
					auto add_row = [&](const double2 (&v)[KP], double wc) {
#pragma unroll
						for (int s = 0; s < KP; s++) {
							g[2 * s] = fma(v[s].x, wc, g[2 * s]);
							g[2 * s + 1] = fma(v[s].y, wc, g[2 * s + 1]);
						}
					};
					/* ids run a pair ahead and weights one entry, and nothing is
					 * computed from a load in the trip that issues it: the warp
					 * issues in order, so an operation on a fresh load would hold
					 * the accumulation of the current entry back */
					unsigned pr = a3_lds_u32(x);
					unsigned prn = nq > 2 ? a3_lds_u32(x + 4 * A3_THREADS) : 0u;
					unsigned e0 = pr & 0xffffu;
					double w0 = HAS_W ? a3_lds_f64(w_addr(e0)) : 0.0;
					for (int j = 0; j < nq; j += 2) {
						const unsigned e1 = pr >> 16;
						double2 v[KP];
						double wc = HAS_W ? w0 : copies(e0);
						load_row(e0, v);
						if (HAS_W)
							w0 = j + 1 < nq ? a3_lds_f64(w_addr(e1)) : 0.0;
						add_row(v, wc);
						if (j + 1 >= nq)
							break;
						pr = prn;
						x += 4 * A3_THREADS;
						prn = j + 4 < nq ? a3_lds_u32(x + 4 * A3_THREADS) : 0u;
						e0 = pr & 0xffffu;
						wc = HAS_W ? w0 : copies(e1);
						load_row(e1, v);
						if (HAS_W)
							w0 = j + 2 < nq ? a3_lds_f64(w_addr(e0)) : 0.0;
						add_row(v, wc);
					}
				}
				/* the lane's partial sums */
#pragma unroll
				for (int pc = 0; pc < KP; pc++)
					*reinterpret_cast<double2 *>(part_s + (size_t)a3_part_row(tl) * KR + 2 * pc)
						= make_double2(g[2 * pc], g[2 * pc + 1]);
				/* ---- fold: thread <-> (column, piece), lanes in order; the
				 * column's running sum lives in L2.  The sums of the thread's first
				 * two (column, piece) items are requested before the barrier, so the
				 * L2 round trip overlaps the wait for the other warps' pass 2 ---- */
				auto fold_dst = [&](int f) -> double * {
					if (f >= a.ncolmax * KP)
						return nullptr;
					const int cc = f / KP, pc = f - cc * KP;
					if ((int)cst_s[csw + cc + 1] - (int)cst_s[csw + cc] <= 0)
						return nullptr;
					const unsigned info = cst_s[2 * csw + cc];
					return G_u + (size_t)(tro + rb[info >> 8] + (int)(info & 0xff)) * KR + 2 * pc;
				};
				double *dst0 = fold_dst(t), *dst1 = fold_dst(t + A3_THREADS);
				double2 old0 = make_double2(0.0, 0.0), old1 = old0;
				if (dst0)
					old0 = a3_ldcg2(dst0, pol_keep);
				if (dst1)
					old1 = a3_ldcg2(dst1, pol_keep);
				a3_cp_async_wait<0>();	/* the next tile's p rows */
				__syncthreads();
				if (more) {
					if (!P1)
						stage_rb(ltn, buf ^ 1);
					stage_lists(itn, ltn, buf ^ 1);
				} else {
					a3_cp_async_commit();
				}

				auto fold_item = [&](int f, double *dst, double2 old) {
					const int cc = f / KP, pc = f - cc * KP;
					const int lane0 = cst_s[csw + cc], S = cst_s[csw + cc + 1] - lane0;
					const double2 *src = reinterpret_cast<const double2 *>(part_s + 2 * pc);
					auto part = [&](int sx) {
						return src[(size_t)a3_part_row(lane0 + sx) * KP];
					};
					/* four independent partial sums keep four loads in flight */
					double2 acc = make_double2(0.0, 0.0), a1 = acc, a2 = acc, a3 = acc;
					int sx = 0;
					for (; sx + 3 < S; sx += 4) {
						const double2 v0 = part(sx);
						const double2 v1 = part(sx + 1);
						const double2 v2 = part(sx + 2);
						const double2 v3 = part(sx + 3);
						acc.x += v0.x; acc.y += v0.y;
						a1.x += v1.x; a1.y += v1.y;
						a2.x += v2.x; a2.y += v2.y;
						a3.x += v3.x; a3.y += v3.y;
					}
					for (; sx < S; sx++) {
						const double2 v0 = part(sx);
						acc.x += v0.x;
						acc.y += v0.y;
					}
					double2 v = old;
					v.x += (acc.x + a1.x) + (a2.x + a3.x);
					v.y += (acc.y + a1.y) + (a2.y + a3.y);
					a3_stcg2(dst, v, pol_keep);
				};
				if (dst0)
					fold_item(t, dst0, old0);
				if (dst1)
					fold_item(t + A3_THREADS, dst1, old1);
				for (int f = t + 2 * A3_THREADS; f < a.ncolmax * KP; f += A3_THREADS) {
					double *dst = fold_dst(f);
					if (dst)
						fold_item(f, dst, a3_ldcg2(dst, pol_keep));
				}
			}
			if (HAS_A) {
				double *dst = a.Apart + ((size_t)c * a.Ipad + i) * a.K;
#pragma unroll
				for (int k = 0; k < KR; k++)
					if (k < a.K)
						dst[k] = A[k];
			}
		}
		a3_cp_async_wait<0>();

		/* ---- flush the chunk's allele sums: N_klj = p_klj G_klj (the mixture's
		 * sums are the counts themselves) ---- */
		if (P2) {
			__syncthreads();
			double *Np = a.Npart + (size_t)r * a.K * a.T;
			for (int x = t; x < chunk_rows * a.K; x += A3_THREADS) {
				const int k = x / chunk_rows, row = x - k * chunk_rows;
				const size_t gx = (size_t)k * a.T
					+ (a.nat_of ? __ldg(a.nat_of + row0 + row) : row0 + row);
				const double gv = __ldcg(G_u + (size_t)row * KR + k);
				Np[gx] = MODE == A3_ADMIX_EM ? gv * __ldg(a.p_nat + gx) : gv;
			}
		}
		/* ---- log likelihood of the unit ---- */
		if (!HAS_TMP)
			continue;
		double ll = log(prod) + (double)esum * 0.693147180559945309417232121458 + ll_slow;
#pragma unroll
		for (int mm = 16; mm >= 1; mm >>= 1)
			ll += shfl_xor_f64(ll, mm);
		__syncthreads();
		if (lane == 0)
			red[t >> 5] = ll;
		__syncthreads();
		if (t == 0) {
			double sum = 0.0;
			for (int wv = 0; wv < A3_THREADS / 32; wv++)
				sum += red[wv];
			a.llpart[u] = sum;
		}
	}
}
