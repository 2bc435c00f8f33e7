/*
 * mc_tile.cuh -- the round-1 one-pass genotype-streaming kernel (multi-allelic mixture,
 * admixture with K > 16 or ploidy > 8); instantiated in mc_inst_tile.cu.
 *
 * One templated genotype-streaming kernel (`tile_kernel`) covers the four
 * data passes of the reference, selected by MODE:
 *   MODE_ADMIX_EM  e_step_admixture_orig + the sums m_step_admixture_orig
 *                  needs (em_alg.c:325-433, 604-725), fused, never
 *                  materialising d_iklm (multiclust.c:1197 allocates I*K*T)
 *   MODE_ADMIX_LL  logL_admixture (log_likelihood.c:128-144)
 *   MODE_MIX_E     per-individual sum of log p over the observed copies
 *                  (e_step_mixture em_alg.c:793-827, logL_mixture
 *                  log_likelihood.c:189-203)
 *   MODE_MIX_M     allele-count sums of m_step_mixture (em_alg.c:965-986)
 *
 * Mapping:
 *   - loci are dealt to tiles; a CTA owns one tile for one chunk of
 *     individuals ("unit"), units are striped over a persistent grid;
 *   - inside a warp, lane = locus_in_warp * k_split + kh: k_split adjacent
 *     lanes share a locus and each holds KH = ceil(K / k_split) of the K
 *     clusters in registers;
 *   - the tile's p rows and allele-count accumulators live in shared memory in
 *     a lane-interleaved layout [row][kk][lane], so every 64-bit access of a
 *     half warp hits 16 distinct bank pairs whatever the allele codes are, and
 *     each accumulator column is private to one thread: no atomics, and the
 *     summation order is fixed (deterministic);
 *   - genotypes are read straight from HBM as one 16-byte vector per (locus,
 *     block of individuals) from a tile-major layout, so a warp reads a
 *     contiguous run of 16-byte units; they are not staged through shared
 *     memory because the shared-memory pipe is this kernel's binding resource;
 *   - per-individual sums (A_ik) are kept in registers for a block of
 *     individuals, folded across the lanes of a warp by a recursive-halving
 *     shuffle reduction and across warps through a small scratch buffer.
 */
#pragma once

#include "mc_device.cuh"

enum { MODE_ADMIX_EM = 0, MODE_ADMIX_LL = 1, MODE_MIX_E = 2, MODE_MIX_M = 3 };

struct TileArgs {
	/* plan */
	int K, k_split, loci_per_warp, warps, groups;
	int n_tiles, n_chunks, n_units;
	int tile_slots;			/* warps * groups * loci_per_warp */
	long long n_blocks;		/* blocks of IB individuals */
	long long I, Ipad, T;
	/* per tile tables */
	const int *slot_locus;		/* [n_tiles][tile_slots], -1 = padding */
	const int *slot_off;		/* [n_tiles][tile_slots] off[l] */
	const int *slot_J;		/* [n_tiles][tile_slots] J[l] */
	const int *group_rowbase;	/* [n_tiles][groups*warps] first smem row */
	const int *group_rows;		/* [n_tiles][groups*warps] rows (max J) */
	const int *tile_rows;		/* [n_tiles] total rows */
	int max_rows;			/* max over tiles: smem sizing */
	/* data */
	const unsigned char *codes;	/* tile-major units */
	long long tile_stride;		/* bytes between tiles */
	/* parameters */
	const double *p;		/* [K][T] p (admixture) or log p (mixture E) */
	const double *eta;		/* eta rows / v_ik rows */
	long long eta_stride;		/* K, or 0 for a shared row */
	/* outputs */
	double *Apart;			/* [n_tiles][Ipad][K] */
	double *Npart;			/* [n_chunks][K*T] */
	double *llpart;			/* [n_units] */
	/* MODE_MIX_E as the fall-back of the digit-sliced pass (mc_digit.cuh) */
	const int *run_if;
	int *n_chunks_dev;
};

/* one recursive-halving step over N live values: lanes whose `mask` bit is
 * clear keep the lower half, the others the upper half */
template <int N>
__device__ __forceinline__ void rs_step(double *a, int mask, int lane,
	int &lo, int &hi)
{
	constexpr int H = (N + 1) / 2;
	const bool up = (lane & mask) != 0;
#pragma unroll
	for (int j = 0; j < H; j++) {
		double vlo = a[j];
		double vhi = (j + H < N) ? a[j + H] : 0.0;
		double send = up ? vlo : vhi;
		double keep = up ? vhi : vlo;
		a[j] = keep + shfl_xor_f64(send, mask);
	}
	if (up)
		lo += H;
	else
		hi = min(hi, lo + H);
}

/* sum a[0..V) over the lanes that share lane % k_split; on return lane holds
 * the totals of indices [lo, hi) in a[0..hi-lo) */
template <int V>
__device__ __forceinline__ void lane_reduce_scatter(double *a, int k_split,
	int lane, int &lo, int &hi)
{
	constexpr int N1 = (V + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2,
		N4 = (N3 + 1) / 2;
	lo = 0;
	hi = V;
	if (k_split <= 16) rs_step<V>(a, 16, lane, lo, hi);
	if (k_split <= 8) rs_step<N1>(a, 8, lane, lo, hi);
	if (k_split <= 4) rs_step<N2>(a, 4, lane, lo, hi);
	if (k_split <= 2) rs_step<N3>(a, 2, lane, lo, hi);
	if (k_split <= 1) rs_step<N4>(a, 1, lane, lo, hi);
}

/* byte e (0..7) of an 8-byte group of allele codes */
__device__ __forceinline__ unsigned group_byte(const uint2 &u, int e)
{
	return ((e < 4 ? u.x : u.y) >> ((e & 3) * 8)) & 0xffu;
}

__device__ __forceinline__ void prefetch_l1(const void *ptr)
{
	asm volatile("prefetch.global.L1 [%0];" :: "l"(ptr));
}

/* A-blocks (groups of individuals whose sums are folded together) reduced
 * across warps per barrier */
#define MC_NB 4

/*
 * Work inside a thread is organised in GROUPS of 8 allele copies of one locus
 * (4 diploid individuals, 2 tetraploids, ...).  Each group runs in phases of
 * straight-line code so that independent work is in flight together -- with
 * one CTA of a few warps per SM (shared memory holds the tile),
 * instruction-level parallelism is what hides the shared-memory and FP64
 * latencies:
 *   1. the 8 p rows are read from shared memory;
 *   2. the 8 sums tmp = sum_k eta_ik p_klj, folded over the k_split lanes;
 *   3. w = c / tmp and the log-likelihood term for the 8 copies;
 *   4. A_ik += p_klj w (registers);
 *   5. B_klj += eta_ik w, a read-modify-write of this thread's private
 *      shared-memory column, in copy order (two copies may share an allele).
 * The allele codes of the next group are loaded, and the eta rows of the next
 * individuals prefetched into L1, while the current group is computed.
 */
template <int KH, int PP, int MODE>
__global__ void __launch_bounds__(256, 1) tile_kernel(const TileArgs a)
{
	if (MODE == MODE_MIX_E) {
		if (a.run_if && !*a.run_if)
			return;
		if (a.n_chunks_dev && blockIdx.x == 0 && threadIdx.x == 0)
			*a.n_chunks_dev = a.n_tiles;
	}
	constexpr int IB = (PP >= 2) ? 16 / PP : 8;	/* individuals per unit */
	constexpr int UB = IB * PP;			/* bytes per unit: 8 or 16 */
	constexpr int NH = UB / 8;			/* groups of 8 copies per unit */
	constexpr int NI = (PP >= 8) ? 1 : 8 / PP;	/* individuals per group */
	constexpr int EPI = 8 / NI;			/* copies per individual in a group */
	constexpr int V = NI * KH;			/* A values per lane */
	constexpr bool SPAN = (PP > 8);			/* one individual spans both groups */
	constexpr int ABU = SPAN ? 1 : NH;		/* A-blocks per unit */
	constexpr int HPA = SPAN ? NH : 1;		/* groups per A-block */
	constexpr bool HAS_A = (MODE == MODE_ADMIX_EM || MODE == MODE_MIX_E);
	constexpr bool HAS_B = (MODE == MODE_ADMIX_EM || MODE == MODE_MIX_M);
	constexpr bool HAS_P = (MODE != MODE_MIX_M);
	constexpr bool HAS_E = (MODE != MODE_MIX_E);
	constexpr bool HAS_LL = (MODE == MODE_ADMIX_EM || MODE == MODE_ADMIX_LL);

	extern __shared__ double smem[];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int W = a.warps, NG = a.groups, LW = a.loci_per_warp;
	const int ks = a.k_split;
	const int kh = lane % ks, lw = lane / ks;
	const int k0 = kh * KH;
	constexpr int RS = KH * 32;		/* doubles per smem row */
	/* one extra all-zero row (index max_rows) serves the missing copies */
	const int zrow = a.max_rows;
	double *p_s = smem;
	double *B_s = p_s + (HAS_P ? (size_t)(a.max_rows + 1) * RS : 0);
	double *scr = B_s + (HAS_B ? (size_t)(a.max_rows + 1) * RS : 0);
	/* scratch: [2][MC_NB][W][ks][V] then W doubles for the ll reduction */
	double *llred = scr + (HAS_A ? 2 * MC_NB * W * ks * V : 0);
	int set = 0;
	/* clamped cluster index: lanes past K read a valid eta entry and meet
	 * p = 0 in their (never flushed) columns */
	int kofs[KH];
#pragma unroll
	for (int kk = 0; kk < KH; kk++)
		kofs[kk] = min(k0 + kk, a.K - 1);

	for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
		const int t = u % a.n_tiles, c = u / a.n_tiles;
		const long long b0 = a.n_blocks * c / a.n_chunks;
		const long long b1 = a.n_blocks * (c + 1) / a.n_chunks;
		const int *s_locus = a.slot_locus + (size_t)t * a.tile_slots;
		const int *s_off = a.slot_off + (size_t)t * a.tile_slots;
		const int *s_J = a.slot_J + (size_t)t * a.tile_slots;
		const int *g_rowbase = a.group_rowbase + (size_t)t * NG * W;
		const int *g_rows = a.group_rows + (size_t)t * NG * W;

		__syncthreads();	/* previous unit has left shared memory */
		/* ---- stage this thread's columns of the tile ---- */
		for (int g = 0; g < NG; g++) {
			const int s = (g * W + w) * LW + lw;
			const int loc = s_locus[s], Jl = s_J[s], ol = s_off[s];
			const int rb = g_rowbase[g * W + w], nr = g_rows[g * W + w];
			for (int j = 0; j < nr; j++)
#pragma unroll
				for (int kk = 0; kk < KH; kk++) {
					const int x = ((rb + j) * KH + kk) * 32 + lane;
					if (HAS_P) {
						double v = 0.0;
						if (loc >= 0 && j < Jl && k0 + kk < a.K)
							v = a.p[(size_t)(k0 + kk) * a.T + ol + j];
						p_s[x] = v;
					}
					if (HAS_B)
						B_s[x] = 0.0;
				}
		}
#pragma unroll
		for (int kk = 0; kk < KH; kk++) {
			const int x = (zrow * KH + kk) * 32 + lane;
			if (HAS_P)
				p_s[x] = 0.0;
			if (HAS_B)
				B_s[x] = 0.0;
		}
		/* columns are thread private: no barrier needed before use */

		double prod = 1.0, ll_slow = 0.0;
		long long esum = 0;
		const unsigned char *tile_codes = a.codes + (size_t)t * a.tile_stride;
		const long long nab = (b1 - b0) * ABU;
		/* position (A-block, group in block, locus group) of the 8-byte
		 * code group that is loaded one step ahead */
		long long pab = 0;
		int phh = 0, pg = 0;
		auto code_ptr = [&]() {
			const long long b = b0 + pab / ABU;
			const int h = (int)(pab % ABU) + phh;
			const int s = (pg * W + w) * LW + lw;
			return reinterpret_cast<const uint2 *>(tile_codes
				+ ((size_t)b * a.tile_slots + s) * UB + h * 8);
		};
		uint2 cw_next = make_uint2(0xffffffffu, 0xffffffffu);
		if (nab > 0)
			cw_next = __ldg(code_ptr());

		for (long long ab = 0; ab < nab; ab++) {
			const long long b = b0 + ab / ABU;
			const int h0 = (int)(ab % ABU);
			const long long i0 = b * IB + (h0 * 8) / PP;
			const int slot = (int)(ab % MC_NB);
			double A[HAS_A ? V : 1];
			if (HAS_A) {
#pragma unroll
				for (int v = 0; v < V; v++)
					A[v] = 0.0;
			}
			double e[HAS_E ? V : 1];
			if (HAS_E) {
#pragma unroll
				for (int n = 0; n < NI; n++) {
					const long long i = min(i0 + n, a.I - 1);
					const double *er = a.eta + (size_t)i * a.eta_stride;
#pragma unroll
					for (int kk = 0; kk < KH; kk++)
						e[n * KH + kk] = __ldg(er + kofs[kk]);
				}
				/* next A-block's rows -> L1 */
				if (lane < NI && a.eta_stride) {
					const long long in = min(i0 + (SPAN ? 1 : NI) + lane, a.I - 1);
					prefetch_l1(a.eta + (size_t)in * a.eta_stride + k0);
				}
			}
#pragma unroll
			for (int hh = 0; hh < HPA; hh++)
			for (int g = 0; g < NG; g++) {
				const int rb = g_rowbase[g * W + w];
				const uint2 cw = cw_next;
				if (++pg == NG) {
					pg = 0;
					if (++phh == HPA) {
						phh = 0;
						++pab;
					}
				}
				if (pab < nab)
					cw_next = __ldg(code_ptr());
				int x0[8];
				bool valid[8];
#pragma unroll
				for (int q = 0; q < 8; q++) {
					const unsigned code = group_byte(cw, q);
					valid[q] = code != MC_MISSING;
					x0[q] = (valid[q] ? rb + (int)code : zrow) * RS + lane;
				}
				/* phase 1: p rows */
				double pr[HAS_P ? 8 * KH : 1];
				if (HAS_P) {
#pragma unroll
					for (int q = 0; q < 8; q++)
#pragma unroll
						for (int kk = 0; kk < KH; kk++)
							pr[q * KH + kk] = p_s[x0[q] + kk * 32];
				}
				if (MODE == MODE_MIX_E) {
#pragma unroll
					for (int q = 0; q < 8; q++)
#pragma unroll
						for (int kk = 0; kk < KH; kk++)
							A[(q / EPI) * KH + kk] += pr[q * KH + kk];
					continue;
				}
				double wgt[8];
				if (MODE == MODE_MIX_M) {
#pragma unroll
					for (int q = 0; q < 8; q++)
						wgt[q] = valid[q] ? 1.0 : 0.0;
				} else {
					/* phase 2: tmp */
					double tmp[8];
#pragma unroll
					for (int q = 0; q < 8; q++) {
						double v = 0.0;
#pragma unroll
						for (int kk = 0; kk < KH; kk++)
							v = fma(e[(q / EPI) * KH + kk], pr[q * KH + kk], v);
						tmp[q] = v;
					}
#pragma unroll
					for (int m = 1; m < 32; m <<= 1)
						if (m < ks) {
#pragma unroll
							for (int q = 0; q < 8; q++)
								tmp[q] += shfl_xor_f64(tmp[q], m);
						}
					/* phase 3: weights and log-likelihood terms.
					 * log(tmp) = exponent*ln2 + log(mantissa): the
					 * mantissas are multiplied up and logged once;
					 * zero / subnormal / non-finite sums (never
					 * seen in a healthy fit) take the slow path
					 * after the straight-line code */
					unsigned bad = 0;
#pragma unroll
					for (int q = 0; q < 8; q++) {
						tmp[q] = valid[q] ? tmp[q] : 1.0;
						/* positive and normal <=> high word in
						 * [0x00100000, 0x7ff00000) */
						bad |= (unsigned)(__double2hiint(tmp[q]) - 0x00100000)
							>= 0x7fe00000u;
						if (MODE == MODE_ADMIX_EM)
							wgt[q] = valid[q] ? mc_rcp(tmp[q]) : 0.0;
					}
					if (!bad) {
						int es = 0;
#pragma unroll
						for (int q = 0; q < 8; q++) {
							const int hi = __double2hiint(tmp[q]);
							es += hi >> 20;
							prod *= __hiloint2double((hi & 0x000fffff) | 0x3ff00000,
								__double2loint(tmp[q]));
						}
						/* prod < 2^8: fold its exponent away */
						const int hi = __double2hiint(prod);
						es += (hi >> 20) - 9 * 1023;
						prod = __hiloint2double((hi & 0x000fffff) | 0x3ff00000,
							__double2loint(prod));
						esum += es;
					} else {
						for (int q = 0; q < 8; q++)
							ll_slow += log(tmp[q]);
					}
				}
				if (MODE == MODE_ADMIX_EM) {
					/* phase 4: A */
#pragma unroll
					for (int q = 0; q < 8; q++)
#pragma unroll
						for (int kk = 0; kk < KH; kk++)
							A[(q / EPI) * KH + kk] = fma(pr[q * KH + kk], wgt[q],
								A[(q / EPI) * KH + kk]);
				}
				if (HAS_B) {
					/* phase 5: B, copy by copy */
#pragma unroll
					for (int q = 0; q < 8; q++) {
						double bv[KH];
#pragma unroll
						for (int kk = 0; kk < KH; kk++)
							bv[kk] = B_s[x0[q] + kk * 32];
#pragma unroll
						for (int kk = 0; kk < KH; kk++)
							B_s[x0[q] + kk * 32] = fma(e[(q / EPI) * KH + kk],
								wgt[q], bv[kk]);
					}
				}
			}
			if (HAS_A) {
				/* fold the A-block's sums over loci: lanes now, warps
				 * once per MC_NB A-blocks */
				int lo, hi;
				lane_reduce_scatter<V>(A, ks, lane, lo, hi);
				double *mine = scr + ((((size_t)set * MC_NB + slot) * W + w) * ks + kh) * V;
#pragma unroll
				for (int v = 0; v < V; v++)
					if (v < hi - lo)
						mine[lo + v] = A[v];
				if (slot == MC_NB - 1 || ab == nab - 1) {
					__syncthreads();
					const int nslot = slot + 1;
					for (int idx = threadIdx.x; idx < nslot * ks * V; idx += blockDim.x) {
						const int sl = idx / (ks * V), r = idx % (ks * V);
						const int rkh = r / V, v = r % V;
						const int n = v / KH, kk = v % KH;
						const int k = rkh * KH + kk;
						const long long abs_ = ab - slot + sl;
						const long long is = (b0 + abs_ / ABU) * IB
							+ ((int)(abs_ % ABU) * 8) / PP + n;
						double sum = 0.0;
						for (int ww = 0; ww < W; ww++)
							sum += scr[((((size_t)set * MC_NB + sl) * W + ww) * ks + rkh) * V + v];
						if (k < a.K)
							a.Apart[((size_t)t * a.Ipad + is) * a.K + k] = sum;
					}
					set ^= 1;
				}
			}
		}

		/* ---- flush this unit ---- */
		if (HAS_B) {
			double *Np = a.Npart + (size_t)c * a.K * a.T;
			for (int g = 0; g < NG; g++) {
				const int s = (g * W + w) * LW + lw;
				const int loc = s_locus[s], Jl = s_J[s], ol = s_off[s];
				const int rb = g_rowbase[g * W + w];
				if (loc < 0)
					continue;
				for (int j = 0; j < Jl; j++)
#pragma unroll
					for (int kk = 0; kk < KH; kk++)
						if (k0 + kk < a.K) {
							const int x = ((rb + j) * KH + kk) * 32 + lane;
							/* d_iklj = eta p / tmp: the factor p_klj is
							 * common to the whole column, applied once */
							double v = B_s[x];
							if (MODE == MODE_ADMIX_EM)
								v *= p_s[x];
							Np[(size_t)(k0 + kk) * a.T + ol + j] = v;
						}
			}
		}
		if (HAS_LL) {
			double ll = (kh == 0) ? (log(prod) + (double)esum * 0.693147180559945309417232121458 + ll_slow) : 0.0;
#pragma unroll
			for (int m = 16; m >= 1; m >>= 1)
				ll += shfl_xor_f64(ll, m);
			__syncthreads();
			if (lane == 0)
				llred[w] = ll;
			__syncthreads();
			if (threadIdx.x == 0) {
				double sum = 0.0;
				for (int ww = 0; ww < W; ww++)
					sum += llred[ww];
				a.llpart[u] = sum;
			}
		}
	}
}

