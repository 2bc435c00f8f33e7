/*
 * mc_kernels.cuh -- sm_100a kernels of the MULTICLUST EM hot path.
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

#include <cuda_runtime.h>
#include <stdint.h>

#define MC_MISSING 255

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
};

__device__ __forceinline__ double mc_rcp(double x)
{
#ifdef MC_EXACT_DIV
	return 1.0 / x;
#else
	/* MUFU.RCP64H seed + two Newton steps: ~1 ulp, a third of the
	 * instructions of the IEEE division sequence */
	double r, e;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	return r;
#endif
}

__device__ __forceinline__ double shfl_xor_f64(double v, int mask)
{
	int lo = __double2loint(v), hi = __double2hiint(v);
	lo = __shfl_xor_sync(0xffffffffu, lo, mask);
	hi = __shfl_xor_sync(0xffffffffu, hi, mask);
	return __hiloint2double(hi, lo);
}

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

/* ------------------------------------------------------------------ */
/* natural [I][L][P] codes -> tile-major 16-byte units                  */

__global__ void k_tile_codes(const unsigned char *nat, unsigned char *out,
	const int *slot_locus, int n_tiles, int tile_slots, long long n_blocks,
	long long I, int L, int P, int PP, int IB, long long tile_stride)
{
	const long long n = (long long)n_tiles * n_blocks * tile_slots;
	const int UB = IB * PP;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int s = (int)(x % tile_slots);
		const long long b = (x / tile_slots) % n_blocks;
		const int t = (int)(x / tile_slots / n_blocks);
		const int loc = slot_locus[(size_t)t * tile_slots + s];
		unsigned char *o = out + (size_t)t * tile_stride + ((size_t)b * tile_slots + s) * UB;
		for (int ii = 0; ii < IB; ii++) {
			const long long i = b * IB + ii;
			for (int ap = 0; ap < PP; ap++) {
				unsigned char c = MC_MISSING;
				if (loc >= 0 && i < I && ap < P)
					c = nat[((size_t)i * L + loc) * P + ap];
				o[ii * PP + ap] = c;
			}
		}
	}
}

/* ------------------------------------------------------------------ */
/* Michelot projection with a floor (simplex.c:109-143) on a strided row */

__device__ __forceinline__ void project_row(double *x, int n, double floor_)
{
	unsigned fixed[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };	/* n <= 256 */
	int nfree = n;
	while (nfree) {
		double csum = 0.0;
		for (int i = 0; i < n; i++)
			csum += x[i];
		const double shift = (csum - 1.0) / nfree;
		bool done = true;
		for (int i = 0; i < n; i++)
			if (!(fixed[i >> 5] >> (i & 31) & 1u)) {
				double v = x[i] - shift;
				if (v < floor_) {
					v = floor_;
					fixed[i >> 5] |= 1u << (i & 31);
					nfree--;
					done = false;
				}
				x[i] = v;
			}
		if (done)
			break;
	}
}

/* p rows: one thread per (k, l) */
__global__ void k_project_p(double *p, const int *J, const int *off, int K,
	int L, long long T, double lb)
{
	const long long n = (long long)K * L;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int k = (int)(x / L), l = (int)(x % L);
		project_row(p + (size_t)k * T + off[l], J[l], lb);
	}
}

/* eta rows: one thread per row of length K */
__global__ void k_project_eta(double *eta, long long rows, int K, double lb)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
		i += (long long)gridDim.x * blockDim.x)
		project_row(eta + (size_t)i * K, K, lb);
}

/* ------------------------------------------------------------------ */
/* admixture M-step, eta side (em_alg.c:650-702): A_ik = sum over tiles,
 * D_ik = eta_ik A_ik, eta_ik = D_ik / sum_k D_ik, projection             */

/* individuals per block of k_admix_eta: 32, fewer when K is large */
static inline int eta_rows(int K) { return K <= 128 ? 32 : (4096 / K > 0 ? 4096 / K : 1); }

__global__ void k_admix_eta(const double *Apart, int n_tiles, long long Ipad,
	long long I, int K, const double *eta_f, long long eta_stride,
	double *eta_t, double *D, int per_indiv, int do_proj, double lb, int ETA_ROWS)
{
	/* one thread per (individual, k) adds the per-tile partial sums: a warp
	 * reads contiguous runs of Apart; then one thread per individual
	 * normalises and projects its row */
	extern __shared__ double rows[];	/* [ETA_ROWS][K] */
	const int n = ETA_ROWS * K;
	for (long long i0 = (long long)blockIdx.x * ETA_ROWS; i0 < I;
		i0 += (long long)gridDim.x * ETA_ROWS) {
		for (int x = threadIdx.x; x < n; x += blockDim.x) {
			const long long i = i0 + x / K;
			if (i >= I)
				continue;
			const int k = x % K;
			/* four partial sums (tiles t % 4), eight loads in flight: the
			 * pass is a pure stream over Apart; the order is fixed */
			const double *src = Apart + (size_t)i0 * K + x;
			const size_t ts = (size_t)Ipad * K;
			double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
			int t = 0;
			for (; t + 8 <= n_tiles; t += 8) {
				const double v0 = __ldg(src + (size_t)t * ts), v1 = __ldg(src + (size_t)(t + 1) * ts);
				const double v2 = __ldg(src + (size_t)(t + 2) * ts), v3 = __ldg(src + (size_t)(t + 3) * ts);
				const double v4 = __ldg(src + (size_t)(t + 4) * ts), v5 = __ldg(src + (size_t)(t + 5) * ts);
				const double v6 = __ldg(src + (size_t)(t + 6) * ts), v7 = __ldg(src + (size_t)(t + 7) * ts);
				a0 += v0; a1 += v1; a2 += v2; a3 += v3;
				a0 += v4; a1 += v5; a2 += v6; a3 += v7;
			}
			for (; t < n_tiles; t++)
				a0 += __ldg(src + (size_t)t * ts);
			const double acc = (a0 + a1) + (a2 + a3);
			const double d = eta_f[(size_t)i * eta_stride + k] * acc;
			D[(size_t)i * K + k] = d;
			rows[x] = d;
		}
		__syncthreads();
		if (per_indiv && threadIdx.x < ETA_ROWS && i0 + threadIdx.x < I) {
			const double *r = rows + threadIdx.x * K;
			double *row = eta_t + (size_t)(i0 + threadIdx.x) * K;
			double s = 0.0;
			for (int k = 0; k < K; k++)
				s += r[k];
			for (int k = 0; k < K; k++)
				row[k] = r[k] / s;
			if (do_proj)
				project_row(row, K, lb);
		}
		__syncthreads();
	}
}

/* ------------------------------------------------------------------ */
/* deterministic reductions: fixed grid, fixed tree                      */

#define RED_BLOCKS 296
#define RED_THREADS 256

__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1)
		v += shfl_xor_f64(v, m);
	__syncthreads();
	if ((threadIdx.x & 31) == 0)
		sh[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0.0;
	if (threadIdx.x == 0)
		for (int w = 0; w < (int)(blockDim.x >> 5); w++)
			s += sh[w];
	return s;	/* valid in thread 0 */
}

/* out[b*ncol + c] = sum over this block's rows of x[r*ncol + c], c < ncol <= 8 */
__global__ void k_colsum_partial(const double *x, long long rows, int ncol,
	long long row_stride, double *part)
{
	__shared__ double sh[RED_THREADS / 32];
	for (int c = 0; c < ncol; c++) {
		double v = 0.0;
		for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
			r += (long long)gridDim.x * blockDim.x)
			v += x[(size_t)r * row_stride + c];
		v = block_sum(v, sh);
		if (threadIdx.x == 0)
			part[(size_t)blockIdx.x * ncol + c] = v;
	}
}

/* final stage: one block, out[c] (+)= sum_b part[b*ncol + c] */
__global__ void k_colsum_final(const double *part, int nblocks, int ncol,
	double *out)
{
	__shared__ double sh[RED_THREADS / 32];
	for (int c = 0; c < ncol; c++) {
		double v = 0.0;
		for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
			v += part[(size_t)b * ncol + c];
		v = block_sum(v, sh);
		if (threadIdx.x == 0)
			out[c] = v;
	}
}

/* N[x] = sum_c Npart[c][x] (+ add) : allele-count sums over chunks */
__global__ void k_sum_chunks(const double *Npart, int n_chunks, long long n,
	double add, double *out)
{
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		double s = add;
		for (int c = 0; c < n_chunks; c++)
			s += Npart[(size_t)c * n + x];
		out[x] = s;
	}
}

/* p side of both M-steps (em_alg.c:706-752, 965-1010): normalise each (k,l)
 * row of the count sums and project.  The sums arrive as n_chunks partial
 * vectors (added in chunk order, exactly like k_sum_chunks) plus the mixture's
 * pseudo-count `add` on every slot (em_alg.c:972). */
__global__ void k_update_p(const double *N, int n_chunks, long long chunk_stride,
	double add, double *p_t, const int *J, const int *off, int K, int L, long long T,
	int do_proj, double lb)
{
	const long long n = (long long)K * L;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int k = (int)(x / L), l = (int)(x % L);
		const double *nr = N + (size_t)k * T + off[l];
		double *row = p_t + (size_t)k * T + off[l];
		const int Jl = J[l];
		double s = 0.0;
		for (int j = 0; j < Jl; j++) {
			double t = 0.0;
			for (int c = 0; c < n_chunks; c++)
				t += nr[(size_t)c * chunk_stride + j];
			const double v = add + t;
			row[j] = v;
			s += v;
		}
		for (int j = 0; j < Jl; j++)
			row[j] = row[j] / s;
		if (do_proj)
			project_row(row, Jl, lb);
	}
}

/* pooled eta (em_alg.c:604-648, 916-962): eta_k = S_k / sum S, projection */
__global__ void k_update_eta_pooled(const double *S, double *eta_t, int K,
	int do_proj, double lb)
{
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		double s = 0.0;
		for (int k = 0; k < K; k++)
			s += S[k];
		for (int k = 0; k < K; k++)
			eta_t[k] = S[k] / s;
		if (do_proj)
			project_row(eta_t, K, lb);
	}
}

/* log p table for the mixture passes; zero_skip reproduces the E-step's
 * "p == 0 contributes nothing" rule (em_alg.c:797-804) */
__global__ void k_log_table(const double *p, double *lp, long long n,
	int zero_skip)
{
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const double v = p[x];
		lp[x] = (zero_skip && v == 0.0) ? 0.0 : log(v);
	}
}

/* ------------------------------------------------------------------ */
/* parameter-space kernels of the acceleration schemes                   */

__global__ void k_delta(double *d, const double *xt, const double *xf,
	long long n)
{
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x)
		d[x] = xt[x] - xf[x];
}

/* {u.u, u.(v-u), (v-u).(v-u)} (accel_em.c:142-184) -> part[b][3] */
__global__ void k_step_dots(const double *u, const double *v, long long n,
	double *part)
{
	__shared__ double sh[RED_THREADS / 32];
	double a = 0.0, b = 0.0, c = 0.0;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const double uu = u[x], r = v[x] - uu;
		a += uu * uu;
		b += uu * r;
		c += r * r;
	}
	a = block_sum(a, sh);
	b = block_sum(b, sh);
	c = block_sum(c, sh);
	if (threadIdx.x == 0) {
		part[(size_t)blockIdx.x * 3 + 0] = a;
		part[(size_t)blockIdx.x * 3 + 1] = b;
		part[(size_t)blockIdx.x * 3 + 2] = c;
	}
}

/* {u1.u2, u1.v2} (accel_em.c:291-310) -> part[b][2] */
__global__ void k_qn_dots(const double *u1, const double *u2, const double *v2,
	long long n, double *part)
{
	__shared__ double sh[RED_THREADS / 32];
	double a = 0.0, b = 0.0;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const double uu = u1[x];
		a += uu * u2[x];
		b += uu * v2[x];
	}
	a = block_sum(a, sh);
	b = block_sum(b, sh);
	if (threadIdx.x == 0) {
		part[(size_t)blockIdx.x * 2 + 0] = a;
		part[(size_t)blockIdx.x * 2 + 1] = b;
	}
}

/* accel_em.c:449-466 / 486-503, same expression shapes as the reference */
__global__ void k_accel_update(double *xt, const double *xp, const double *u,
	const double *v, long long n, double s, int qn1)
{
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		if (qn1)
			xt[x] = __dadd_rn(__dadd_rn(xp[x], u[x]), __dmul_rn(s, v[x]));
		else
			xt[x] = __dadd_rn(__dadd_rn(xp[x], -__dmul_rn(__dmul_rn(2.0, s), u[x])),
				__dmul_rn(__dmul_rn(s, s), __dadd_rn(v[x], -u[x])));
	}
}

#define MC_QMAX 3
struct QnArgs {
	const double *v[MC_QMAX];	/* v of the pair used by row j */
	double coef[MC_QMAX][MC_QMAX][2];	/* Ainv[j][n], cutu[n] */
	int q;
};

/* accel_em.c:364-402 */
__global__ void k_qn_update(double *xt, const double *xp, const double *uu,
	long long n, const QnArgs qa)
{
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		double acc = __dadd_rn(xp[x], uu[x]);
		for (int j = 0; j < qa.q; j++) {
			const double vv = qa.v[j][x];
			for (int m = 0; m < qa.q; m++)
				acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(vv, qa.coef[j][m][0]),
					qa.coef[j][m][1]));
		}
		xt[x] = acc;
	}
}

/* argmax_k of the posterior, first maximum wins (write_file.c:369-375,590-598) */
__global__ void k_partition(const double *post, long long I, int K, int *I_K)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < I;
		i += (long long)gridDim.x * blockDim.x) {
		const double *r = post + (size_t)i * K;
		int best = 0;
		double m = r[0];
		for (int k = 1; k < K; k++)
			if (r[k] > m) {
				m = r[k];
				best = k;
			}
		I_K[i] = best;
	}
}

/* sums of the posterior rows over the individuals of every sampling locale
 * (write_file.c:446-459, 658-666: the popq tables).  Block b owns LS_ROWS
 * consecutive individuals; thread (locale, k) adds its block's rows in index
 * order, a second kernel adds the block sums in block order: deterministic. */
#define LS_ROWS 256
__global__ void k_locale_partial(const double *post, const int *locale, long long I, int K,
	int n_loc, double *part /* [blocks][n_loc * K] */)
{
	__shared__ int loc_s[LS_ROWS];
	const long long i0 = (long long)blockIdx.x * LS_ROWS;
	const int rows = (int)(I - i0 < LS_ROWS ? I - i0 : LS_ROWS);
	for (int x = threadIdx.x; x < rows; x += blockDim.x)
		loc_s[x] = locale[i0 + x];
	__syncthreads();
	for (int x = threadIdx.x; x < n_loc * K; x += blockDim.x) {
		const int n = x / K, k = x - n * K;
		double s = 0.0;
		for (int j = 0; j < rows; j++)
			if (loc_s[j] == n)
				s += post[(size_t)(i0 + j) * K + k];
		part[(size_t)blockIdx.x * n_loc * K + x] = s;
	}
}

__global__ void k_locale_final(const double *part, int blocks, int n, double *out)
{
	for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
		double s = 0.0;
		for (int b = 0; b < blocks; b++)
			s += part[(size_t)b * n + x];
		out[x] = s;
	}
}

/* ------------------------------------------------------------------ */
/* admixture initialiser (rnd_init.c:456-482): hard assignment counts     */

/* one CTA per individual, threads stride over its loci (coalesced reads of the
 * codes and of z); D_i is counted in shared memory, the allele counts with
 * integer atomics in global memory: order-independent, so deterministic */
__global__ void k_init_counts(const unsigned char *nat, const unsigned char *z,
	long long I, int L, int P, int K, const int *off, long long T,
	double *D /* [I][K] */, unsigned *N /* [K][T] */)
{
	__shared__ unsigned Dsm[256];
	for (long long i = blockIdx.x; i < I; i += gridDim.x) {
		for (int k = threadIdx.x; k < K; k += blockDim.x)
			Dsm[k] = 0u;
		__syncthreads();
		for (int l = threadIdx.x; l < L; l += blockDim.x) {
			const unsigned char *c = nat + ((size_t)i * L + l) * P;
			const unsigned char *zz = z + ((size_t)i * L + l) * P;
			for (int ap = 0; ap < P; ap++) {
				if (c[ap] == MC_MISSING)
					continue;
				bool seen = false;
				for (int b = 0; b < ap; b++)
					seen |= (c[b] == c[ap] && zz[b] == zz[ap]);
				if (seen)
					continue;
				atomicAdd(&Dsm[zz[ap]], 1u);
				atomicAdd(&N[(size_t)zz[ap] * T + off[l] + c[ap]], 1u);
			}
		}
		__syncthreads();
		for (int k = threadIdx.x; k < K; k += blockDim.x)
			D[(size_t)i * K + k] = (double)Dsm[k];
		__syncthreads();
	}
}

/* the reference's cluster draws made on the device: thread b continues glibc's
 * TYPE_3 rand() -- x[n] = x[n-31] + x[n-3] mod 2^32, draw = x[n] >> 1 -- from
 * the 31 words in front of its block of draws and writes z = draw % K for the
 * block's allele copies (rnd_init.c:460-481).  The 31 words stay in registers:
 * 496 = 16 x 31 draws per round, all indices static. */
#define MC_RAND_ROUND 496
__global__ void k_rand_assign(const unsigned *hist, long long n_blocks, long long block_draws,
	long long n, unsigned K, unsigned char *z)
{
	const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (b >= n_blocks)
		return;
	unsigned h[31];
#pragma unroll
	for (int j = 0; j < 31; j++)
		h[j] = hist[(size_t)b * 31 + j];
	const long long first = b * block_draws;
	const long long cnt = n - first < block_draws ? n - first : block_draws;
	unsigned char *out = z + first;
	long long done = 0;
	for (; done + MC_RAND_ROUND <= cnt; done += MC_RAND_ROUND) {
#pragma unroll
		for (int g = 0; g < 31; g++) {
			unsigned w[4] = { 0u, 0u, 0u, 0u };
#pragma unroll
			for (int q = 0; q < 16; q++) {
				const int i = (g * 16 + q) % 31;
				h[i] += h[(i + 28) % 31];
				w[q >> 2] |= ((h[i] >> 1) % K) << ((q & 3) * 8);
			}
			*reinterpret_cast<uint4 *>(out + done + g * 16) = make_uint4(w[0], w[1], w[2], w[3]);
		}
	}
	if (done < cnt) {	/* tail of the last block: a circular buffer in local memory */
		unsigned t[31];
#pragma unroll
		for (int j = 0; j < 31; j++)
			t[j] = h[j];
		int f = 0;
		for (; done < cnt; done++) {
			t[f] += t[f >= 3 ? f - 3 : f + 28];
			out[done] = (unsigned char)((t[f] >> 1) % K);
			if (++f == 31)
				f = 0;
		}
	}
}

/* ------------------------------------------------------------------ */
/* mixture initialiser (rnd_init.c:192-339) on the device                 */

/* L1 distance between the allele-count vectors of two genotypes at one locus
 * (rnd_init.c:238-247), from the codes: sum_j |n_a(j) - n_b(j)| over the
 * alleles either of them carries.  Integer, so the comparison below is exact. */
__device__ __forceinline__ int locus_distance(const unsigned char *ca, const unsigned char *cb, int P)
{
	int d = 0;
	for (int x = 0; x < P; x++) {
		if (ca[x] == MC_MISSING)
			continue;
		bool first = true;
		for (int y = 0; y < x; y++)
			first &= ca[y] != ca[x];
		if (!first)
			continue;
		int na = 0, nb = 0;
		for (int y = 0; y < P; y++) {
			na += ca[y] == ca[x];
			nb += cb[y] == ca[x];
		}
		d += abs(na - nb);
	}
	for (int x = 0; x < P; x++) {	/* alleles only b carries */
		if (cb[x] == MC_MISSING)
			continue;
		bool first = true, in_a = false;
		int nb = 0;
		for (int y = 0; y < x; y++)
			first &= cb[y] != cb[x];
		for (int y = 0; y < P; y++) {
			in_a |= ca[y] == cb[x];
			nb += cb[y] == cb[x];
		}
		if (first && !in_a)
			d += nb;
	}
	return d;
}

/* nearest centre of every individual, strictly smaller distance wins, centres
 * keep their own cluster (rnd_init.c:220-258); one CTA per individual, threads
 * stride over the loci; also counts the cluster sizes */
__global__ void k_mix_assign(const unsigned char *nat, const unsigned char *centers,
	const int *center_idx, long long I, int L, int P, int K, int *part, unsigned *nk)
{
	__shared__ int red[32];
	const size_t row = (size_t)L * P;
	for (long long i = blockIdx.x; i < I; i += gridDim.x) {
		int mine = -1;		/* a centre keeps the first cluster it is the centre of */
		for (int k = K - 1; k >= 0; k--)
			if (center_idx[k] == i)
				mine = k;
		long long best = -1;
		int bk = 0;
		if (K > 1 && !(mine == 0)) {
			for (int k = 0; k < K; k++) {
				if (mine == k) {	/* uniform over the block */
					bk = k;
					break;
				}
				int d = 0;
				for (int l = threadIdx.x; l < L; l += blockDim.x)
					d += locus_distance(nat + (size_t)i * row + (size_t)l * P,
						centers + (size_t)k * row + (size_t)l * P, P);
				for (int m = 16; m >= 1; m >>= 1)
					d += __shfl_xor_sync(0xffffffffu, d, m);
				__syncthreads();
				if ((threadIdx.x & 31) == 0)
					red[threadIdx.x >> 5] = d;
				__syncthreads();
				long long tot = 0;
				for (int w = 0; w < (int)(blockDim.x >> 5); w++)
					tot += red[w];
				if (best < 0 || tot < best) {
					best = tot;
					bk = k;
				}
			}
		}
		if (threadIdx.x == 0) {
			part[i] = bk;
			atomicAdd(&nk[bk], 1u);
		}
		__syncthreads();
	}
}

/* allele counts of every cluster (rnd_init.c:296-318): every observed copy adds one */
__global__ void k_mix_init_counts(const unsigned char *nat, const int *part, long long I,
	int L, int P, const int *off, long long T, unsigned *N /* [K][T] */)
{
	for (long long i = blockIdx.x; i < I; i += gridDim.x) {
		const int k = part[i];
		for (int l = threadIdx.x; l < L; l += blockDim.x) {
			const unsigned char *c = nat + ((size_t)i * L + l) * P;
			for (int ap = 0; ap < P; ap++)
				if (c[ap] != MC_MISSING)
					atomicAdd(&N[(size_t)k * T + off[l] + c[ap]], 1u);
		}
	}
}

/* eta_k = (1 + n_k) / (I + K); p_klj = (1 + (K - k) S_klj) / sum_j (...)
 * (rnd_init.c:274-338), from the summed counts */
__global__ void k_mix_init_finish(const double *S, const double *nk, double *eta_t, double *p_t,
	const int *J, const int *off, int K, int L, long long T, long long I_total)
{
	const long long n = (long long)K * L;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int k = (int)(x / L), l = (int)(x % L);
		const double *sr = S + (size_t)k * T + off[l];
		double *row = p_t + (size_t)k * T + off[l];
		double sum = 0.0;
		for (int m = 0; m < J[l]; m++) {
			row[m] = 1.0 + (K - k) * sr[m];
			sum += row[m];
		}
		for (int m = 0; m < J[l]; m++)
			row[m] /= sum;
		if (l == 0)
			eta_t[k] = (1.0 + nk[k]) / (double)(I_total + K);
	}
}

__global__ void k_u32_to_f64(const unsigned *x, double *y, long long n)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
		i += (long long)gridDim.x * blockDim.x)
		y[i] = (double)x[i];
}

/* eta rows from D (em_alg.c:650-702) */
__global__ void k_eta_from_D(const double *D, double *eta_t, long long I, int K,
	int do_proj, double lb)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < I;
		i += (long long)gridDim.x * blockDim.x) {
		double s = 0.0;
		for (int k = 0; k < K; k++)
			s += D[(size_t)i * K + k];
		double *row = eta_t + (size_t)i * K;
		for (int k = 0; k < K; k++)
			row[k] = D[(size_t)i * K + k] / s;
		if (do_proj)
			project_row(row, K, lb);
	}
}

/* ------------------------------------------------------------------ */
/* synthetic data straight into HBM (include/mc_synth.h)                 */

__global__ void k_synth_fill(unsigned char *nat, long long I, int L,
	mcs_params g, long long i_first, unsigned *present /* [L][8] */,
	unsigned *missing /* [L] */)
{
	const long long n = I * (long long)L;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const long long i = x / L;
		const int l = (int)(x % L);
		for (int ap = 0; ap < g.ploidy; ap++) {
			const unsigned char c = mcs_code(&g, i_first + i, l, ap);
			nat[(size_t)x * g.ploidy + ap] = c;
			if (c == MC_MISSING)
				atomicOr(&missing[l], 1u);
			else
				atomicOr(&present[(size_t)l * 8 + (c >> 5)], 1u << (c & 31));
		}
	}
}

__global__ void k_synth_remap(unsigned char *nat, long long I, int L, int P,
	const unsigned char *map /* [L][256] */)
{
	const long long n = I * (long long)L;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int l = (int)(x % L);
		for (int ap = 0; ap < P; ap++) {
			const unsigned char c = nat[(size_t)x * P + ap];
			if (c != MC_MISSING)
				nat[(size_t)x * P + ap] = map[(size_t)l * 256 + c];
		}
	}
}
