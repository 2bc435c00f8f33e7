/*
 * mc_admix3_build.cuh -- one-time layout builders of the two-pass admixture
 * kernel (tile codes, entry lists, column and lane tables).  Included by
 * mc_cuda.cu only.
 */
#pragma once

#include "mc_admix3.cuh"

/* ---------------------------------------------------------------------- */
/* one-time layout builders                                                 */

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

/* natural [I][L][P] codes -> A3_NC bytes per (tile, individual): LT = A3_NC / PP
 * loci x PP copies, thread-major inside a tile */
__global__ void k3_build_codes(const unsigned char *nat, unsigned char *codes,
	long long I, int L, int P, int PP, int n_itiles, int n_ltiles)
{
	const int LT = A3_NC / PP;
	const long long n = (long long)n_itiles * n_ltiles * A3_THREADS;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int t = (int)(x % A3_THREADS);
		const long long r = x / A3_THREADS;
		const int lt = (int)(r % n_ltiles);
		const long long i = (r / n_ltiles) * A3_IT + t;
		for (int h = 0; h < A3_NC / 8; h++) {
			unsigned char b[8];
			for (int q = 0; q < 8; q++) {
				const int l = lt * LT + (h * 8 + q) / PP, a = q % PP;
				b[q] = (i < I && l < L && a < P) ? nat[((size_t)i * L + l) * P + a] : 255;
			}
			*reinterpret_cast<uint2 *>(codes + (size_t)x * A3_NC + h * 8)
				= *reinterpret_cast<uint2 *>(b);
		}
	}
}

/* entry lists of one (itile, ltile): for every real allele column in `colinfo`
 * order, one entry per (individual, allele) carrying it:
 *     i | first copy << 9 | (count - 1) << 12.
 * The column owns S consecutive pass-2 lanes, S chosen per tile from the tile's
 * own counts so that no lane gets more than q = ceil(entries / A3_THREADS) (+1)
 * entries; lane seg reads the entries at start + s * S + seg, s = 0, 1, ...  The eta rows
 * of 8 individuals with different i % 8 lie in different bank groups, so an
 * entry is dealt to a slot whose (lane + step) % 8 equals i % 8 wherever such a
 * slot is still free; the rest fill the remaining positions in ascending order
 * of i. */
__global__ void k3_build_csc(const unsigned char *codes, int PP, int n_ltiles,
	int ncolmax, int cap, const int *lt_ncol, const unsigned short *colinfo,
	unsigned short *csc, unsigned short *colstart)
{
	extern __shared__ unsigned char sm3[];
	unsigned char *cd = sm3;				/* [A3_IT][A3_NC] */
	int *cnt = reinterpret_cast<int *>(sm3 + (size_t)A3_IT * A3_NC);	/* [ncolmax] */
	int *lane_first = cnt + ncolmax;				/* [ncolmax + 1] */
	/* carriers of every column as a bit mask over the tile's individuals: the
	 * sweeps below then visit carriers only (a sixth of the individuals at
	 * config 3) instead of testing every individual three times */
	constexpr int MW = A3_IT / 32;
	unsigned *mask = reinterpret_cast<unsigned *>(lane_first + ncolmax + 1);	/* [ncolmax][MW] */
	const int lt = blockIdx.x % n_ltiles;
	const int ncol = lt_ncol[lt];
	const unsigned short *ci = colinfo + (size_t)lt * ncolmax;
	unsigned short *out = csc + (size_t)blockIdx.x * cap;
	const int csw = ((ncolmax + 1 + 7) / 8) * 8;
	unsigned short *cs = colstart + (size_t)blockIdx.x * (3 * csw + A3_THREADS / 2);
	const uint2 *src = reinterpret_cast<const uint2 *>(codes)
		+ (size_t)blockIdx.x * A3_THREADS * (A3_NC / 8);

	for (int x = threadIdx.x; x < A3_IT * (A3_NC / 8); x += blockDim.x)
		reinterpret_cast<uint2 *>(cd)[x] = src[x];
	__syncthreads();
	for (int x = threadIdx.x; x < ncol * MW; x += blockDim.x) {
		const int c = x / MW, w = x - c * MW;
		const int ll = ci[c] >> 8, j = ci[c] & 0xff;
		unsigned m = 0;
		for (int b = 0; b < 32; b++) {
			const unsigned char *pc = cd + (w * 32 + b) * A3_NC + ll * PP;
			bool has = false;
			for (int a = 0; a < PP; a++)
				has |= pc[a] == j;
			m |= (unsigned)has << b;
		}
		mask[x] = m;
	}
	__syncthreads();
	for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
		int n = 0;
		for (int w = 0; w < MW; w++)
			n += __popc(mask[c * MW + w]);
		cnt[c] = n;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		int acc = 0;
		for (int c = 0; c < ncol; c++) {
			const int n = cnt[c];
			cs[c] = (unsigned short)acc;
			acc += n;
		}
		for (int c = ncol; c < csw; c++)
			cs[c] = (unsigned short)acc;
		/* lanes of THIS tile: the smallest list length q with
		 * sum_c ceil(n_c / q) <= A3_THREADS, column c gets ceil(n_c / q) lanes */
		int q = (acc + A3_THREADS - 1) / A3_THREADS;
		if (q < 1)
			q = 1;
		for (;; q++) {
			int lanes = 0;
			for (int c = 0; c < ncol; c++)
				lanes += (cnt[c] + q - 1) / q;
			if (lanes <= A3_THREADS)
				break;
		}
		int l0 = 0;
		for (int c = 0; c < ncol; c++) {
			lane_first[c] = l0;
			cs[csw + c] = (unsigned short)l0;
			l0 += (cnt[c] + q - 1) / q;
		}
		lane_first[ncol] = l0;
		for (int c = ncol; c < csw; c++)
			cs[csw + c] = (unsigned short)l0;
		for (int c = 0; c < csw; c++)	/* column -> locus_in_tile << 8 | allele */
			cs[2 * csw + c] = c < ncol ? ci[c] : 0;
	}
	__syncthreads();
	/* the column of every pass-2 lane, two lanes per 16-bit word */
	for (int x = threadIdx.x; x < A3_THREADS / 2; x += blockDim.x) {
		unsigned v = 0;
		for (int h = 0; h < 2; h++) {
			const int ln = 2 * x + h;
			int col = 255;
			if (ln < lane_first[ncol]) {
				int lo = 0, hi = ncol - 1;	/* last column with lane_first <= ln */
				while (lo < hi) {
					const int mid = (lo + hi + 1) >> 1;
					if (lane_first[mid] <= ln)
						lo = mid;
					else
						hi = mid - 1;
				}
				col = lo;
				/* columns without entries own no lane: step to the owner */
				while (lane_first[col + 1] <= ln)
					col++;
			}
			v |= (unsigned)col << (8 * h);
		}
		cs[3 * csw + x] = (unsigned short)v;
	}
	for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
		const int ll = ci[c] >> 8, j = ci[c] & 0xff;
		const int n = cnt[c], start = cs[c];
		const int lane0 = lane_first[c], S = lane_first[c + 1] - lane0;
		if (!n)
			continue;
		const int q = n / S, rem = n - q * S;	/* lane seg holds q + (seg < rem) entries */
		for (int x = 0; x < n; x++)
			out[start + x] = 0xffff;
		/* slot (seg, s) belongs to residue class (lane0 + seg + s) % 8: in step s
		 * the 8 lanes of a quarter warp then want 8 different residues, and every
		 * lane meets every residue once in 8 steps, so a column finds room for
		 * all residues however few lanes it owns.  Two sweeps over the carriers:
		 * the first places the entries that find a slot of their class, the
		 * second the others */
		for (int sweep = 0; sweep < 2; sweep++) {
			int cs_s[8], cs_seg[8];		/* next free slot of every class */
			for (int r = 0; r < 8; r++) {
				cs_s[r] = 0;
				cs_seg[r] = (r - lane0) & 7;
			}
			int fill = 0;
			for (int w = 0; w < MW; w++)
			for (unsigned mm = mask[c * MW + w]; mm; mm &= mm - 1) {
				const int ii = w * 32 + __ffs((int)mm) - 1;
				int cn = 0, first = 0;
				for (int a = PP - 1; a >= 0; a--)
					if (cd[ii * A3_NC + ll * PP + a] == j) {
						cn++;
						first = a;
					}
				const int r = ii & 7;
				/* skip over slots that do not exist (segments beyond S, the
				 * short last step) */
				int s_ = cs_s[r], seg = cs_seg[r];
				while (s_ <= q && (seg >= S || (s_ == q && seg >= rem))) {
					s_++;
					seg = (r - lane0 - s_) & 7;
				}
				const bool ok = s_ < q || (s_ == q && seg < rem);
				if (ok) {
					if (sweep == 0)
						out[start + s_ * S + seg] = (unsigned short)(ii | first << 9
							| (cn - 1) << 12);
					cs_s[r] = s_;
					cs_seg[r] = seg + 8;
				} else {
					cs_s[r] = q + 1;
					if (sweep == 1) {
						while (out[start + fill] != 0xffff)
							fill++;
						out[start + fill++] = (unsigned short)(ii | first << 9
							| (cn - 1) << 12);
					}
				}
			}
		}
	}
}

