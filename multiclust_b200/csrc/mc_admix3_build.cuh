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
 * loci x PP copies, thread-major inside a tile; an allele is stored as its row inside
 * the kernel's p tile (perm_of: allele slot -> row of the locus, mc_cuda.cu) */
__global__ void k3_build_codes(const unsigned char *nat, unsigned char *codes,
	long long I, int L, int P, int PP, int n_itiles, int n_ltiles,
	const int *off, const unsigned char *perm_of)
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
				unsigned char v = (i < I && l < L && a < P)
					? nat[((size_t)i * L + l) * P + a] : 255;
				if (v != 255)
					v = perm_of[off[l] + v];
				b[q] = v;
			}
			*reinterpret_cast<uint2 *>(codes + (size_t)x * A3_NC + h * 8)
				= *reinterpret_cast<uint2 *>(b);
		}
	}
}

/* the parameter slot (or log p table) in the kernel's row order: pp[k][g] = p[k][nat_of[g]] */
__global__ void k3_permute_rows(const double *p, double *pp, const int *nat_of, int K,
	long long T)
{
	const long long n = (long long)K * T;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const long long k = x / T, g = x - k * T;
		pp[x] = p[k * T + nat_of[g]];
	}
}

/* shared memory of k3_build_csc */
static inline size_t a3_build_smem_bytes(int ncolmax)
{
	return (size_t)A3_IT * A3_NC + sizeof(int) * (2 * (size_t)ncolmax + 2)
		+ sizeof(unsigned) * (size_t)ncolmax * (A3_IT / 32)
		+ sizeof(unsigned short) * (size_t)A3_IT * A3_NC
		+ (size_t)A3_THREADS * (8 + 8 + 1 + 1);
}

/* entry lists of one (itile, ltile): for every real allele column in `colinfo`
 * order, one entry per (individual, allele) carrying it:
 *     i | first copy << 9 | (count - 1) << 12.
 * The column owns S consecutive logical pass-2 lanes, S chosen per tile from the
 * tile's own counts so that no lane gets more than q entries (q the smallest list
 * length for which the columns need at most A3_THREADS lanes).  The list of the tile is
 * stored step-major: entry st of logical lane ln is half (st & 1) of the 32-bit word
 * (st / 2) * A3_THREADS + a3_lane_thread(ln).  With `qmax_out` the launch only counts:
 * it reports the largest q of any tile, which sizes the lists.
 *
 * The eta rows of 8 individuals with different i % 8 lie in different bank groups, so a
 * quarter warp's LDS.128 costs as many wavefronts as the most frequent residue among
 * its 8 entries.  Which of its entries a lane reads in which step is free, and the 8
 * lanes of a quarter warp belong to 8 different columns (a3_thread_lane), so the lists
 * are scheduled in two stages:
 *   deal      a column's carriers, residue class by residue class, go round its S
 *             lanes: every lane gets an even share of every class;
 *   schedule  per quarter warp and step, the active lanes -- the one with the fewest
 *             classes left first -- take the best-stocked class nobody in the step has
 *             taken; a lane that finds none asks a holder of one of its classes to
 *             move to another free class (one augmenting step); if that fails too it
 *             takes its best-stocked class and the step costs a wavefront more.
 * Restated in tools/list_schedule_sim.py on tiles of the bench's generator: 1.21
 * wavefronts per quarter-warp step against 1.52 for dealing entries to fixed
 * (lane + step) % 8 slots (a full augmenting-path matching gives the same; balancing the
 * class totals of the quarter warps while dealing is better still, but couples the
 * columns and runs on one thread per tile). */
__global__ void k3_build_csc(const unsigned char *codes, int PP, int n_ltiles,
	int ncolmax, int cap, const int *lt_ncol, const unsigned short *colinfo,
	unsigned short *csc, unsigned short *colstart, int *qmax_out)
{
	extern __shared__ __align__(16) unsigned char sm3[];
	unsigned char *cd = sm3;				/* [A3_IT][A3_NC] */
	int *cnt = reinterpret_cast<int *>(sm3 + (size_t)A3_IT * A3_NC);	/* [ncolmax] */
	int *lane_first = cnt + ncolmax;				/* [ncolmax + 2] (keeps 8-byte alignment) */
	/* carriers of every column as a bit mask over the tile's individuals */
	constexpr int MW = A3_IT / 32;
	unsigned *mask = reinterpret_cast<unsigned *>(lane_first + ncolmax + 2);	/* [ncolmax][MW] */
	unsigned short *tmp = reinterpret_cast<unsigned short *>(mask + (size_t)ncolmax * MW);	/* [A3_IT * A3_NC] dealt lists, column after column */
	unsigned char *lcnt = reinterpret_cast<unsigned char *>(tmp + A3_IT * A3_NC);	/* [lanes][8] entries left per class */
	unsigned char *lptr = lcnt + A3_THREADS * 8;		/* [lanes][8] next list step of the class */
	unsigned char *lcol = lptr + A3_THREADS * 8;		/* [lanes] column */
	unsigned char *lfill = lcol + A3_THREADS;		/* [lanes] entries of the lane */
	const int lt = blockIdx.x % n_ltiles;
	const int ncol = lt_ncol[lt];
	const unsigned short *ci = colinfo + (size_t)lt * ncolmax;
	unsigned short *out = csc + (size_t)blockIdx.x * cap;
	const int csw = ((ncolmax + 1 + 7) / 8) * 8;
	unsigned short *cs = colstart + (size_t)blockIdx.x * (3 * csw + A3_THREADS);
	const uint2 *src = reinterpret_cast<const uint2 *>(codes)
		+ (size_t)blockIdx.x * A3_THREADS * (A3_NC / 8);

	for (int x = threadIdx.x; x < A3_IT * (A3_NC / 8); x += blockDim.x)
		reinterpret_cast<uint2 *>(cd)[x] = src[x];
	for (int x = threadIdx.x; x < A3_THREADS * 8; x += blockDim.x)
		lcnt[x] = 0;
	for (int x = threadIdx.x; x < A3_THREADS; x += blockDim.x) {
		lcol[x] = 255;
		lfill[x] = 0;
	}
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
		for (int c = 0; c < ncol; c++)
			acc += cnt[c];
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
		if (qmax_out) {		/* first launch: the longest list of any tile */
			atomicMax(qmax_out, q);
		} else {
			acc = 0;
			for (int c = 0; c < ncol; c++) {
				cs[c] = (unsigned short)acc;	/* first entry in the dealt order */
				acc += cnt[c];
			}
			for (int c = ncol; c < csw; c++)
				cs[c] = (unsigned short)acc;
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
	}
	if (qmax_out)
		return;
	__syncthreads();
	/* the column of every logical pass-2 lane */
	for (int c = threadIdx.x; c < ncol; c += blockDim.x)
		for (int ln = lane_first[c]; ln < lane_first[c + 1]; ln++)
			lcol[ln] = (unsigned char)c;
	__syncthreads();

	/* ---- deal: one thread per column ---- */
	for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
		const int n = cnt[c];
		if (!n)
			continue;
		const int ll = ci[c] >> 8, j = ci[c] & 0xff;
		const int start = cs[c], lane0 = lane_first[c], S = lane_first[c + 1] - lane0;
		const int q = n / S, rem = n - q * S;	/* lane seg holds q + (seg < rem) entries */
		for (int seg = 0; seg < S; seg++)
			lfill[lane0 + seg] = (unsigned char)(q + (seg < rem));
		int k = 0, seg = 0;
		for (int r = 0; r < 8; r++)
		for (int w = 0; w < MW; w++)
		for (unsigned mm = mask[c * MW + w] & (0x01010101u << r); mm; mm &= mm - 1) {
			const int ii = w * 32 + __ffs((int)mm) - 1;
			int cn = 0, first = 0;
			for (int a = PP - 1; a >= 0; a--)
				if (cd[ii * A3_NC + ll * PP + a] == j) {
					cn++;
					first = a;
				}
			/* the k-th carrier in class order is entry k / S of lane k % S */
			tmp[start + k++] = (unsigned short)(ii | first << 9 | (cn - 1) << 12);
			lcnt[(lane0 + seg) * 8 + r]++;
			if (++seg == S)
				seg = 0;
		}
	}
	__syncthreads();
	/* a lane's dealt list is sorted by class: where each class starts; and the lane
	 * table of the kernel: column | entries << 8 (255: idle) */
	for (int ln = threadIdx.x; ln < A3_THREADS; ln += blockDim.x) {
		cs[3 * csw + ln] = (unsigned short)(lcol[ln] | (unsigned)lfill[ln] << 8);
		int acc = 0;
		for (int r = 0; r < 8; r++) {
			lptr[ln * 8 + r] = (unsigned char)acc;
			acc += lcnt[ln * 8 + r];
		}
	}
	__syncthreads();

	/* ---- schedule: one thread per quarter warp; fixed-length loops, so the threads of
	 * the warp stay together ---- */
	for (int qw = threadIdx.x; qw < A3_NQ; qw += blockDim.x) {
		int lstart[8], lS[8], lseg[8], lquota[8];
		int maxq = 0;
#pragma unroll
		for (int b = 0; b < 8; b++) {
			const int ln = b * A3_NQ + qw;
			const int c = lcol[ln];
			lstart[b] = lS[b] = lseg[b] = lquota[b] = 0;
			if (c == 255)
				continue;
			lstart[b] = cs[c];
			lS[b] = lane_first[c + 1] - lane_first[c];
			lseg[b] = ln - lane_first[c];
			lquota[b] = lfill[ln];
			maxq = max(maxq, lquota[b]);
		}
		/* the 8 per-class counts of lane slot b, one byte each */
		auto counts = [&](int b) {
			return *reinterpret_cast<const unsigned long long *>(lcnt + (b * A3_NQ + qw) * 8);
		};
		auto classes = [](unsigned long long c8) {	/* bit r: class r has entries */
			unsigned m = 0;
#pragma unroll
			for (int r = 0; r < 8; r++)
				m |= (unsigned)((c8 >> (8 * r) & 0xff) != 0) << r;
			return m;
		};
		auto best = [](unsigned long long c8, unsigned among) {	/* best-stocked class */
			int r = 0, bc = -1;
#pragma unroll
			for (int rr = 0; rr < 8; rr++) {
				const int cc = (int)(c8 >> (8 * rr) & 0xff);
				if ((among >> rr & 1) && cc > bc) {
					bc = cc;
					r = rr;
				}
			}
			return r;
		};
		for (int st = 0; st < maxq; st++) {
			unsigned used = 0, todo = 0;
			unsigned holder = 0xffffffffu;	/* nibble r: lane slot holding class r */
			unsigned choice = 0;		/* nibble b: class of lane slot b */
#pragma unroll
			for (int b = 0; b < 8; b++)
				todo |= (unsigned)(st < lquota[b]) << b;
			const unsigned active = todo;
			unsigned nopt = 0;		/* nibble b: classes lane slot b has left */
#pragma unroll
			for (int b = 0; b < 8; b++)
				nopt |= (unsigned)__popc(classes(counts(b))) << (4 * b);
			for (int round = 0; round < 8 && todo; round++) {
				/* the lane with the fewest classes left */
				int b = 0, bn = 99;
#pragma unroll
				for (int bb = 0; bb < 8; bb++) {
					const int n = (int)(nopt >> (4 * bb) & 0xfu);
					if ((todo >> bb & 1) && n < bn) {
						bn = n;
						b = bb;
					}
				}
				todo &= ~(1u << b);
				const unsigned long long c8 = counts(b);
				const unsigned av = classes(c8);
				int r;
				if (av & ~used) {
					r = best(c8, av & ~used);
					used |= 1u << r;
					holder = (holder & ~(0xfu << (4 * r))) | (unsigned)b << (4 * r);
				} else {
					/* every class of this lane is taken: ask a holder to move */
					unsigned tr = av;
					r = -1;
					for (int k = 0; k < 8 && tr && r < 0; k++) {
						const int rr = best(c8, tr);
						tr &= ~(1u << rr);
						const unsigned h = holder >> (4 * rr) & 0xfu;
						if (h == 0xfu)
							continue;
						const unsigned long long h8 = counts((int)h);
						const unsigned alt = classes(h8) & ~used;
						if (!alt)
							continue;
						const int r2 = best(h8, alt);
						used |= 1u << r2;
						holder = (holder & ~(0xfu << (4 * r2))) | h << (4 * r2);
						choice = (choice & ~(0xfu << (4 * h))) | (unsigned)r2 << (4 * h);
						holder = (holder & ~(0xfu << (4 * rr))) | (unsigned)b << (4 * rr);
						r = rr;
					}
					if (r < 0)
						r = best(c8, av);	/* a wavefront more */
				}
				choice = (choice & ~(0xfu << (4 * b))) | (unsigned)r << (4 * b);
			}
#pragma unroll
			for (int b = 0; b < 8; b++) {
				if (!(active >> b & 1))
					continue;
				const int ln = b * A3_NQ + qw, r = (int)(choice >> (4 * b) & 0xfu);
				const int from = lptr[ln * 8 + r]++;
				lcnt[ln * 8 + r]--;
				/* entry st of the lane: half (st & 1) of word (st / 2) * A3_THREADS + thread */
				out[((st >> 1) * A3_THREADS + a3_lane_thread(ln)) * 2 + (st & 1)]
					= tmp[lstart[b] + from * lS[b] + lseg[b]];
			}
		}
	}
}
