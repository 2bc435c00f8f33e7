/*
 * mc_kernels.cuh -- the small sm_100a kernels of the MULTICLUST EM hot path:
 * tails of the E/M steps, reductions, parameter-space kernels of the
 * acceleration schemes, projections, initialisers, result derivation, the
 * synthetic-data generator.  Included by mc_cuda.cu only (the definitions are
 * not inline).  The genotype-streaming kernels live in mc_admix3.cuh,
 * mc_dense.cuh and mc_tile.cuh and are instantiated in the mc_inst_*.cu files.
 */
#pragma once

#include "mc_device.cuh"

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
	int do_proj, double lb, const double *S, double *eta_t)
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
	/* pooled eta (em_alg.c:604-648, 916-962): eta_k = S_k / sum S, projection;
	 * one thread of the last block, S == nullptr with per-individual eta */
	if (S && blockIdx.x == gridDim.x - 1 && threadIdx.x == blockDim.x - 1) {
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
	int zero_skip, const int *run_if)
{
	if (run_if && !*run_if)		/* fall-back pass that is not needed */
		return;
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
/* parametric bootstrap (bootstrap.c:77-175)                              */

/* the raw draws of k_rand_assign's stream: out[d] = rand() of draw d */
__global__ void k_rand_raw(const unsigned *hist, long long n_blocks, long long block_draws,
	long long n, unsigned *out_all)
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
	unsigned *out = out_all + first;
	long long done = 0;
	for (; done + MC_RAND_ROUND <= cnt; done += MC_RAND_ROUND) {
#pragma unroll
		for (int g = 0; g < 124; g++) {
			unsigned w[4];
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const int i = (g * 4 + q) % 31;
				h[i] += h[(i + 28) % 31];
				w[q] = h[i] >> 1;
			}
			*reinterpret_cast<uint4 *>(out + done + g * 4) = make_uint4(w[0], w[1], w[2], w[3]);
		}
	}
	if (done < cnt) {
		unsigned t[31];
#pragma unroll
		for (int j = 0; j < 31; j++)
			t[j] = h[j];
		int f = 0;
		for (; done < cnt; done++) {
			t[f] += t[f >= 3 ? f - 3 : f + 28];
			out[done] = t[f] >> 1;
			if (++f == 31)
				f = 0;
		}
	}
}

/* inverse-CDF walk of bootstrap.c:96-105, 109-114: the first index whose
 * running sum reaches r, the last one if none does */
__device__ __forceinline__ int boot_pick(const double *w, int n, double r)
{
	int j = 0;
	double sum = 0.0;
	while (j < n && r > sum)
		sum = __dadd_rn(sum, w[j++]);
	return j ? j - 1 : 0;
}

/* one thread per (individual, locus) of rows [i0, i1): the P allele copies of
 * the bootstrap sample.  draws[] holds the stream from draw `d0` on. */
__global__ void k_bootstrap_codes(const unsigned *draws, long long d0, long long i0,
	long long i1, int L, int P, int K, long long T, const int *off, const int *J,
	const double *eta, long long eta_stride, const double *p, int admixture,
	unsigned char *nat)
{
	const long long n = (i1 - i0) * L;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const long long i = i0 + x / L;
		const int l = (int)(x % L);
		const int Jl = J[l];
		unsigned char *out = nat + ((size_t)i * L + l) * P;
		const long long per = (long long)L * P;
		int k = 0;
		long long base;
		if (admixture) {
			base = 2 * (i * per + (long long)l * P);
		} else {
			const long long bi = i * (1 + per);
			k = boot_pick(eta, K, __ddiv_rn((double)draws[bi - d0], 2147483647.0));
			base = bi + 1 + (long long)l * P;
		}
		for (int a = 0; a < P; a++) {
			double r;
			if (admixture) {
				r = __ddiv_rn((double)draws[base + 2 * a - d0], 2147483647.0);
				k = boot_pick(eta + (size_t)i * eta_stride, K, r);
				r = __ddiv_rn((double)draws[base + 2 * a + 1 - d0], 2147483647.0);
			} else {
				r = __ddiv_rn((double)draws[base + a - d0], 2147483647.0);
			}
			out[a] = Jl > 0 ? (unsigned char)boot_pick(p + (size_t)k * T + off[l], Jl, r)
				: (unsigned char)MC_MISSING;
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
