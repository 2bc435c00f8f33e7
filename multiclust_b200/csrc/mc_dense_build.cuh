/*
 * mc_dense_build.cuh -- layout builders and the mixture tail of the dense DMMA
 * path (packed counts, dense p table).  Included by mc_cuda.cu only.
 */
#pragma once

#include "mc_dense.cuh"

/* ---------------------------------------------------------------------- */
/* one-time layout builders                                                 */

/* largest allele code in the data (255 = missing is skipped) */
__global__ void k_dense_maxcode(const unsigned char *nat, long long n, unsigned *out)
{
	unsigned m = 0;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const unsigned c = nat[x];
		if (c != 255u && c > m)
			m = c;
	}
	for (int s = 16; s >= 1; s >>= 1)
		m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
	if ((threadIdx.x & 31) == 0 && m)
		atomicMax(out, m);
}

/* natural [I][L][P] codes -> c0 | c1 << 4 per (individual, locus), tile major */
__global__ void k_dense_counts(const unsigned char *nat, unsigned char *cnt,
	long long I, int L, int P, int n_itiles, int n_ltiles)
{
	const long long n = (long long)n_itiles * n_ltiles * DN_IT;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int row = (int)(x % DN_IT);
		const long long tix = x / DN_IT;
		const int lt = (int)(tix % n_ltiles);
		const long long i = (tix / n_ltiles) * DN_IT + row;
		unsigned w[4] = { 0u, 0u, 0u, 0u };
		if (i < I)
			for (int s = 0; s < DN_TL; s++) {
				const int l = lt * DN_TL + s;
				if (l >= L)
					break;
				const unsigned char *c = nat + ((size_t)i * L + l) * P;
				unsigned c0 = 0, c1 = 0;
				for (int a = 0; a < P; a++) {
					c0 += c[a] == 0;
					c1 += c[a] == 1;
				}
				w[s >> 2] |= (c0 | c1 << 4) << ((s & 3) * 8);
			}
		*reinterpret_cast<uint4 *>(cnt + (size_t)x * 16) = make_uint4(w[0], w[1], w[2], w[3]);
	}
}

/* p [K][T] -> dense [locus][k][allele] with the fragment pitch.  take_log: 0 =
 * the table as it is; 1 = log p with the E-step's "p == 0 contributes nothing"
 * rule (em_alg.c:797-804); 2 = plain log p (log_likelihood.c:190-199), where
 * minus infinity (log 0, only without the projection) becomes -DBL_MAX so that
 * a zero count still contributes zero */
__global__ void k_dense_p(const double *p, double *pd, const int *off, const int *J,
	int K, int L, long long T, int n_loci_pad, int PL, int K8, int take_log,
	const int *run_if)
{
	if (run_if && !*run_if)
		return;
	const long long n = (long long)n_loci_pad * K8 * 2;
	for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n;
		x += (long long)gridDim.x * blockDim.x) {
		const int a = (int)(x & 1), k = (int)((x >> 1) % K8);
		const int l = (int)(x / (2 * K8));
		double v = 0.0;
		if (l < L && k < K && a < J[l]) {
			v = p[(size_t)k * T + off[l] + a];
			if (take_log) {
				v = (take_log == 1 && v == 0.0) ? 0.0 : log(v);
				if (v == -INFINITY)
					v = -1.7976931348623157e308;
			}
		}
		pd[(size_t)l * PL + 2 * k + a] = v;
	}
}

/* ---------------------------------------------------------------------- */
/* mixture tail: a_ik = log eta_k + sum over locus chunks, then the softmax
 * of the E-step (em_alg.c:828-882) or the guarded log-sum-exp of
 * logL_mixture (log_likelihood.c:203-228).  One thread per (individual, k)
 * adds the chunk partial sums (a warp reads contiguous runs of Apart), one
 * thread per individual finishes the row. */
#define MT_ROWS 32
/* ... and the sums the step needs over individuals, fused in: every block adds
 * up the log likelihood and (E-step) the K posterior columns of its rows in row
 * order and leaves them in part[block][2 K + 1] = {S_0..S_{K-1}, ll, max_i v_i0..};
 * k_mix_final adds the blocks in block order.  Fixed order, no atomics. */
__global__ void k_mix_tail(const double *Apart, int n_chunks, const int *n_chunks_dev,
	long long Ipad, long long I, int K, const double *eta, double *vik, double *part,
	int ll_only)
{
	if (n_chunks_dev)	/* left by whichever kernel ran (mc_digit.cuh) */
		n_chunks = *n_chunks_dev;
	extern __shared__ double mt_rows[];	/* [MT_ROWS][K] then [MT_ROWS] */
	double *mt_ll = mt_rows + MT_ROWS * K;
	const int n = MT_ROWS * K;
	double acc_col = 0.0;			/* thread k < K: S_k; thread K: ll */
	double max_col = 0.0;			/* thread k < K: max_i v_ik */
	for (long long i0 = (long long)blockIdx.x * MT_ROWS; i0 < I; i0 += (long long)gridDim.x * MT_ROWS) {
		for (int x = threadIdx.x; x < n; x += blockDim.x) {
			const long long i = i0 + x / K;
			if (i >= I)
				continue;
			const double *src = Apart + (size_t)i0 * K + x;
			const size_t ts = (size_t)Ipad * K;
			double acc = 0.0;
			for (int c = 0; c < n_chunks; c++)
				acc += __ldg(src + (size_t)c * ts);
			mt_rows[x] = acc + log(eta[x % K]);
		}
		__syncthreads();
		if (threadIdx.x < MT_ROWS && i0 + threadIdx.x < I) {
			const long long i = i0 + threadIdx.x;
			double *v = mt_rows + threadIdx.x * K;
			double mx = -INFINITY;
			for (int k = 0; k < K; k++)
				mx = v[k] > mx ? v[k] : mx;
			if (!ll_only) {
				double s = 0.0;
				for (int k = 0; k < K; k++)
					s += exp(v[k] - mx);
				for (int k = 0; k < K; k++) {
					const double vv = exp(v[k] - mx) / s;
					vik[(size_t)i * K + k] = vv;
					v[k] = vv;
				}
				mt_ll[threadIdx.x] = log(s) + mx;
			} else {
				double te = exp(mx), scale = 0.0, s = 0.0;
				if (te == 0.0 || te == HUGE_VAL) {
					scale = (te == HUGE_VAL) ? mx : -mx;
					do {
						scale *= 0.5;
						te = exp(scale);
					} while (te == HUGE_VAL);
					scale = mx - scale;
				}
				for (int k = 0; k < K; k++)
					s += exp(v[k] - scale);
				mt_ll[threadIdx.x] = log(s) + scale;
			}
		}
		__syncthreads();
		const int rows = (int)(I - i0 < MT_ROWS ? I - i0 : MT_ROWS);
		if ((int)threadIdx.x < K && !ll_only) {
			for (int r = 0; r < rows; r++) {
				const double vv = mt_rows[r * K + threadIdx.x];
				acc_col += vv;
				max_col = fmax(max_col, vv);
			}
		} else if ((int)threadIdx.x == K) {
			for (int r = 0; r < rows; r++)
				acc_col += mt_ll[r];
		}
		__syncthreads();
	}
	if ((int)threadIdx.x <= K)
		part[(size_t)blockIdx.x * (2 * K + 1) + threadIdx.x] = acc_col;
	if ((int)threadIdx.x < K)
		part[(size_t)blockIdx.x * (2 * K + 1) + K + 1 + threadIdx.x] = max_col;
}

/* out_ll = sum_b part[b][K]; out_S[k] = sum_b part[b][k] (E-step only).  vscale
 * (nullable, E-step only): the power-of-two scaling of the posterior column k
 * for the digit-sliced M pass, {2^(64 - e_k)} then {2^(e_k - 64)} with
 * max_i v_ik < 2^e_k (mc_digit.cuh). */
__global__ void k_mix_final(const double *part, int blocks, int K, double *out_ll,
	double *out_S, int ll_only, int *flag, double *vscale)
{
	/* one warp per column: lanes stride over the blocks, then a fixed tree */
	const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
	for (int x = threadIdx.x >> 5; x <= K; x += nw) {
		if (ll_only && x < K)
			continue;
		double s = 0.0, m = 0.0;
		for (int b = lane; b < blocks; b += 32) {
			s += part[(size_t)b * (2 * K + 1) + x];
			if (x < K && vscale)
				m = fmax(m, part[(size_t)b * (2 * K + 1) + K + 1 + x]);
		}
		for (int o = 16; o >= 1; o >>= 1) {
			s += shfl_xor_f64(s, o);
			m = fmax(m, shfl_xor_f64(m, o));
		}
		if (lane)
			continue;
		if (x < K && vscale) {
			int e = 0;
			if (m > 0.0 && m <= 1.0)
				frexp(m, &e);		/* m = f 2^e, 0.5 <= f < 1 */
			e = e < -900 ? -900 : e;
			vscale[x] = scalbn(1.0, 64 - e);
			vscale[K + x] = scalbn(1.0, e - 64);
		}
		if (x == K) {
			/* the digit table's flag ends with the pass.  The log-likelihood
			 * pass fell back to the FP64 kernels on it; the E-step has no
			 * fall-back -- only a NaN or a negative p raises it there -- and
			 * reports NaN */
			if (flag) {
				if (flag[0] && !ll_only)
					s = __longlong_as_double(0x7ff8000000000000LL);
				flag[0] = 0;
			}
			*out_ll = s;
		} else {
			out_S[x] = s;
		}
	}
}
