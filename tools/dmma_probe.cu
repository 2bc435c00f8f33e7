// DMMA probe: what the FP64 tensor path (mma.sync ... f64) delivers on B200 next to the
// FP64 vector pipe, and whether the two overlap.  BASELINE.json's north star sends the dense
// contractions (mixture one-hot(X) . log p^T, biallelic admixture) to DMMA only if it beats
// the SIMT path; this is the rate half of that decision (profiles/r02_dmma_probe.txt).
//   dmma<SHAPE>   U independent accumulator chains of mma.sync.m8n8k4 / m16n8k4 / m16n8k8 /
//                 m16n8k16 (f64), operands in registers
//   dfma          8 independent DFMA chains (reference rate, tools/lds_probe3.cu measured 0.50)
//   mixed         DMMA and DFMA chains interleaved in one warp: separate pipes would add up
//   lds+dmma      one LDS.64 operand fetch per DMMA (the fragment-from-shared-memory pattern)
// Build: nvcc -arch=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
		: "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double *c, const double *a, double b)
{
	asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
		: "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double *c, const double *a, const double *b)
{
	asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
		: "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
		: "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double *c, const double *a, const double *b)
{
	asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
		"{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
		: "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
		: "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
		  "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// SHAPE 0: m8n8k4 (256 FMA), 1: m16n8k4 (512), 2: m16n8k8 (1024), 3: m16n8k16 (2048)
template <int SHAPE, int U>
__global__ void dmma(int iters, double *out, long long *cyc)
{
	double c[U][4], a[8], b[4];
#pragma unroll
	for (int u = 0; u < U; u++)
#pragma unroll
		for (int j = 0; j < 4; j++)
			c[u][j] = threadIdx.x * 1e-3 + u + j;
#pragma unroll
	for (int j = 0; j < 8; j++)
		a[j] = 1.0 + 1e-9 * (threadIdx.x + j);
#pragma unroll
	for (int j = 0; j < 4; j++)
		b[j] = 1.0 - 1e-9 * (threadIdx.x + j);
	__syncthreads();
	const long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < U; u++) {
			if (SHAPE == 0) mma884(c[u][0], c[u][1], a[0], b[0]);
			if (SHAPE == 1) mma1684(c[u], a, b[0]);
			if (SHAPE == 2) mma1688(c[u], a, b);
			if (SHAPE == 3) mma16816(c[u], a, b);
		}
	}
	const long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int u = 0; u < U; u++)
#pragma unroll
		for (int j = 0; j < 4; j++)
			s += c[u][j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		cyc[0] = t1 - t0;
}

// ND m8n8k4 + NF DFMA per inner step, all chains independent
template <int ND, int NF>
__global__ void mixed(int iters, double *out, long long *cyc, double x)
{
	double c[ND > 0 ? ND : 1][2], f[NF > 0 ? NF : 1];
#pragma unroll
	for (int u = 0; u < ND; u++)
		c[u][0] = c[u][1] = threadIdx.x * 1e-3 + u;
#pragma unroll
	for (int u = 0; u < NF; u++)
		f[u] = threadIdx.x + u;
	const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
	const double m = x + 1e-9 * threadIdx.x;
	__syncthreads();
	const long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
#pragma unroll
			for (int u = 0; u < ND; u++)
				mma884(c[u][0], c[u][1], a, b);
#pragma unroll
			for (int u = 0; u < NF; u++)
				f[u] = fma(f[u], m, 1e-3);
		}
	}
	const long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int u = 0; u < ND; u++)
		s += c[u][0] + c[u][1];
#pragma unroll
	for (int u = 0; u < NF; u++)
		s += f[u];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		cyc[0] = t1 - t0;
}

// one conflict-free LDS.64 (the B fragment) per m8n8k4, A fragment resident
template <int U>
__global__ void lds_dmma(int iters, double *out, long long *cyc)
{
	extern __shared__ double sm[];
	for (int x = threadIdx.x; x < 4096; x += blockDim.x)
		sm[x] = 1.0 + 1e-9 * x;
	__syncthreads();
	double c[U][2];
#pragma unroll
	for (int u = 0; u < U; u++)
		c[u][0] = c[u][1] = threadIdx.x * 1e-3 + u;
	const double a = 1.0 + 1e-9 * threadIdx.x;
	const int lane = threadIdx.x & 31;
	/* B fragment: element (row lane%4, col lane/4) of a tile with a pitch of 36 doubles */
	const double *bp = sm + (lane & 3) * 36 + (lane >> 2);
	unsigned rot = 0;
	const long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < U; u++) {
			const double b = bp[(rot + u * 144) & 2047];
			mma884(c[u][0], c[u][1], a, b);
		}
		rot += 8;
	}
	const long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int u = 0; u < U; u++)
		s += c[u][0] + c[u][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		cyc[0] = t1 - t0;
}

static double *d_out;
static long long *d_cyc;

template <typename F> static void run(const char *name, int warps, double fma_per_warp_iter,
	double inst_per_warp_iter, F launch)
{
	const int iters = 2000;
	launch(warps * 32, 10);	/* warm-up */
	cudaDeviceSynchronize();
	launch(warps * 32, iters);
	cudaError_t e = cudaDeviceSynchronize();
	long long cyc = 0;
	cudaMemcpy(&cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost);
	if (e != cudaSuccess) {
		printf("%-44s warps=%2d  %s\n", name, warps, cudaGetErrorString(e));
		return;
	}
	const double clk = (double)cyc / iters;
	printf("%-44s warps=%2d  %8.2f clk/iter  %7.1f FMA/clk/SM  %6.2f clk per warp-inst/SM\n",
		name, warps, clk, fma_per_warp_iter * warps / clk, clk / (inst_per_warp_iter * warps));
}

int main()
{
	cudaMalloc(&d_out, sizeof(double) * 1024 * 148);
	cudaMalloc(&d_cyc, sizeof(long long));
	const int ws[] = { 4, 8, 16, 32 };
	for (int w : ws) {
		run("DMMA m8n8k4  x8 chains", w, 8 * 256.0, 8, [&](int th, int it) {
			dmma<0, 8><<<148, th>>>(it, d_out, d_cyc); });
		run("DMMA m16n8k4 x8 chains", w, 8 * 512.0, 8, [&](int th, int it) {
			dmma<1, 8><<<148, th>>>(it, d_out, d_cyc); });
		run("DMMA m16n8k8 x8 chains", w, 8 * 1024.0, 8, [&](int th, int it) {
			dmma<2, 8><<<148, th>>>(it, d_out, d_cyc); });
		run("DMMA m16n8k16 x8 chains", w, 8 * 2048.0, 8, [&](int th, int it) {
			dmma<3, 8><<<148, th>>>(it, d_out, d_cyc); });
		run("DMMA m8n8k4  x2 chains (latency)", w, 2 * 256.0, 2, [&](int th, int it) {
			dmma<0, 2><<<148, th>>>(it, d_out, d_cyc); });
		run("DFMA only (8 chains x4)", w, 32 * 32.0, 32, [&](int th, int it) {
			mixed<0, 8><<<148, th>>>(it, d_out, d_cyc, 1.0000001); });
		run("DMMA only (4 chains x4)", w, 16 * 256.0, 16, [&](int th, int it) {
			mixed<4, 0><<<148, th>>>(it, d_out, d_cyc, 1.0000001); });
		run("mixed 4 DMMA + 8 DFMA (x4)", w, 16 * 256.0 + 32 * 32.0, 48, [&](int th, int it) {
			mixed<4, 8><<<148, th>>>(it, d_out, d_cyc, 1.0000001); });
		run("mixed 2 DMMA + 16 DFMA (x4)", w, 8 * 256.0 + 64 * 32.0, 72, [&](int th, int it) {
			mixed<2, 16><<<148, th>>>(it, d_out, d_cyc, 1.0000001); });
		run("LDS.64 B fragment + DMMA m8n8k4 x8", w, 8 * 256.0, 8, [&](int th, int it) {
			lds_dmma<8><<<148, th, 4096 * sizeof(double)>>>(it, d_out, d_cyc); });
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
