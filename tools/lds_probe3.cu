// Third probe: what the two data-dependent access shapes of the admixture kernel cost on one SM.
//   pass 1  32 lanes (individuals) at one locus read the p value of THEIR allele: few distinct,
//           consecutive words  -> LDS.64 / LDS.128 with multi-lane broadcast
//   pass 2  32 lanes read 16 B of 32 unrelated eta rows -> LDS.128 gather, random vs. conflict-free
//   FP64    DFMA issue rate alone, with a constant-bank operand, and interleaved with LDS
// Build: nvcc -arch=sm_100a -O3 -o lds_probe3 lds_probe3.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int W>
__global__ void gather(const int *pat, int npat, int iters, double *out, long long *cyc)
{
	extern __shared__ unsigned char sm[];
	for (int x = threadIdx.x; x < 65536 / 8; x += blockDim.x)
		reinterpret_cast<double *>(sm)[x] = 1.0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int idx[8];
#pragma unroll
	for (int u = 0; u < 8; u++)
		idx[u] = pat[((warp * 8 + u) % npat) * 32 + lane];
	double acc = 0;
	unsigned rot = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			const unsigned o = ((unsigned)idx[u] * W + rot) & 65535u;
			if (W == 8) acc += *reinterpret_cast<double *>(sm + o);
			if (W == 16) { double2 v = *reinterpret_cast<double2 *>(sm + o); acc += v.x + v.y; }
		}
		rot = (rot + 2048) & 65535u;
	}
	long long t1 = clock64();
	out[threadIdx.x] = acc;
	if (threadIdx.x == 0)
		cyc[0] = t1 - t0;
}

__constant__ double cpar[64];

// MODE 0: 8 independent DFMA chains, register operands; 1: multiplier from the constant bank
template <int MODE>
__global__ void dfma(int iters, double *out, long long *cyc, double x)
{
	double a[8];
#pragma unroll
	for (int u = 0; u < 8; u++)
		a[u] = threadIdx.x + u;
	double m = x + threadIdx.x * 1e-9;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 4; r++)
#pragma unroll
			for (int u = 0; u < 8; u++)
				a[u] = MODE ? fma(a[u], cpar[(r * 8 + u) & 63], m) : fma(a[u], m, 1e-3);
	}
	long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int u = 0; u < 8; u++)
		s += a[u];
	out[threadIdx.x] = s;
	if (threadIdx.x == 0)
		cyc[0] = t1 - t0;
}

// pass-1 shape: per "copy" K=10 broadcast LDS.64 + 20 DFMA (tmp, then A)
__global__ void pass1_shape(const int *pat, int npat, int iters, double *out, long long *cyc)
{
	extern __shared__ unsigned char sm[];
	double *ps = reinterpret_cast<double *>(sm);	/* [k][row]: 10 x 512 */
	for (int x = threadIdx.x; x < 65536 / 8; x += blockDim.x)
		ps[x] = 1.0 / (1 + (x & 15));
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double e[10], A[10];
#pragma unroll
	for (int k = 0; k < 10; k++) {
		e[k] = 0.1 + 1e-3 * lane;
		A[k] = 0;
	}
	int row[4];
#pragma unroll
	for (int u = 0; u < 4; u++)
		row[u] = pat[((warp * 4 + u) % npat) * 32 + lane];
	unsigned rot = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < 4; u++) {
			double pr[10], s0 = 0, s1 = 0;
			const unsigned r = (row[u] + rot) & 511u;
#pragma unroll
			for (int k = 0; k < 10; k++)
				pr[k] = ps[k * 512 + r];
#pragma unroll
			for (int k = 0; k < 10; k += 2) {
				s0 = fma(e[k], pr[k], s0);
				s1 = fma(e[k + 1], pr[k + 1], s1);
			}
			const double w = 1.0 / (s0 + s1);	/* stands for the reciprocal */
#pragma unroll
			for (int k = 0; k < 10; k++)
				A[k] = fma(pr[k], w, A[k]);
		}
		rot = (rot + 32) & 511u;
	}
	long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int k = 0; k < 10; k++)
		s += A[k];
	out[threadIdx.x] = s;
	if (threadIdx.x == 0)
		cyc[0] = t1 - t0;
}

static int *d_pat;
static double *d_out;
static long long *d_cyc;

template <int W> static void run_gather(const char *name, const int *h, int npat)
{
	cudaMemcpy(d_pat, h, sizeof(int) * 32 * npat, cudaMemcpyHostToDevice);
	cudaFuncSetAttribute(gather<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
	for (int warps : { 4, 16, 32 }) {
		const int iters = 20000;
		long long c;
		for (int rep = 0; rep < 2; rep++) {
			gather<W><<<1, warps * 32, 65536>>>(d_pat, npat, iters, d_out, d_cyc);
			cudaDeviceSynchronize();
		}
		cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
		printf("%-58s LDS.%-3d warps=%2d  %.2f clk/warp-load\n", name, W * 8, warps,
			(double)c / iters / 8 / warps);
	}
}

int main()
{
	const int NP = 64;
	static int h[NP * 32];
	cudaMalloc(&d_pat, sizeof(int) * 32 * NP);
	cudaMalloc(&d_out, sizeof(double) * 2048);
	cudaMalloc(&d_cyc, sizeof(long long) * 8);
	srand(1);
	// index unit = W bytes
	for (int i = 0; i < NP * 32; i++) h[i] = 5;
	run_gather<8>("all lanes one word", h, NP);
	run_gather<16>("all lanes one word", h, NP);
	for (int J : { 2, 4, 8, 11, 16, 21 }) {
		char nm[96];
		for (int i = 0; i < NP * 32; i++) h[i] = rand() % J;
		snprintf(nm, sizeof nm, "%d distinct consecutive words, random lanes", J);
		run_gather<8>(nm, h, NP);
		run_gather<16>(nm, h, NP);
	}
	for (int i = 0; i < NP * 32; i++) h[i] = i % 32;
	run_gather<8>("coalesced", h, NP);
	run_gather<16>("coalesced", h, NP);
	// pass 2: 16-byte pieces of rows of 80 bytes (unit 16 B: row * 5)
	for (int i = 0; i < NP * 32; i++) h[i] = (rand() % 256) * 5;
	run_gather<16>("eta rows of 80 B, random rows", h, NP);
	for (int i = 0; i < NP * 32; i++) h[i] = ((rand() % 32) * 8 + (i % 8)) * 5;
	run_gather<16>("eta rows of 80 B, row%8 == lane%8", h, NP);
	for (int i = 0; i < NP * 32; i++) h[i] = (rand() % 256);
	run_gather<16>("eta [kp][i] layout, random i", h, NP);
	run_gather<8>("eta [k][i] layout, random i", h, NP);
	for (int i = 0; i < NP * 32; i++) h[i] = (rand() % 16) * 16 + (i % 16);
	run_gather<8>("eta [k][i] layout, i%16 == lane%16", h, NP);

	double hc[64];
	for (int i = 0; i < 64; i++) hc[i] = 1.0 - 1e-9 * i;
	cudaMemcpyToSymbol(cpar, hc, sizeof hc);
	for (int warps : { 4, 8, 16, 32 }) {
		const int iters = 20000;
		long long c0, c1;
		for (int rep = 0; rep < 2; rep++) { dfma<0><<<1, warps * 32>>>(iters, d_out, d_cyc, 0.999999); cudaDeviceSynchronize(); }
		cudaMemcpy(&c0, d_cyc, sizeof c0, cudaMemcpyDeviceToHost);
		for (int rep = 0; rep < 2; rep++) { dfma<1><<<1, warps * 32>>>(iters, d_out, d_cyc, 0.999999); cudaDeviceSynchronize(); }
		cudaMemcpy(&c1, d_cyc, sizeof c1, cudaMemcpyDeviceToHost);
		printf("DFMA warps=%2d  reg operands %.2f clk/warp-inst/SM   constant-bank operand %.2f\n", warps,
			(double)c0 / iters / 32 / warps, (double)c1 / iters / 32 / warps);
	}
	cudaFuncSetAttribute(pass1_shape, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
	for (int J : { 1, 11, 21 }) {
		for (int i = 0; i < NP * 32; i++) h[i] = rand() % J;
		cudaMemcpy(d_pat, h, sizeof(int) * 32 * NP, cudaMemcpyHostToDevice);
		for (int warps : { 8, 16, 32 }) {
			const int iters = 5000;
			long long c;
			for (int rep = 0; rep < 2; rep++) { pass1_shape<<<1, warps * 32, 65536>>>(d_pat, NP, iters, d_out, d_cyc); cudaDeviceSynchronize(); }
			cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
			printf("pass-1 shape J=%2d warps=%2d  %.1f clk per warp-copy (10 LDS.64 + 20 DFMA + rcp)  = %.3f clk/copy/SM\n",
				J, warps, (double)c / iters / 4 / warps, (double)c / iters / 4 / warps / 32);
		}
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
