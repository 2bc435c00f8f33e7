// Micro-benchmark: cost of LDS.128 / LDS.64 gathers for different lane -> bank-window
// assignments (which lanes form a conflict group?).  Build: nvcc -arch=sm_100a -O3 -o lds_probe lds_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe128(const int *pat, int iters, double *out, long long *cyc)
{
	extern __shared__ double2 sm2[];
	for (int x = threadIdx.x; x < 4096; x += blockDim.x)
		sm2[x] = make_double2(x, 1.0);
	__syncthreads();
	int idx = pat[threadIdx.x & 31];
	double acc = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			double2 v = sm2[(idx + u * 256) & 4095];
			acc += v.x + v.y;
		}
		idx = (idx + 8 * (int)acc * 0) & 4095;
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0)
		cyc[blockIdx.x] = t1 - t0;
}

int main()
{
	const int NP = 10;
	int h[NP][32];
	const char *name[NP];
	for (int l = 0; l < 32; l++) {
		// index unit = 16 bytes; window = idx mod 8; line = idx / 8
		name[0] = "window = lane%8 (distinct lines)";        h[0][l] = (l % 8) + 8 * l;
		name[1] = "window = lane/4 (distinct lines)";        h[1][l] = (l / 4) + 8 * l;
		name[2] = "window = (lane/8)*2 + lane%2";            h[2][l] = ((l / 8) * 2 + l % 2) + 8 * l;
		name[3] = "all lanes same address (broadcast)";     h[3][l] = 5;
		name[4] = "window = 0 for all, distinct lines";     h[4][l] = 8 * l;
		name[5] = "consecutive 16B words (coalesced)";      h[5][l] = l;
		name[6] = "window = lane%4 + 4*(lane/16)";          h[6][l] = (l % 4 + 4 * (l / 16)) + 8 * l;
		name[7] = "window = (lane%16)/2";                    h[7][l] = ((l % 16) / 2) + 8 * l;
		name[8] = "window = lane%8, lanes 0-15 only distinct, 16-31 same as l-16"; h[8][l] = (l % 8) + 8 * (l % 16);
		name[9] = "window = (lane*5)%8";                     h[9][l] = ((l * 5) % 8) + 8 * l;
	}
	int *d_pat; double *d_out; long long *d_cyc;
	cudaMalloc(&d_pat, sizeof(int) * 32);
	cudaMalloc(&d_out, sizeof(double) * 1024);
	cudaMalloc(&d_cyc, sizeof(long long) * 8);
	cudaFuncSetAttribute(probe128, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
	for (int p = 0; p < NP; p++) {
		cudaMemcpy(d_pat, h[p], sizeof(int) * 32, cudaMemcpyHostToDevice);
		long long c = 0;
		const int iters = 20000;
		probe128<<<1, 32, 65536>>>(d_pat, iters, d_out, d_cyc);
		cudaDeviceSynchronize();
		probe128<<<1, 32, 65536>>>(d_pat, iters, d_out, d_cyc);
		cudaDeviceSynchronize();
		cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
		printf("%-70s %.2f cycles per LDS.128 (1 warp)\n", name[p], (double)c / iters / 8);
		probe128<<<1, 256, 65536>>>(d_pat, iters, d_out, d_cyc);
		cudaDeviceSynchronize();
		cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
		printf("%-70s %.2f cycles per warp-LDS.128 (8 warps: throughput)\n", "", (double)c / iters / 8 / 8);
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
