// Does SHFL share the shared-memory data path?  SHFL.IDX alone, LDS.64 alone, both interleaved.
// Build: nvcc -arch=sm_100a -O3 -o shfl_probe shfl_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>	// 1 = shfl, 2 = lds, 3 = both
__global__ void k(int iters, double *out, long long *cyc, int seed)
{
	extern __shared__ double sm[];
	for (int x = threadIdx.x; x < 8192; x += blockDim.x)
		sm[x] = x;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	unsigned src = (lane * 7 + seed) & 31, idx = (lane * 13 + seed) & 8191;
	unsigned a0 = lane, a1 = lane + 1, a2 = lane + 2, a3 = lane + 3;
	double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < 4; u++) {
			if (MODE & 1) {
				a0 = __shfl_sync(0xffffffffu, a0, src) + 1;
				a1 = __shfl_sync(0xffffffffu, a1, src ^ 1) + 1;
				a2 = __shfl_sync(0xffffffffu, a2, src ^ 2) + 1;
				a3 = __shfl_sync(0xffffffffu, a3, src ^ 3) + 1;
			}
			if (MODE & 2) {
				d0 += sm[(idx + u * 32) & 8191];
				d1 += sm[(idx + u * 32 + 512) & 8191];
				d2 += sm[(idx + u * 32 + 1024) & 8191];
				d3 += sm[(idx + u * 32 + 1536) & 8191];
			}
		}
		idx = (idx + 128) & 8191;
	}
	long long t1 = clock64();
	out[threadIdx.x] = d0 + d1 + d2 + d3 + a0 + a1 + a2 + a3;
	if (threadIdx.x == 0)
		cyc[0] = t1 - t0;
}

int main()
{
	double *d_out; long long *d_cyc, c;
	cudaMalloc(&d_out, sizeof(double) * 1024);
	cudaMalloc(&d_cyc, 64);
	cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
	cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
	cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
	const int iters = 20000;
	for (int warps : { 8, 16, 32 }) {
		double r[4] = { 0, 0, 0, 0 };
		for (int mode = 1; mode <= 3; mode++) {
			for (int rep = 0; rep < 2; rep++) {
				if (mode == 1) k<1><<<1, warps * 32, 65536>>>(iters, d_out, d_cyc, 3);
				if (mode == 2) k<2><<<1, warps * 32, 65536>>>(iters, d_out, d_cyc, 3);
				if (mode == 3) k<3><<<1, warps * 32, 65536>>>(iters, d_out, d_cyc, 3);
				cudaDeviceSynchronize();
			}
			cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
			r[mode] = (double)c / iters / 16 / warps;
		}
		printf("warps=%2d  clk per warp-op per SM: SHFL.IDX alone %.2f | LDS.64 alone %.2f | 1 SHFL + 1 LDS.64 together %.2f\n",
			warps, r[1], r[2], r[3]);
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
