// Second probe: sustained shared-memory load throughput per SM for 32/64/128-bit conflict-free
// loads whose addresses change every iteration (no hoisting).
#include <cstdio>
#include <cuda_runtime.h>

template <int W>
__global__ void probe(int iters, double *out, long long *cyc, int stride)
{
	extern __shared__ unsigned char sm[];
	for (int x = threadIdx.x; x < 65536 / 4; x += blockDim.x)
		reinterpret_cast<float *>(sm)[x] = 1.0f;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned off = (lane * W * stride + warp * 1024) & 65535;
	double acc = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			const unsigned o = (off + u * 4096) & (65535 & ~(W - 1));
			if (W == 4) acc += *reinterpret_cast<float *>(sm + o);
			if (W == 8) acc += *reinterpret_cast<double *>(sm + o);
			if (W == 16) { double2 v = *reinterpret_cast<double2 *>(sm + o); acc += v.x + v.y; }
		}
		off = (off + 512 + (it & 1) * 512) & 65535;
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0)
		cyc[blockIdx.x] = t1 - t0;
}

template <int W> void run(const char *name, int stride)
{
	double *d_out; long long *d_cyc, c;
	cudaMalloc(&d_out, sizeof(double) * 2048);
	cudaMalloc(&d_cyc, sizeof(long long) * 8);
	cudaFuncSetAttribute(probe<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
	const int iters = 20000;
	for (int warps : { 1, 4, 8, 16, 32 }) {
		probe<W><<<1, warps * 32, 65536>>>(iters, d_out, d_cyc, stride);
		cudaDeviceSynchronize();
		probe<W><<<1, warps * 32, 65536>>>(iters, d_out, d_cyc, stride);
		cudaDeviceSynchronize();
		cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
		const double per = (double)c / iters / 8 / warps;
		printf("%-28s warps=%2d  %.2f clk per warp-load  -> %.0f B/clk/SM\n", name, warps, per, 32.0 * W / per);
	}
	cudaFree(d_out); cudaFree(d_cyc);
}

int main()
{
	run<4>("LDS.32 coalesced", 1);
	run<8>("LDS.64 coalesced", 1);
	run<16>("LDS.128 coalesced", 1);
	run<16>("LDS.128 stride 5 (80 B rows)", 5);
	run<8>("LDS.64 stride 5", 5);
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
