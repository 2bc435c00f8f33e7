// Integer tensor-path probe: what mma.sync (legacy warp-level path) delivers on B200 for
// u8 x u8 -> s32 (IMMA m16n8k32) and f16 x f16 -> f32 (HMMA m16n8k16), per SM and per
// warp instruction.  The mixture model's contractions on biallelic data are
// (small integer counts) x (FP64 table); with the table cut into 8-bit digits they become
// exact integer GEMMs ("digit-sliced mixture path").  This is the rate half of
// that decision, next to tools/dmma_probe.cu (profiles/r02_imma_probe.txt).
//   imma     U independent accumulator chains of mma.sync.m16n8k32.s32.u8.u8.s32
//   hmma     U independent chains of mma.sync.m16n8k16.f32.f16.f16.f32
//   imma+alu IMMA with the nibble unpack (2 LOP3 + 1 SHF per A register) interleaved
// Build: nvcc -arch=sm_100a -O3 -o imma_probe imma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int *c, const unsigned *a, const unsigned *b)
{
	asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
		: "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
		: "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void hmma(float *c, const unsigned *a, const unsigned *b)
{
	asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
		: "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
		: "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// KIND 0: IMMA, 1: HMMA, 2: IMMA + unpack ALU work
template <int KIND, int U>
__global__ void probe(int iters, int *out, long long *cyc, unsigned seed)
{
	int ci[U][4];
	float cf[U][4];
	unsigned a[4], b[2], packed[2];
#pragma unroll
	for (int u = 0; u < U; u++)
#pragma unroll
		for (int j = 0; j < 4; j++) {
			ci[u][j] = threadIdx.x + u + j;
			cf[u][j] = threadIdx.x * 1e-3f + u + j;
		}
#pragma unroll
	for (int j = 0; j < 4; j++)
		a[j] = KIND == 1 ? 0x3c003c00u : 0x01020100u + threadIdx.x % 2;
	b[0] = KIND == 1 ? 0x3c003c00u : 0x7f017f01u;
	b[1] = KIND == 1 ? 0x38003c00u : 0x017f017fu;
	packed[0] = seed * 0x9e3779b9u + threadIdx.x;
	packed[1] = seed * 0x85ebca6bu + threadIdx.x;
	__syncthreads();
	const long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < U; u++) {
			if (KIND == 2) {
				// one packed word (c0 | c1 << 4 per byte) -> two A registers
				const unsigned w = packed[u & 1] + (unsigned)it;
				a[(2 * u) & 3] = w & 0x0f0f0f0fu;
				a[(2 * u + 1) & 3] = (w >> 4) & 0x0f0f0f0fu;
			}
			if (KIND == 1)
				hmma(cf[u], a, b);
			else
				imma(ci[u], a, b);
		}
	}
	const long long t1 = clock64();
	int s = 0;
#pragma unroll
	for (int u = 0; u < U; u++)
#pragma unroll
		for (int j = 0; j < 4; j++)
			s += ci[u][j] + (int)cf[u][j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		cyc[0] = t1 - t0;
}

template <int KIND, int U>
static void run(const char *name, int warps, double mac_per_inst)
{
	int *out;
	long long *cyc, h;
	const int iters = 4096;
	cudaMalloc(&out, sizeof(int) * 148 * 1024);
	cudaMalloc(&cyc, sizeof(long long));
	probe<KIND, U><<<148, warps * 32>>>(64, out, cyc, 1);
	probe<KIND, U><<<148, warps * 32>>>(iters, out, cyc, 2);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) {
		printf("%s: %s\n", name, cudaGetErrorString(e));
		exit(1);
	}
	cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	const double per_iter = (double)h / iters;
	const double insts = (double)warps * U;
	printf("%-44s warps=%2d %9.2f clk/iter %9.1f MAC/clk/SM %7.2f clk per warp-inst/SM\n",
		name, warps, per_iter, insts * mac_per_inst / per_iter, per_iter / insts);
	cudaFree(out);
	cudaFree(cyc);
}

int main()
{
	const int ws[] = {4, 8, 16};
	for (int w : ws) {
		run<0, 8>("IMMA m16n8k32 u8.u8 x8 chains", w, 16.0 * 8 * 32);
		run<0, 2>("IMMA m16n8k32 u8.u8 x2 chains (latency)", w, 16.0 * 8 * 32);
		run<2, 8>("IMMA m16n8k32 + nibble unpack x8", w, 16.0 * 8 * 32);
		run<1, 8>("HMMA m16n8k16 f16 -> f32 x8 chains", w, 16.0 * 8 * 16);
		run<1, 2>("HMMA m16n8k16 f16 -> f32 x2 chains (latency)", w, 16.0 * 8 * 16);
	}
	return 0;
}
