/*
 * mc_device.cuh -- device helpers shared by every kernel header.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define MC_MISSING 255

__device__ __forceinline__ double mc_rcp(double x)
{
#ifdef MC_EXACT_DIV
	return 1.0 / x;
#else
	/* MUFU.RCP64H seed + two Newton steps: ~1 ulp, a third of the
	 * instructions of the IEEE division sequence */
	double r, e;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	return r;
#endif
}

__device__ __forceinline__ double shfl_xor_f64(double v, int mask)
{
	int lo = __double2loint(v), hi = __double2hiint(v);
	lo = __shfl_xor_sync(0xffffffffu, lo, mask);
	hi = __shfl_xor_sync(0xffffffffu, hi, mask);
	return __hiloint2double(hi, lo);
}

