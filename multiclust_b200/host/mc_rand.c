/* mc_rand.c -- see mc_rand.h */
#include "mc_rand.h"

void mcr_seed(mcr_state *s, unsigned int seed)
{
	int32_t *r = s->r;
	int64_t word;

	if (seed == 0)
		seed = 1;
	r[0] = (int32_t)seed;
	for (int i = 1; i < 31; i++) {
		/* 16807 * r[i-1] mod (2^31 - 1), without overflow */
		const int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
		word = 16807 * lo - 2836 * hi;
		if (word < 0)
			word += 2147483647;
		r[i] = (int32_t)word;
	}
	s->f = 3;
	s->b = 0;
	for (int i = 0; i < 310; i++)	/* the generator is run 310 times */
		(void)mcr_next(s);
}

/* ---- linear-recurrence view (see mc_rand.h) ---- */

void mcr_history(const mcr_state *s, uint32_t h[MCR_LAG])
{
	/* slot f holds x[n-31], the slots after it the younger words */
	for (int j = 0; j < MCR_LAG; j++)
		h[j] = (uint32_t)s->r[(s->f + j) % MCR_LAG];
}

void mcr_from_history(mcr_state *s, const uint32_t h[MCR_LAG])
{
	for (int j = 0; j < MCR_LAG; j++)
		s->r[j] = (int32_t)h[j];
	s->f = 0;
	s->b = MCR_LAG - 3;
}

void mcr_step_history(uint32_t h[MCR_LAG], long long m)
{
	/* in place on a circular buffer, rotated back to linear order at the end */
	int f = 0;
	uint32_t t[MCR_LAG];

	for (long long n = 0; n < m; n++) {
		h[f] += h[(f + MCR_LAG - 3) % MCR_LAG];
		if (++f == MCR_LAG)
			f = 0;
	}
	for (int j = 0; j < MCR_LAG; j++)
		t[j] = h[(f + j) % MCR_LAG];
	for (int j = 0; j < MCR_LAG; j++)
		h[j] = t[j];
}

static void mat_mul(const uint32_t *A, const uint32_t *B, uint32_t *C)
{
	for (int i = 0; i < MCR_LAG; i++)
		for (int j = 0; j < MCR_LAG; j++) {
			uint32_t acc = 0;
			for (int k = 0; k < MCR_LAG; k++)
				acc += A[i * MCR_LAG + k] * B[k * MCR_LAG + j];
			C[i * MCR_LAG + j] = acc;
		}
}

void mcr_jump_matrix(long long m, uint32_t M[MCR_LAG * MCR_LAG])
{
	uint32_t P[MCR_LAG * MCR_LAG], T[MCR_LAG * MCR_LAG];

	/* one step: h'[j] = h[j+1], h'[30] = h[0] + h[28] */
	for (int i = 0; i < MCR_LAG * MCR_LAG; i++) {
		P[i] = 0;
		M[i] = 0;
	}
	for (int j = 0; j + 1 < MCR_LAG; j++)
		P[j * MCR_LAG + j + 1] = 1;
	P[(MCR_LAG - 1) * MCR_LAG + 0] = 1;
	P[(MCR_LAG - 1) * MCR_LAG + MCR_LAG - 3] = 1;
	for (int j = 0; j < MCR_LAG; j++)
		M[j * MCR_LAG + j] = 1;
	for (; m > 0; m >>= 1) {
		if (m & 1) {
			mat_mul(P, M, T);
			for (int i = 0; i < MCR_LAG * MCR_LAG; i++)
				M[i] = T[i];
		}
		mat_mul(P, P, T);
		for (int i = 0; i < MCR_LAG * MCR_LAG; i++)
			P[i] = T[i];
	}
}

void mcr_apply(const uint32_t M[MCR_LAG * MCR_LAG], uint32_t h[MCR_LAG])
{
	uint32_t t[MCR_LAG];

	for (int i = 0; i < MCR_LAG; i++) {
		uint32_t acc = 0;
		for (int k = 0; k < MCR_LAG; k++)
			acc += M[i * MCR_LAG + k] * h[k];
		t[i] = acc;
	}
	for (int i = 0; i < MCR_LAG; i++)
		h[i] = t[i];
}
