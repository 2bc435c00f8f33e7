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
