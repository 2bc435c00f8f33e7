/*
 * mc_rand.h -- the random stream of the initialisers as an explicit object.
 *
 * The reference draws every initialisation from glibc rand() -- one stream
 * shared by all initialisations and all K (multiclust.c:516-531; srand only
 * with -r, multiclust.c:1592-1596) -- so bit-exact initial parameters need
 * that stream.  glibc's rand() is the TYPE_3 additive feedback generator
 * r[i] = r[i-3] + r[i-31] (published in glibc's random_r.c and in many
 * descriptions of it); restating it here with its state in a struct makes the
 * stream a value that can be copied: the sharded multi-start mode snapshots
 * the state in front of every fit and lets each device regenerate its own
 * draws.  tests/test_host_rand.py checks the
 * sequence against the C library for several seeds.
 */
#ifndef MC_RAND_H
#define MC_RAND_H

#include <stdint.h>

typedef struct {
	int32_t r[34];
	int f, b;	/* front / rear positions of the lag-(3, 31) recurrence */
} mcr_state;

void mcr_seed(mcr_state *s, unsigned int seed);		/* srand(seed) */
static inline int mcr_next(mcr_state *s)		/* rand() */
{
	int32_t *r = s->r;
	const uint32_t v = (uint32_t)r[s->f] + (uint32_t)r[s->b];
	r[s->f] = (int32_t)v;
	if (++s->f >= 31)
		s->f = 0;
	if (++s->b >= 31)
		s->b = 0;
	return (int)(v >> 1);
}

#endif
