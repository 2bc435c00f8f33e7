/*
 * ref_hooks.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Force-included (gcc -include) when oracle/Makefile compiles the reference's
 * em_alg.c in place from /root/reference.  The reference only prints the
 * per-iteration log likelihood with "%.2f" (em_alg.c:123-136), which is far
 * too coarse for a 1e-9 relative parity check.  stop() tests isnan(loglik)
 * before anything else (em_alg.c:106), so redefining isnan for that one
 * translation unit lets the harness record every log likelihood handed to
 * stop() at full precision without touching the reference sources.
 */
#ifndef REF_HOOKS_H
#define REF_HOOKS_H

#include <math.h>

void ref_hook_ll(double ll);

#undef isnan
#define isnan(x) (ref_hook_ll((double)(x)), __builtin_isnan(x))

#endif
