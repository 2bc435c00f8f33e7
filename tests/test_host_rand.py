"""multiclust_b200/host/mc_rand.{h,c} restates glibc's rand() (the TYPE_3
additive-feedback generator the reference's initialisers draw from,
rnd_init.c:205-217, 460-481) with its state in a struct.  The sequence is
checked here against the C library itself for several seeds, including the
no-srand() default (seed 1) and the seed-0 special case."""
import os
import subprocess

from common import ROOT

HOST = os.path.join(ROOT, "multiclust_b200", "host")

PROG = r"""
#include <stdio.h>
#include <stdlib.h>
#include "mc_rand.h"
int main(void)
{
	static const unsigned int seeds[] = { 1, 0, 2, 42, 1234567, 20261018, 4294967295u };
	mcr_state s, copy;
	/* default stream: no srand() at all */
	mcr_seed(&s, 1);
	for (int i = 0; i < 100000; i++)
		if (mcr_next(&s) != rand()) { printf("default stream differs at %d\n", i); return 1; }
	for (unsigned k = 0; k < sizeof seeds / sizeof *seeds; k++) {
		srand(seeds[k]);
		mcr_seed(&s, seeds[k]);
		for (int i = 0; i < 200000; i++) {
			if (i == 777)
				copy = s;	/* the state is a value */
			if (mcr_next(&s) != rand()) {
				printf("seed %u differs at %d\n", seeds[k], i);
				return 1;
			}
		}
		srand(seeds[k]);
		for (int i = 0; i < 777; i++)
			(void)rand();
		for (int i = 0; i < 1000; i++)
			if (mcr_next(&copy) != rand()) { printf("snapshot differs\n"); return 1; }
	}
	printf("ok\n");
	return 0;
}
"""


def test_mc_rand_equals_glibc_rand(tmp_path):
    src = tmp_path / "t.c"
    src.write_text(PROG)
    exe = str(tmp_path / "t")
    subprocess.check_call(["gcc", "-std=c17", "-O2", "-I" + HOST, "-o", exe, str(src),
                           os.path.join(HOST, "mc_rand.c")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout


JUMP = r"""
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mc_rand.h"
int main(void)
{
	static uint32_t M[MCR_LAG * MCR_LAG];
	const long long jumps[] = { 0, 1, 2, 30, 31, 32, 496, 31744, 1000003 };
	for (unsigned k = 0; k < sizeof jumps / sizeof *jumps; k++) {
		mcr_state a, b;
		uint32_t h[MCR_LAG], g[MCR_LAG];
		mcr_seed(&a, 12345);
		for (int i = 0; i < 77; i++)
			(void)mcr_next(&a);	/* an arbitrary position, f != 0 */
		b = a;
		mcr_history(&a, h);
		memcpy(g, h, sizeof g);
		mcr_jump_matrix(jumps[k], M);
		mcr_apply(M, h);		/* by matrix */
		mcr_step_history(g, jumps[k]);	/* by stepping the history */
		for (long long i = 0; i < jumps[k]; i++)
			(void)mcr_next(&b);	/* by drawing */
		if (memcmp(h, g, sizeof h)) { printf("matrix != stepping at %lld\n", jumps[k]); return 1; }
		mcr_from_history(&a, h);
		for (int i = 0; i < 100; i++)
			if (mcr_next(&a) != mcr_next(&b)) { printf("stream differs after jump %lld\n", jumps[k]); return 1; }
	}
	printf("ok\n");
	return 0;
}
"""


def test_mc_rand_jump_ahead(tmp_path):
    """the generator as a linear recurrence: advancing by a matrix power, by
    stepping the 31-word history and by drawing give the same stream"""
    src = tmp_path / "j.c"
    src.write_text(JUMP)
    exe = str(tmp_path / "j")
    subprocess.check_call(["gcc", "-std=c17", "-O2", "-I" + HOST, "-o", exe, str(src),
                           os.path.join(HOST, "mc_rand.c")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout
