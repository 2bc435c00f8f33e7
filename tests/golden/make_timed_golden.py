#!/usr/bin/env python
"""Golden output of the reference's timed repetition mode (-w n <r>, multiclust.c:201-347).

`-w` repeats the whole estimation without writing files, prints one compact line per
repetition (the model state plus elapsed seconds, converged repetitions, best log likelihood
so far) and a block of statistics over the repetitions.  This script runs the STOCK reference
binary (oracle/_ref/multiclust, compiled in place from /root/reference by oracle/Makefile) on
mc_gen data and stores its stdout; tests/test_cli_gpu.py::test_cli_timed_matches_reference
runs the product binary with the same arguments and compares everything but the clock
readings.

    python tests/golden/make_timed_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
MC_GEN = os.path.join(ROOT, "multiclust_b200", "host", "mc_gen")
REF = os.path.join(ROOT, "oracle", "_ref", "multiclust")

SMALL = dict(I=40, L=30, K=3, jmax=4, miss=300, P=2)
CASES = [
    # fixed iteration counts: nothing converges, the initialisation statistics stay 0 and
    # the reference prints its 0/0 as -nan
    ("admix_fixed", SMALL, "-a -k 3 -n 2 -T 8 -E 1e-30 -w n 3"),
    # converged fits: initialisation / iteration statistics
    ("mix_converged", SMALL, "-k 2 -n 3 -w n 2"),
    ("mix_s1_seeded", SMALL, "-k 3 -s 1 -n 3 -r 5 -w n 3"),
    ("admix_converged", dict(I=30, L=20, K=2, jmax=4, miss=200, P=2), "-a -k 2 -n 3 -w n 2"),
    # a K sweep: the "Average K (AIC / BIC)" branch
    ("admix_sweep", SMALL, "-a -1 2 -2 3 -n 2 -T 6 -E 1e-30 -w n 2"),
    ("mix_sweep_converged", SMALL, "-1 2 -2 4 -n 2 -w n 3"),
]


def main():
    for name, gen, cmd in CASES:
        tmp = tempfile.mkdtemp(prefix="mctimed_")
        subprocess.check_call([MC_GEN, "--I", str(gen["I"]), "--L", str(gen["L"]),
                               "--K", str(gen["K"]), "--jmax", str(gen["jmax"]),
                               "--miss", str(gen["miss"]), "--P", str(gen["P"]),
                               "--stru", os.path.join(tmp, "d.stru")],
                              stdout=subprocess.DEVNULL)
        r = subprocess.run([REF, "-f", "d.stru"] + cmd.split(), cwd=tmp,
                           capture_output=True, text=True)
        last = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        if r.returncode != 0 or not last.startswith(("Maximum iterations", "Average K (BIC)")):
            sys.exit("%s: reference failed or aborted (%d): %s | %s" % (
                name, r.returncode, last, r.stderr[-500:]))
        with open(os.path.join(HERE, "timed_%s.json" % name), "w") as fp:
            json.dump({"gen": gen, "cmd": cmd, "stdout": r.stdout}, fp, indent=1)
        print(name, "ok:", len(r.stdout.splitlines()), "lines;", last)


if __name__ == "__main__":
    main()
