#!/usr/bin/env python
"""Golden output of the reference's parametric bootstrap (-b, bootstrap.c, multiclust.c:675-708).

The bootstrap has no state worth dumping beyond what it prints: per sample the summary line of
the H0 and the Ha fit (maximum log likelihood, AIC, BIC at %f) and the two test statistics, then
the p-value.  This script runs the STOCK reference binary (oracle/_ref/multiclust, compiled in
place from /root/reference by oracle/Makefile) on mc_gen data and stores its stdout; the GPU test
(tests/test_cli_gpu.py::test_cli_bootstrap_matches_reference) runs the product binary on the same
generated file with the same arguments and compares line by line, numbers to 2e-6.

    python tests/golden/make_bootstrap_golden.py
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
    # cases the reference completes: its two numeric aborts (a log likelihood that decreases
    # by one rounding error exits with status 0, em_alg.c:116) hit many bootstrap fits, and
    # with -c every K reaches the same likelihood, so the test refuses to run at all
    ("admix", SMALL, "-a -k 3 -n 2 -b 3 -T 8 -E 1e-30"),
    ("mix", SMALL, "-k 3 -n 2 -b 3 -T 6 -E 1e-30"),
    ("admix_tetra", dict(I=30, L=25, K=2, jmax=3, miss=200, P=4),
     "-a -k 3 -p 4 -n 2 -b 2 -T 7 -E 1e-30"),
    ("mix_biallelic_s1", dict(I=50, L=60, K=3, jmax=2, miss=0, P=2),
     "-k 3 -s 1 -n 2 -b 2 -T 10 -E 1e-30"),
    ("admix_biallelic_s5", dict(I=50, L=64, K=2, jmax=2, miss=0, P=2),
     "-a -k 2 -s 5 -n 3 -b 2 -T 12 -E 1e-30"),
    ("admix_converged", dict(I=30, L=20, K=2, jmax=4, miss=200, P=2),
     "-a -k 2 -n 3 -b 2"),
    ("mix_seeded_k1", SMALL, "-k 2 -n 3 -b 2 -r 7 -T 9 -E 1e-30"),
    ("mix_converged", SMALL, "-k 3 -n 3 -b 2 -r 11"),
]


def main():
    for name, gen, cmd in CASES:
        tmp = tempfile.mkdtemp(prefix="mcboot_")
        subprocess.check_call([MC_GEN, "--I", str(gen["I"]), "--L", str(gen["L"]),
                               "--K", str(gen["K"]), "--jmax", str(gen["jmax"]),
                               "--miss", str(gen["miss"]), "--P", str(gen["P"]),
                               "--stru", os.path.join(tmp, "d.stru")],
                              stdout=subprocess.DEVNULL)
        os.mkdir(os.path.join(tmp, "out"))
        r = subprocess.run([REF, "-f", "d.stru"] + cmd.split() + ["-d", "out/"], cwd=tmp,
                           capture_output=True, text=True)
        if r.returncode != 0 or not r.stdout.strip().splitlines()[-1].startswith("p-value"):
            sys.exit("%s: reference failed or aborted (%d): %s" % (name, r.returncode,
                                                                   r.stderr[-500:]))
        with open(os.path.join(HERE, "bootstrap_%s.json" % name), "w") as fp:
            json.dump({"gen": gen, "cmd": cmd, "stdout": r.stdout}, fp, indent=1)
        print(name, "ok:", r.stdout.strip().splitlines()[-1])


if __name__ == "__main__":
    main()
