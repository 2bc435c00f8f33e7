#!/usr/bin/env python
"""Golden output of the reference's parametric bootstrap (-b, bootstrap.c, multiclust.c:675-708).

The bootstrap has no state worth dumping beyond what it prints: per sample the summary line of
the H0 and the Ha fit (maximum log likelihood, AIC, BIC at %f) and the two test statistics, then
the p-value.  This script runs the STOCK reference binary (oracle/_ref/multiclust, compiled in
place from /root/reference by oracle/Makefile) on mc_gen data and stores its stdout; the GPU test
(tests/test_cli_gpu.py::test_cli_bootstrap_matches_reference) runs the product binary on the same
generated file with the same arguments and compares line by line, numbers to 2e-6.

It also stores single bootstrap SAMPLES made by the reference's own parametric_bootstrap()
from dumped parameters and a re-seeded rand() stream (oracle/ref_harness.c --bootstrap-sample):
tests/golden/bootsample_*.npz pin the sampler itself -- the Python restatement on the CPU
(tests/test_oracle_golden.py) and mc_bootstrap_data on the GPU (tests/test_init_rand_gpu.py).

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


# one bootstrap sample made by the reference's own parametric_bootstrap() from known
# parameters and a known rand() stream (oracle/ref_harness.c --bootstrap-sample): name,
# generator arguments, command line of the fit whose final parameters are the H0 estimates,
# seed the generator is reset to in front of the sample
SAMPLES = [
    ("admix", SMALL, "-a -k 3 -n 1 -b 1 -T 5 -E 1e-30", 5),
    ("admix_pooled", SMALL, "-a -c -k 3 -n 1 -b 1 -T 5 -E 1e-30", 6),
    ("mix", SMALL, "-k 3 -n 1 -b 1 -T 5 -E 1e-30", 7),
    ("admix_tetra", dict(I=30, L=25, K=2, jmax=3, miss=200, P=4),
     "-a -k 2 -p 4 -n 1 -b 1 -T 4 -E 1e-30", 8),
]


def make_samples():
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import orc
    for name, gen, cmd, seed in SAMPLES:
        tmp = tempfile.mkdtemp(prefix="mcboot_")
        stru = os.path.join(tmp, "d.stru")
        subprocess.check_call([MC_GEN, "--I", str(gen["I"]), "--L", str(gen["L"]),
                               "--K", str(gen["K"]), "--jmax", str(gen["jmax"]),
                               "--miss", str(gen["miss"]), "--P", str(gen["P"]),
                               "--stru", stru], stdout=subprocess.DEVNULL)
        pre = os.path.join(tmp, "r")
        r = orc.run_ref(["-f", stru] + cmd.split(), dump=pre, bootstrap_seed=seed)
        if r.returncode != 0:
            sys.exit("%s: reference harness failed: %s" % (name, r.stderr[-500:]))
        parsed = orc.read_mcb(pre + ".parse.mcb")
        K = int(cmd.split()[cmd.split().index("-k") + 1])
        st = orc.read_state("%s.K%d.init0.final.bin" % (pre, K))
        T = int(parsed["J"].sum())
        counts = np.fromfile(pre + ".bootstrap.bin", dtype="<i4").reshape(gen["I"], T)
        np.savez_compressed(os.path.join(HERE, "bootsample_%s.npz" % name),
                            J=parsed["J"], codes=parsed["codes"], K=K, seed=seed,
                            admixture=st["admixture"], per_indiv=st["per_indiv"],
                            eta=st["eta"], p=st["p"], counts=counts.astype(np.int8))
        print("sample", name, "ok: copies per individual and locus",
              sorted(set(np.add.reduceat(counts, np.concatenate([[0], np.cumsum(
                  parsed["J"])[:-1]])[parsed["J"] > 0], axis=1).ravel().tolist())))


def main():
    make_samples()
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
