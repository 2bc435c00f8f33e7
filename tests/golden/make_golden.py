#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ from the UNMODIFIED reference.

The reference ships no tests, fixtures or golden vectors,
so parity is pinned on outputs of the reference itself: oracle/Makefile
compiles /root/reference in place into oracle/_ref/, oracle/ref_harness.c
drives its own read_file / initialize_model / em_step / em_2_steps /
accelerated_em_step / log_likelihood and dumps full-precision state, and this
script packs those dumps into small .npz files.  Run it in the build container
(needs /root/reference); the GPU box only reads the committed files.

    python tests/golden/make_golden.py

Every case is: mc_gen (seeded synthetic STRUCTURE text) -> reference parser ->
reference initialiser (glibc rand()) -> reference EM driver.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import orc  # noqa: E402

MC_GEN = os.path.join(ROOT, "multiclust_b200", "host", "mc_gen")

# name, generator arguments, reference command line (without -f)
BASE = dict(I=60, L=40, K=3, jmax=5, miss=300, P=2)
CASES = [
    ("admix_em", BASE, "-a -k 3 -T 30 -E 1e-30 -n 2"),
    ("admix_s1", BASE, "-a -k 3 -s 1 -T 24 -E 1e-30 -n 2"),
    ("admix_s2", BASE, "-a -k 3 -s 2 -T 24 -E 1e-30 -n 1"),
    ("admix_s3", BASE, "-a -k 3 -s 3 -T 24 -E 1e-30 -n 1"),
    ("admix_s4", BASE, "-a -k 3 -s 4 -T 24 -E 1e-30 -n 1"),
    ("admix_s5", BASE, "-a -k 3 -s 5 -T 24 -E 1e-30 -n 1"),
    ("admix_s6", BASE, "-a -k 3 -s 6 -T 24 -E 1e-30 -n 1"),
    ("admix_c_em", BASE, "-a -c -k 3 -T 4 -E 1e-30 -n 2"),
    ("admix_c_s3", BASE, "-a -c -k 3 -s 3 -T 4 -E 1e-30 -n 1"),
    ("admix_noproj", BASE, "-a -k 3 -T 12 -E 1e-30 -n 1 --projection"),
    ("admix_seed", BASE, "-a -k 4 -r 7 -T 10 -E 1e-30 -n 1"),
    ("admix_conv", BASE, "-a -k 2 -n 1"),
    ("admix_k1", BASE, "-a -k 1 -n 1"),
    ("admix_tetra", dict(I=40, L=30, K=3, jmax=4, miss=500, P=4),
     "-a -k 3 -p 4 -T 15 -E 1e-30 -n 1"),
    ("admix_biallelic", dict(I=50, L=64, K=2, jmax=2, miss=0, P=2),
     "-a -k 2 -s 3 -T 20 -E 1e-30 -n 1"),
    ("admix_k10", dict(I=48, L=70, K=6, jmax=20, miss=500, P=2),
     "-a -k 10 -T 12 -E 1e-30 -n 1"),
    ("mix_em", BASE, "-k 3 -T 12 -E 1e-30 -n 2"),
    ("mix_s1", BASE, "-k 3 -s 1 -T 24 -E 1e-30 -n 1"),
    ("mix_s3", BASE, "-k 3 -s 3 -T 24 -E 1e-30 -n 1"),
    ("mix_s4", BASE, "-k 3 -s 4 -T 24 -E 1e-30 -n 1"),
    ("mix_s5", BASE, "-k 3 -s 5 -T 24 -E 1e-30 -n 1"),
    ("mix_s6", BASE, "-k 3 -s 6 -T 24 -E 1e-30 -n 1"),
    ("mix_conv", BASE, "-k 2 -n 1"),
    ("mix_biallelic_k5", dict(I=80, L=120, K=5, jmax=2, miss=0, P=2),
     "-k 5 -s 1 -T 7 -E 1e-30 -n 1"),
    # K sweep with several initialisations: one shared rand() stream
    ("admix_sweep", dict(I=30, L=20, K=3, jmax=4, miss=200, P=2),
     "-a -1 2 -2 4 -T 6 -E 1e-30 -n 2"),
    # mixture sweep from K=1 (one fit only, multiclust.c:630-631) on few
    # individuals, so the centre draws collide and are re-drawn (rnd_init.c:205-217)
    ("mix_sweep", dict(I=12, L=30, K=3, jmax=4, miss=200, P=2),
     "-1 1 -2 6 -n 3"),
]


def parse_cmd(cmd):
    """the options the oracle / host driver need, from the command line"""
    w = cmd.split()
    o = dict(admixture=0, eta_constrained=0, accel=0, do_projection=1,
             max_iter=0, abs_error=1e-4, rel_error=0.0, n_init=50, seed=-1,
             min_K=6, max_K=6, ploidy=2)
    i = 0
    while i < len(w):
        a = w[i]
        if a == "-a":
            o["admixture"] = 1
        elif a == "-c":
            o["eta_constrained"] = 1
        elif a == "--projection":
            o["do_projection"] = 0
        elif a == "-k":
            o["min_K"] = o["max_K"] = int(w[i + 1]); i += 1
        elif a == "-1":
            o["min_K"] = int(w[i + 1]); i += 1
        elif a == "-2":
            o["max_K"] = int(w[i + 1]); i += 1
        elif a == "-s":
            o["accel"] = int(w[i + 1]); i += 1
        elif a == "-T":
            o["max_iter"] = int(w[i + 1]); i += 1
        elif a == "-E":
            o["abs_error"] = float(w[i + 1]); i += 1
        elif a == "-n":
            o["n_init"] = int(w[i + 1]); i += 1
        elif a == "-r":
            o["seed"] = int(w[i + 1]); i += 1
        elif a == "-p":
            o["ploidy"] = int(w[i + 1]); i += 1
        else:
            raise ValueError(a)
        i += 1
    return o


def make_case(name, gen, cmd, outdir):
    tmp = tempfile.mkdtemp(prefix="mcgold_")
    try:
        stru = os.path.join(tmp, "d.stru")
        mcb = os.path.join(tmp, "d.mcb")
        subprocess.check_call([MC_GEN, "--I", str(gen["I"]), "--L", str(gen["L"]),
                               "--K", str(gen["K"]), "--jmax", str(gen["jmax"]),
                               "--miss", str(gen["miss"]), "--P", str(gen["P"]),
                               "--stru", stru, "--mcb", mcb])
        pre = os.path.join(tmp, "r")
        r = orc.run_ref(["-f", stru] + cmd.split(), dump=pre, steps=True)
        if r.returncode != 0:
            raise RuntimeError("%s: reference harness failed: %s" % (name, r.stderr))
        parsed = orc.read_mcb(pre + ".parse.mcb")
        gen_mcb = orc.read_mcb(mcb)
        # mc_gen's own recoding must equal the reference parser's
        for key in ("J", "nreal", "labels", "codes", "locale"):
            assert np.array_equal(parsed[key], gen_mcb[key]), (name, key)
        tr = orc.read_trace(pre + ".trace.txt")
        opts = parse_cmd(cmd)
        out = dict(J=parsed["J"], codes=parsed["codes"], locale=parsed["locale"],
                   nreal=parsed["nreal"], labels=parsed["labels"])
        with open(pre + ".trace.txt") as fp:
            first = fp.readline().split()
        bounds = float(first[1])
        fits = []
        for (K, init), fit in sorted(tr["fit"].items()):
            tag = "K%d.init%d" % (K, init)
            st = orc.read_state("%s.%s.start.bin" % (pre, tag))
            fi = orc.read_state("%s.%s.final.bin" % (pre, tag))
            key = "K%d_i%d_" % (K, init)
            out[key + "start_eta"] = st["eta"]
            out[key + "start_p"] = st["p"]
            out[key + "final_eta"] = fi["eta"]
            out[key + "final_p"] = fi["p"]
            out[key + "final_post"] = fi["posterior"]
            out[key + "ll"] = np.array(tr["ll"].get((K, init), []))
            steps = tr["steps"].get((K, init), [])
            out[key + "step_logL"] = np.array([s["logL"] for s in steps])
            out[key + "step_pindex"] = np.array([s["pindex"] for s in steps], dtype=np.int32)
            out[key + "step_n_iter"] = np.array([s["n_iter"] for s in steps], dtype=np.int32)
            # parameters after the first and the last top-level step
            for which in (1, len(steps)):
                path = "%s.%s.step%d.bin" % (pre, tag, which)
                if which >= 1 and os.path.exists(path):
                    s = orc.read_state(path)
                    out[key + "step%d_eta" % which] = s["eta"]
                    out[key + "step%d_p" % which] = s["p"]
            fits.append(dict(K=K, init=init, **fit))
        meta = dict(name=name, gen=gen, cmd=cmd, options=opts, bound=bounds,
                    fits=fits)
        out["meta"] = np.array(json.dumps(meta))
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        return meta
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    if not os.path.isdir("/root/reference"):
        sys.exit("make_golden.py needs /root/reference (run it in the build container)")
    orc.build()
    from multiclust_b200 import build as mcbuild
    mcbuild.build_host()
    only = set(sys.argv[1:])  # optional: names of the cases to (re)generate
    for name, gen, cmd in CASES:
        if only and name not in only:
            continue
        meta = make_case(name, gen, cmd, HERE)
        if not meta["fits"]:
            sys.exit("%s: the reference aborted (exit(0) on a log likelihood "
                     "decrease, em_alg.c:115); pick a shorter run" % name)
        print("%-18s %2d fits  ll[0]=%r" % (name, len(meta["fits"]),
                                             meta["fits"][0]["logL"]))


if __name__ == "__main__":
    main()
