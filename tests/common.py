"""Helpers shared by the test modules (data generation, golden fixtures)."""
import glob
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MC_GEN = os.path.join(ROOT, "multiclust_b200", "host", "mc_gen")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files if k != "meta"}
    g["meta"] = json.loads(str(z["meta"]))
    return g


def ensure_mc_gen():
    if not os.path.exists(MC_GEN):
        from multiclust_b200 import build
        build.build_host()
    return MC_GEN


def gen_data(tmpdir, I, L, K=3, jmax=5, miss=300, P=2, seed=None, stru=False):
    """run mc_gen; returns the MCB1 dict (and the STRUCTURE path when asked)"""
    from oracle import orc
    ensure_mc_gen()
    mcb = os.path.join(str(tmpdir), "d_%d_%d_%d_%d.mcb" % (I, L, jmax, P))
    cmd = [MC_GEN, "--I", str(I), "--L", str(L), "--K", str(K), "--jmax", str(jmax),
           "--miss", str(miss), "--P", str(P), "--mcb", mcb]
    if seed is not None:
        cmd += ["--seed", str(seed)]
    spath = None
    if stru:
        spath = mcb[:-4] + ".stru"
        cmd += ["--stru", spath]
    subprocess.check_call(cmd)
    d = orc.read_mcb(mcb)
    d["mcb_path"] = mcb
    d["stru_path"] = spath
    return d


def random_params(rng, I, K, J, per_indiv=True):
    """a strictly positive, normalised parameter set on the flat layout"""
    T = int(np.sum(J))
    eta = rng.random((I if per_indiv else 1, K)) + 0.05
    eta /= eta.sum(axis=1, keepdims=True)
    p = rng.random((K, T)) + 0.05
    off = np.concatenate([[0], np.cumsum(J)])
    for l in range(len(J)):
        if J[l]:
            s = p[:, off[l]:off[l + 1]].sum(axis=1, keepdims=True)
            p[:, off[l]:off[l + 1]] /= s
    return eta.ravel().copy(), p.ravel().copy()
