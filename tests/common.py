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
    # (bootsample_*.npz are single bootstrap samples: tests/test_bootstrap_golden.py)
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith("bootsample_"))


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


# ---- the reference's rand() stream and bootstrap sampler, restated ----

def glibc_stream(seed, n):
    """the 31 words in front of the first draw, then the n words of n draws"""
    r = [0] * 31
    r[0] = seed if seed else 1
    for i in range(1, 31):
        hi, lo = divmod(r[i - 1], 127773)
        w = 16807 * lo - 2836 * hi
        r[i] = w + 2147483647 if w < 0 else w
    # srand() leaves front = 3, rear = 0: the first update is r[3] += r[0], so in
    # linear terms the oldest word is r[3] -- the history is r rotated by 3
    x = r[3:] + r[:3]
    for _ in range(310 + n):
        x.append((x[-31] + x[-3]) & 0xffffffff)
    return np.array(x[310:], dtype=np.uint64)   # history (31) + n draws



def boot_pick(w, r):
    """bootstrap.c:96-105: first index whose running sum reaches r, else the last"""
    j, acc = 0, 0.0
    while j < len(w) and r > acc:
        acc += w[j]
        j += 1
    return j - 1 if j else 0


def bootstrap_numpy(draws, I, L, P, K, J, off, eta, p, admixture, per_indiv):
    """the reference's loops in the default parse mode: every copy is drawn"""
    out = np.full((I, L, P), 255, dtype=np.uint8)
    T = int(off[-1])
    p = p.reshape(K, T)
    d = 0
    for i in range(I):
        k = 0
        if not admixture:
            k = boot_pick(eta, draws[d] / 2147483647.0); d += 1
        for l in range(L):
            for a in range(P):
                if admixture:
                    row = eta.reshape(I, K)[i] if per_indiv else eta
                    k = boot_pick(row, draws[d] / 2147483647.0); d += 1
                r = draws[d] / 2147483647.0; d += 1
                if J[l] > 0:
                    out[i, l, a] = boot_pick(p[k, off[l]:off[l] + J[l]], r)
    assert d == draws.size
    return out



def codes_to_counts(codes, J):
    """allele codes [I][L][P] (255 = missing) -> counts [I][sum J] per allele slot"""
    I, L, P = codes.shape
    off = np.concatenate([[0], np.cumsum(J)]).astype(np.int64)
    out = np.zeros((I, int(off[-1])), dtype=np.int64)
    for l in range(L):
        for a in range(P):
            c = codes[:, l, a].astype(np.int64)
            ok = c != 255
            np.add.at(out, (np.nonzero(ok)[0], off[l] + c[ok]), 1)
    return out
