"""Time the dense (biallelic) configurations: BASELINE config 2 (mixture, I=10k, L=5k, K=5,
diploid) and one GPU's share of config 5 (admixture, I=125k, L=50k, K=8, tetraploid).
  python tools/dense_time.py [c2] [c5] [c5mix] [--kernel N] [--steps N]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiclust_b200 import Context, SynthParams


def params(ctx, K, per_indiv, seed=3):
    rng = np.random.default_rng(seed)
    J = ctx.get_J(); T = int(J.sum())
    eta = rng.random((ctx.I if per_indiv else 1, K)) + 0.1
    eta /= eta.sum(axis=1, keepdims=True)
    p = rng.random((K, T)) + 0.1
    seg = np.repeat(np.arange(len(J)), J)
    for k in range(K):
        p[k] /= np.bincount(seg, weights=p[k], minlength=len(J))[seg]
    return eta.ravel(), p.ravel()


def run(name, I, L, K, P, admixture, kernel, steps=10, miss=0):
    ctx = Context(0)
    ctx.set_option(ctx.OPT_KERNEL, kernel)
    sp = SynthParams(seed=20261018, K=K, jmax=2, miss_bp=miss, ploidy=P)
    t0 = time.perf_counter()
    ctx.set_data_synth(I, L, sp)
    t1 = time.perf_counter()
    lb = min(1e-8, 0.5 / I / P)
    ctx.alloc_model(K, admixture=admixture, q=0, eta_lb=lb, p_lb=lb)
    t2 = time.perf_counter()
    eta, p = params(ctx, K, bool(admixture))
    ctx.set_params(0, eta, p)
    lls = [ctx.em_step(0, 0) for _ in range(3)]
    ctx.profile_enable(True)
    ctx.sync()
    t3 = time.perf_counter()
    for _ in range(steps):
        lls.append(ctx.em_step(0, 0))
    ctx.sync()
    t4 = time.perf_counter()
    n, ms = ctx.profile_read()
    ctx.profile_enable(False)
    for _ in range(3):          # the replayed launch graphs exist from the third call on
        ctx.loglik(0)
        ctx.em_step(0, 0)
    ctx.sync()
    t5 = time.perf_counter()
    for _ in range(steps):
        ctx.loglik(0)
    ctx.sync()
    t6 = time.perf_counter()
    for _ in range(steps):
        ctx.em_step(0, 0)
    ctx.sync()
    t7 = time.perf_counter()
    ok = all(b >= a - 1e-9 * abs(a) for a, b in zip(lls, lls[1:]))
    print("%s I=%d L=%d K=%d P=%d kernel=%d: %.3f ms per EM step (streaming kernels %.3f ms in %d "
          "launches per step), replayed as a graph %.3f ms, loglik pass %.3f ms; synth %.2f s, plan %.2f s; ll %.6f -> %.6f "
          "monotone=%s; plan %s" % (name, I, L, K, P, kernel, (t4 - t3) / steps * 1e3,
                                    ms / steps, n // steps, (t7 - t6) / steps * 1e3,
                                    (t6 - t5) / steps * 1e3, t1 - t0,
                                    t2 - t1, lls[0], lls[-1], ok, ctx.plan()), flush=True)
    ctx.close()


if __name__ == "__main__":
    args = sys.argv[1:]
    kernel = 0
    if "--kernel" in args:
        kernel = int(args[args.index("--kernel") + 1])
    steps = int(args[args.index("--steps") + 1]) if "--steps" in args else 0
    named = [x for x in args if x in ("c2", "c5", "c5mix")]
    if "c2" in args or not named:
        run("C2 mixture", 10000, 5000, 5, 2, 0, kernel, steps=steps or 20)
    if "c5" in args or not named:
        run("C5 share admixture", 125000, 50000, 8, 4, 1, kernel, steps=steps or 5)
    if "c5mix" in args:     # the mixture kernels at a size that fills the machine
        run("C5-sized mixture", 125000, 50000, 8, 4, 0, kernel, steps=steps or 5)
