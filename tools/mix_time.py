"""Mixture EM step on multi-allelic data of BASELINE config 2's size (I=10k, L=5k, K=5,
diploid, <=20 alleles per locus, 5 % missing): the digit-sliced integer kernels on column
pairs (mc_digit.cuh) against the two-pass gather kernel's mixture modes (admix3 A3_MIX_E /
A3_MIX_M) and the round-1 one-pass tile kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiclust_b200 import Context, SynthParams

I, L, K = 10000, 5000, 5
for kernel, name in ((0, "auto (digit-sliced IMMA kernels on column pairs)"),
                     (2, "two-pass gather kernel, mixture modes"), (1, "one-pass tile kernel")):
    ctx = Context(0)
    ctx.set_option(ctx.OPT_KERNEL, kernel)
    ctx.set_data_synth(I, L, SynthParams(seed=20261018, K=K, jmax=20, miss_bp=500, ploidy=2))
    lb = min(1e-8, 0.5 / I / 2)
    ctx.alloc_model(K, admixture=0, q=0, eta_lb=lb, p_lb=lb)
    rng = np.random.default_rng(3)
    J = ctx.get_J(); T = int(J.sum())
    eta = np.full(K, 1.0 / K)
    p = rng.random((K, T)) + 0.1
    seg = np.repeat(np.arange(len(J)), J)
    for k in range(K):
        p[k] /= np.bincount(seg, weights=p[k], minlength=len(J))[seg]
    ctx.set_params(0, eta, p.ravel())
    lls = [ctx.em_step(0, 0) for _ in range(3)]
    ctx.profile_enable(True)
    ctx.sync(); t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        lls.append(ctx.em_step(0, 0))
    ctx.sync(); dt = (time.perf_counter() - t0) / n
    nk, ms = ctx.profile_read()
    bytes_ = I * L * 2
    print("%s: %.3f ms per EM step (streaming kernels %.3f ms in %d launches per step); HBM floor of "
          "one pass over the codes %.3f ms; ll %.6f -> %.6f monotone=%s; kernel family %d"
          % (name, dt * 1e3, ms / n, nk // n, bytes_ / 6456.2e9 * 1e3, lls[0], lls[-1],
             all(b >= a - 1e-9 * abs(a) for a, b in zip(lls, lls[1:])), ctx.plan()["two_pass"]),
          flush=True)
    ctx.close()
