"""where a small problem's EM step goes: streaming kernel vs everything else"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multiclust_b200 import Context, SynthParams
for I, L, K in ((200, 100, 3), (2000, 1000, 6), (20000, 2000, 6)):
    ctx = Context(0)
    ctx.set_data_synth(I, L, SynthParams(seed=1, K=4, jmax=6, miss_bp=200, ploidy=2))
    ctx.alloc_model(K, admixture=1, q=0, eta_lb=1e-8, p_lb=1e-8)
    J = ctx.get_J(); T = int(J.sum())
    rng = np.random.default_rng(7)
    eta = rng.random((I, K)) + 0.05; eta /= eta.sum(1, keepdims=True)
    p = rng.random((K, T)) + 0.05
    off = np.concatenate([[0], np.cumsum(J)])
    p /= np.repeat(np.add.reduceat(p, off[:-1], axis=1), J, axis=1)
    ctx.set_params(0, eta.ravel().copy(), p.ravel().copy())
    for _ in range(5): ctx.em_step(0, 0)
    ctx.profile_read(); ctx.profile_enable(True)
    n0 = ctx.launch_count()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): ctx.em_step(0, 0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ctx.profile_enable(False)
    nk, kms = ctx.profile_read()
    print("I=%d L=%d K=%d: %.1f us per EM step, streaming kernel %.1f us, %d launches per step, plan grid %d"
          % (I, L, K, (t1 - t0) / 200 * 1e6, kms / max(nk, 1) * 1e3, (ctx.launch_count() - n0) // 200, ctx.plan()["grid"]))
    ctx.close()
