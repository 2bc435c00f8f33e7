"""time the log-likelihood-only pass and a SQUAREM-style cycle's pieces at config 3"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multiclust_b200 import Context, SynthParams
I, L, K = 100000, 10000, 10
ctx = Context(0)
ctx.set_data_synth(I, L, SynthParams(seed=20261018, K=K, jmax=20, miss_bp=500, ploidy=2))
lb = 1e-8
ctx.alloc_model(K, admixture=1, q=1, eta_lb=lb, p_lb=lb)
J = ctx.get_J(); T = int(J.sum())
rng = np.random.default_rng(7)
eta = rng.random((I, K)) + 0.05; eta /= eta.sum(1, keepdims=True)
p = rng.random((K, T)) + 0.05
off = np.concatenate([[0], np.cumsum(J)])
p /= np.repeat(np.add.reduceat(p, off[:-1], axis=1), J, axis=1)
ctx.set_params(0, eta.ravel().copy(), p.ravel().copy())
for _ in range(3):
    ctx.em_step(0, 0); ctx.loglik(0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): ctx.loglik(0)
torch.cuda.synchronize(); t1 = time.perf_counter()
for _ in range(10): ctx.em_step(0, 0)
torch.cuda.synchronize(); t2 = time.perf_counter()
print("C3: loglik pass %.2f ms, EM step %.2f ms" % ((t1 - t0) * 100, (t2 - t1) * 100))
