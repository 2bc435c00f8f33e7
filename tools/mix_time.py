"""time the mixture-model EM step (BASELINE config 2 shape) and a log-likelihood pass"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from multiclust_b200 import Context, SynthParams
I, L, K = 10000, 5000, 5
ctx = Context(0)
ctx.set_data_synth(I, L, SynthParams(seed=20261018, K=K, jmax=2, miss_bp=0, ploidy=2))
J = ctx.get_J(); T = int(J.sum())
ctx.alloc_model(K, admixture=0, q=1, eta_lb=1e-8, p_lb=1e-8)
rng = np.random.default_rng(3)
eta = rng.random(K) + 0.1; eta /= eta.sum()
p = rng.random((K, T)) + 0.05
off = np.concatenate([[0], np.cumsum(J)])
for l in range(L):
    p[:, off[l]:off[l+1]] /= p[:, off[l]:off[l+1]].sum(1, keepdims=True)
ctx.set_params(0, eta.copy(), p.ravel().copy())
for _ in range(3): ll = ctx.em_step(0, 0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): ll = ctx.em_step(0, 0)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("mixture I=%d L=%d K=%d biallelic: %.3f ms per EM step (ll %.6f), plan %r" % (I, L, K, (t1 - t0) / 20 * 1e3, ll, ctx.plan()))
