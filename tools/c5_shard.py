"""One GPU's share of BASELINE config 5 (I = 1M over 8 GPUs -> 125k individuals,
L = 50k biallelic loci, tetraploid, K = 8): sizes past 2^32 bytes, a few EM
steps and a log-likelihood pass, monotone trajectory, timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multiclust_b200 import Context, SynthParams

I, L, K, P = int(os.environ.get("C5_I", 125000)), int(os.environ.get("C5_L", 50000)), 8, 4
ctx = Context(0)
t0 = time.perf_counter()
ctx.set_data_synth(I, L, SynthParams(seed=20261018, K=K, jmax=2, miss_bp=0, ploidy=P))
torch.cuda.synchronize(); t1 = time.perf_counter()
lb = min(1e-8, 0.5 / I / P)
ctx.alloc_model(K, admixture=1, q=2, eta_lb=lb, p_lb=lb)
torch.cuda.synchronize(); t2 = time.perf_counter()
J = ctx.get_J(); T = int(J.sum())
rng = np.random.default_rng(5)
eta = rng.random((I, K)) + 0.05; eta /= eta.sum(1, keepdims=True)
p = rng.random((K, T)) + 0.05
off = np.concatenate([[0], np.cumsum(J)])
starts = off[:-1]
sums = np.add.reduceat(p, starts, axis=1)
p /= np.repeat(sums, J, axis=1)
ctx.set_params(0, eta.ravel().copy(), p.ravel().copy())
lls = [ctx.em_step(0, 0) for _ in range(3)]
torch.cuda.synchronize(); t3 = time.perf_counter()
for _ in range(5):
    lls.append(ctx.em_step(0, 0))
torch.cuda.synchronize(); t4 = time.perf_counter()
ll_only = ctx.loglik(0)
free, total = torch.cuda.mem_get_info()
print("I=%d L=%d P=%d K=%d: generate %.1f s, plan+layout %.1f s, %.1f ms per EM step, device memory in use %.1f GB"
      % (I, L, P, K, t1 - t0, t2 - t1, (t4 - t3) / 5 * 1e3, (total - free) / 1e9))
print("plan", ctx.plan())
print("ll", lls, ll_only)
assert all(np.isfinite(lls)) and all(b >= a - 1e-9 * abs(a) for a, b in zip(lls, lls[1:]))
assert ll_only >= lls[-1]
print("ok")
