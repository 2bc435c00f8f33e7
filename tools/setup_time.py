import time, sys, ctypes
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from multiclust_b200 import Context, SynthParams
ctx = Context(0)
sp = SynthParams(seed=20261018, K=10, jmax=20, miss_bp=500, ploidy=2)
I, L = 100000, 10000
ctx.set_data_synth(I, L, sp, i_first=0)
codes = np.empty((I, L, 2), dtype=np.uint8)
pinned = torch.from_numpy(codes).pin_memory().numpy()
ctx.lib.mc_get_codes(ctx.h, ctypes.c_void_p(pinned.ctypes.data))
J = ctx.get_J()
for rep in range(2):
    c2 = Context(0)
    c2.set_option(c2.OPT_TIMING, 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c2.set_data(J, pinned)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    c2.alloc_model(10, admixture=1, q=0, eta_lb=1e-8, p_lb=1e-8)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("set_data %.3f s  alloc_model(plan+build) %.3f s" % (t1 - t0, t2 - t1))
    c2.close()

# the e2e sequence of bench.py, with timestamps
from multiclust_b200.sharding import sharded_em_step
rng = np.random.default_rng(7)
K = 10
T = int(J.sum())
eta0 = rng.random((I, K)) + 0.05; eta0 /= eta0.sum(1, keepdims=True); eta0 = eta0.ravel().copy()
p0 = np.full(K * T, 0.1)
off = np.concatenate([[0], np.cumsum(J)])
pp = rng.random((K, T)) + 0.05
for l in range(len(J)):
    if J[l]:
        pp[:, off[l]:off[l+1]] /= pp[:, off[l]:off[l+1]].sum(1, keepdims=True)
p0 = pp.ravel().copy()
c2 = Context(0)
torch.cuda.synchronize(); t = [time.perf_counter()]
c2.set_data(J, pinned); torch.cuda.synchronize(); t.append(time.perf_counter())
c2.alloc_model(K, admixture=1, q=0, eta_lb=1e-8, p_lb=1e-8); torch.cuda.synchronize(); t.append(time.perf_counter())
c2.set_params(0, eta0, p0); torch.cuda.synchronize(); t.append(time.perf_counter())
for _ in range(20):
    sharded_em_step(c2, None, 1, 0, 0, None)
torch.cuda.synchronize(); t.append(time.perf_counter())
eo = np.empty_like(eta0); po = np.empty_like(p0); post = np.empty_like(eta0)
c2.lib.mc_get_params(c2.h, 0, ctypes.c_void_p(eo.ctypes.data), ctypes.c_void_p(po.ctypes.data))
c2.lib.mc_get_posterior(c2.h, ctypes.c_void_p(post.ctypes.data)); torch.cuda.synchronize(); t.append(time.perf_counter())
print("set_data %.3f alloc %.3f set_params %.3f 20 steps %.3f get %.3f" % tuple(b - a for a, b in zip(t, t[1:])))
