"""The admixture initialiser with the cluster draws made on the device
(mc_init_admixture_rand) against (a) the same ABI fed with
the assignment a host loop over the generator draws and (b) the oracle's
initialiser, which calls the C library's rand() like the reference
(rnd_init.c:456-482).  The generator is restated here in a few lines of Python
(glibc TYPE_3: x[n] = x[n-31] + x[n-3] mod 2^32, rand() = x[n] >> 1)."""
import numpy as np
import pytest

from common import gen_data

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("shape,K,block", [
    ((300, 130, 2), 3, 31744),     # two full blocks and a tail
    ((300, 130, 2), 10, 496),      # many single-round blocks
    ((64, 31, 4), 7, 992),         # tetraploid
    ((5, 3, 2), 2, 31744),         # fewer draws than one round
])
def test_device_draws_match_host_draws(tmp_path, shape, K, block):
    from multiclust_b200 import Context
    from oracle import orc
    I, L, P = shape
    d = gen_data(tmp_path, I, L, K=3, jmax=6, miss=400, P=P)
    n = I * L * P
    seed = 7
    x = glibc_stream(seed, n)
    z = ((x[31:] >> np.uint64(1)) % np.uint64(K)).astype(np.uint8)
    nb = max(1, -(-n // block))
    hist = np.stack([x[b * block: b * block + 31] for b in range(nb)]).astype(np.uint32)

    fit = orc.Fit(d["J"], d["codes"], admixture=1)
    fit.alloc(K)
    orc.seed(seed)
    fit.initialize()
    eta_o, p_o = fit.get_params(0)

    ctx = Context(0)
    ctx.set_data(d["J"], d["codes"])
    lb = fit.lower_bound
    ctx.alloc_model(K, admixture=1, q=0, eta_lb=lb, p_lb=lb)
    ctx.init_admixture(0, z)
    eta_h, p_h = ctx.get_params(0)
    ctx.init_admixture_rand(1, hist, block)
    eta_d, p_d = ctx.get_params(1)
    ctx.close()
    assert np.array_equal(eta_d, eta_h) and np.array_equal(p_d, p_h)
    assert np.max(np.abs(eta_d - eta_o)) < 1e-12 and np.max(np.abs(p_d - p_o)) < 1e-12


# ---- parametric bootstrap samples (mc_bootstrap_data, bootstrap.c:77-175) ----

def _pick(w, r):
    """bootstrap.c:96-105: first index whose running sum reaches r, else the last"""
    j, acc = 0, 0.0
    while j < len(w) and r > acc:
        acc += w[j]
        j += 1
    return j - 1 if j else 0


def _bootstrap_numpy(draws, I, L, P, K, J, off, eta, p, admixture, per_indiv):
    """the reference's loops in the default parse mode: every copy is drawn"""
    out = np.full((I, L, P), 255, dtype=np.uint8)
    T = int(off[-1])
    p = p.reshape(K, T)
    d = 0
    for i in range(I):
        k = 0
        if not admixture:
            k = _pick(eta, draws[d] / 2147483647.0); d += 1
        for l in range(L):
            for a in range(P):
                if admixture:
                    row = eta.reshape(I, K)[i] if per_indiv else eta
                    k = _pick(row, draws[d] / 2147483647.0); d += 1
                r = draws[d] / 2147483647.0; d += 1
                if J[l] > 0:
                    out[i, l, a] = _pick(p[k, off[l]:off[l] + J[l]], r)
    assert d == draws.size
    return out


@pytest.mark.parametrize("model", ["admixture", "pooled", "mixture"])
@pytest.mark.parametrize("shape,K,block", [((23, 17, 2), 3, 496), ((9, 40, 4), 5, 31744),
                                           ((40, 7, 1), 2, 992)])
def test_bootstrap_sample_matches_reference_loops(tmp_path, model, shape, K, block):
    """the device's bootstrap sample, bit for bit, against the reference's loops restated in
    Python on the same rand() stream; the admixture initialiser keeps reading the observed
    data (rnd_init.c:460-481 reads dat->IL, which bootstrap.c never rewrites); the observed
    data come back with mc_restore_data"""
    from multiclust_b200 import Context
    I, L, P = shape
    d = gen_data(tmp_path, I, L, K=3, jmax=5, miss=400, P=P)
    J = d["J"]
    off = np.concatenate([[0], np.cumsum(J)]).astype(np.int64)
    admixture = model != "mixture"
    per_indiv = model == "admixture"
    rng = np.random.default_rng(11)
    eta = rng.random((I if per_indiv else 1, K)) + 0.05
    eta /= eta.sum(axis=1, keepdims=True)
    p = rng.random((K, int(off[-1]))) + 0.05
    for l in range(L):
        if J[l]:
            p[:, off[l]:off[l + 1]] /= p[:, off[l]:off[l + 1]].sum(axis=1, keepdims=True)
    n = I * (2 * L * P if admixture else 1 + L * P)
    x = glibc_stream(5, n)
    draws = (x[31:] >> np.uint64(1)).astype(np.float64)
    nb = max(1, -(-n // block))
    hist = np.stack([x[b * block: b * block + 31] for b in range(nb)]).astype(np.uint32)
    want = _bootstrap_numpy(draws, I, L, P, K, J, off, eta.ravel(), p.ravel(), admixture,
                            per_indiv)

    ctx = Context(0)
    try:
        ctx.set_data(J, d["codes"])
        ctx.alloc_model(K, admixture=int(admixture), eta_constrained=int(model == "pooled"), q=0)
        ctx.set_params(0, eta.ravel(), p.ravel())
        z = rng.integers(0, K, size=I * L * P).astype(np.uint8)
        if admixture:
            ctx.init_admixture(1, z)
            before = ctx.get_params(1)
        ctx.save_mle(0)
        ctx.bootstrap_data(hist, block)
        got = ctx.get_codes()
        assert np.array_equal(got, want)
        for rep in range(2):                    # a second sample overwrites the first
            ctx.alloc_model(K, admixture=int(admixture), eta_constrained=int(model == "pooled"),
                            q=0)
            if admixture:
                ctx.init_admixture(1, z)
                after = ctx.get_params(1)
                assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
            ctx.bootstrap_data(hist, block)
            assert np.array_equal(ctx.get_codes(), want)
        ctx.restore_data()
        assert np.array_equal(ctx.get_codes(), d["codes"])
    finally:
        ctx.close()
