"""The admixture initialiser with the cluster draws made on the device
(mc_init_admixture_rand) against (a) the same ABI fed with
the assignment a host loop over the generator draws and (b) the oracle's
initialiser, which calls the C library's rand() like the reference
(rnd_init.c:456-482).  The generator is restated here in a few lines of Python
(glibc TYPE_3: x[n] = x[n-31] + x[n-3] mod 2^32, rand() = x[n] >> 1)."""
import numpy as np
import pytest

from common import gen_data, glibc_stream, bootstrap_numpy, codes_to_counts, ROOT

pytestmark = pytest.mark.gpu


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
    want = bootstrap_numpy(draws, I, L, P, K, J, off, eta.ravel(), p.ravel(), admixture,
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


def _bootsamples():
    import glob
    import os
    return sorted(os.path.basename(f)[len("bootsample_"):-len(".npz")]
                  for f in glob.glob(os.path.join(ROOT, "tests", "golden", "bootsample_*.npz")))


@pytest.mark.parametrize("name", _bootsamples())
@pytest.mark.parametrize("block", [496, 31744])
def test_bootstrap_sample_matches_reference_function(name, block):
    """mc_bootstrap_data against a sample made by the reference's own parametric_bootstrap()
    (bootstrap.c:31-175) from the same parameters and the same rand() stream: equal allele
    counts for every individual and allele slot"""
    import os
    from multiclust_b200 import Context
    g = np.load(os.path.join(ROOT, "tests", "golden", "bootsample_%s.npz" % name))
    J, codes, K = g["J"], g["codes"], int(g["K"])
    I, L, P = codes.shape
    admixture, per_indiv = int(g["admixture"]), int(g["per_indiv"])
    n = I * (2 * L * P if admixture else 1 + L * P)
    x = glibc_stream(int(g["seed"]), n)
    nb = max(1, -(-n // block))
    hist = np.stack([x[b * block: b * block + 31] for b in range(nb)]).astype(np.uint32)
    ctx = Context(0)
    try:
        ctx.set_data(J, codes)
        ctx.alloc_model(K, admixture=admixture, eta_constrained=int(admixture and not per_indiv),
                        q=0)
        ctx.set_params(0, g["eta"], g["p"])
        ctx.save_mle(0)
        ctx.bootstrap_data(hist, block)
        got = codes_to_counts(ctx.get_codes(), J)
        assert np.array_equal(got, g["counts"].astype(np.int64))
    finally:
        ctx.close()
