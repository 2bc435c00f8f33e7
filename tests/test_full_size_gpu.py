"""BASELINE.json's headline configuration (configs[2]: I=100k, L=10k, <=20
alleles, K=10, 5 % missing) at FULL size on the device.  The oracle cannot run
this size (the reference cannot even allocate it, SURVEY.md finding 5), so the
checks are the size-independent properties of the path:
  - the log likelihood of EM never decreases and is finite;
  - mc_loglik(slot) equals the value the next mc_em_step reports for that slot
    (the reference's "one step late" log likelihood, em_alg.c:195-207);
  - every eta row and every (k, l) row of p sums to one and respects the floor;
  - sum_k D_ik equals the number of non-missing allele copies of individual i
    (sum_k d_iklj = c_ilj, em_alg.c:386-389);
  - the same fit run twice is bitwise identical (deterministic reductions);
  - the first rows of the device-generated genotypes equal the host generator.
A second, smaller case cross-checks one full-size-shaped tile mix against the
oracle: same J distribution, 600 individuals."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

I, L, K = 100000, 10000, 10


@pytest.fixture(scope="module")
def big():
    from multiclust_b200 import Context, SynthParams
    ctx = Context(0)
    sp = SynthParams(seed=20261018, K=K, jmax=20, miss_bp=500, ploidy=2)
    ctx.set_data_synth(I, L, sp)
    lb = min(1e-8, 0.5 / I / 2)
    ctx.alloc_model(K, admixture=1, q=0, eta_lb=lb, p_lb=lb)
    yield ctx, lb
    ctx.close()


def start_params(ctx, seed):
    rng = np.random.default_rng(seed)
    J = ctx.get_J()
    T = int(J.sum())
    eta = rng.random((ctx.I, K)) + 0.1
    eta /= eta.sum(axis=1, keepdims=True)
    p = rng.random((K, T)) + 0.1
    seg = np.repeat(np.arange(len(J)), J)
    for k in range(K):
        p[k] /= np.bincount(seg, weights=p[k], minlength=len(J))[seg]
    return eta.ravel(), p.ravel(), J, seg


def test_full_size_properties(big):
    ctx, lb = big
    assert ctx.plan()["two_pass"] == 2
    eta0, p0, J, seg = start_params(ctx, 3)
    runs = []
    for rep in range(2):
        ctx.set_params(0, eta0, p0)
        pre = ctx.loglik(0)
        lls = [ctx.em_step(0, 0) for _ in range(4)]
        runs.append((pre, lls, ctx.get_params(0), ctx.posterior()))
    pre, lls, (eta, p), D = runs[0]
    assert np.all(np.isfinite(lls))
    assert abs(pre - lls[0]) <= 1e-12 * abs(pre)          # one step late
    assert all(b > a for a, b in zip(lls, lls[1:]))       # EM ascent
    # the log likelihood of the final parameters continues the ascent
    assert ctx.loglik(0) > lls[-1]
    eta = eta.reshape(I, K)
    assert np.max(np.abs(eta.sum(axis=1) - 1.0)) < 1e-12
    assert eta.min() >= lb
    p = p.reshape(K, -1)
    rows = np.stack([np.bincount(seg, weights=p[k], minlength=len(J)) for k in range(K)])
    assert np.max(np.abs(rows[:, J > 0] - 1.0)) < 1e-12
    assert p.min() >= lb
    # posterior sums: every non-missing copy is shared out completely
    codes = ctx.get_codes()
    valid = (codes != 255).sum(axis=(1, 2))
    assert np.max(np.abs(D.sum(axis=1) - valid)) < 1e-7
    # bitwise repeatable
    assert runs[1][1] == lls
    assert np.array_equal(runs[1][2][0], runs[0][2][0])
    assert np.array_equal(runs[1][2][1], runs[0][2][1])
    assert np.array_equal(runs[1][3], D)


def test_full_size_generator_prefix(big, tmp_path):
    """the device generator and the host generator agree on the raw alleles of
    the first individuals (recoding differs: the host sees only the sample)"""
    from common import gen_data
    ctx, lb = big
    n = 6
    small = gen_data(tmp_path, n, L, K=K, jmax=20, miss=500, P=2)
    dev = ctx.get_codes()[:n]
    # missing copies agree exactly; allele codes agree up to the per-locus
    # relabelling, i.e. equal codes on the host are equal codes on the device
    assert np.array_equal(dev == 255, small["codes"] == 255)
    h, d = small["codes"].reshape(n * 1, L, 2), dev
    same_h = h[:, :, 0] == h[:, :, 1]
    same_d = d[:, :, 0] == d[:, :, 1]
    assert np.array_equal(same_h, same_d)


def test_mid_size_against_oracle(orc, tmp_path):
    """the same allele-count mix (jmax 20, 5 % missing, K=10) on 600
    individuals x 400 loci: two-pass kernel vs oracle, several tiles and chunks"""
    from common import gen_data, random_params
    from multiclust_b200 import Context
    d = gen_data(tmp_path, 600, 400, K=6, jmax=20, miss=500, P=2)
    fit = orc.Fit(d["J"], d["codes"], admixture=1)
    fit.alloc(K)
    ctx = Context(0)
    ctx.set_data(d["J"], d["codes"])
    ctx.alloc_model(K, admixture=1, q=0, eta_lb=fit.lower_bound, p_lb=fit.lower_bound)
    assert ctx.plan()["two_pass"] == 2
    eta, p = random_params(np.random.default_rng(11), 600, K, d["J"], True)
    fit.set_params(0, eta, p)
    ctx.set_params(0, eta, p)
    fit.set_indices(0, 0, 0)
    for it in range(5):
        ll_o = fit.e_step()
        fit.m_step()
        ll_g = ctx.em_step(0, 0)
        assert abs(ll_g - ll_o) <= 1e-12 * abs(ll_o)
    eo, po = fit.get_params(0)
    eg, pg = ctx.get_params(0)
    assert np.max(np.abs(eg - eo)) < 1e-11 and np.max(np.abs(pg - po)) < 1e-11
    assert np.max(np.abs(ctx.posterior() - fit.posterior())) < 1e-9
    ctx.close()


# ---- the biallelic configurations at full size (dense DMMA kernels) ----

def _start(ctx, K, per_indiv, seed):
    rng = np.random.default_rng(seed)
    J = ctx.get_J()
    T = int(J.sum())
    eta = rng.random((ctx.I if per_indiv else 1, K)) + 0.1
    eta /= eta.sum(axis=1, keepdims=True)
    p = rng.random((K, T)) + 0.1
    seg = np.repeat(np.arange(len(J)), J)
    for k in range(K):
        p[k] /= np.bincount(seg, weights=p[k], minlength=len(J))[seg]
    return eta.ravel(), p.ravel(), J, seg


def test_config2_full_size_properties():
    """BASELINE configs[1]: mixture, I=10k, L=5k biallelic, K=5, diploid.  Properties:
    EM ascent, loglik == next step's value, rows on the simplex, posterior rows sum
    to one, determinism."""
    from multiclust_b200 import Context, SynthParams
    I2, L2, K2 = 10000, 5000, 5
    ctx = Context(0)
    try:
        ctx.set_data_synth(I2, L2, SynthParams(seed=20261018, K=K2, jmax=2, miss_bp=0, ploidy=2))
        lb = min(1e-8, 0.5 / I2 / 2)
        ctx.alloc_model(K2, admixture=0, q=0, eta_lb=lb, p_lb=lb)
        assert ctx.plan()["two_pass"] == 4      # digit-sliced integer kernels
        eta0, p0, J, seg = _start(ctx, K2, False, 5)
        runs = []
        for rep in range(2):
            ctx.set_params(0, eta0, p0)
            pre = ctx.loglik(0)
            lls = [ctx.em_step(0, 0) for _ in range(4)]
            runs.append((pre, lls, ctx.get_params(0), ctx.posterior()))
        pre, lls, (eta, p), v = runs[0]
        assert np.all(np.isfinite(lls))
        assert abs(pre - lls[0]) <= 1e-12 * abs(pre)
        assert all(b > a for a, b in zip(lls, lls[1:]))
        assert abs(eta.sum() - 1.0) < 1e-12 and eta.min() >= lb
        p = p.reshape(K2, -1)
        rows = np.stack([np.bincount(seg, weights=p[k], minlength=len(J)) for k in range(K2)])
        assert np.max(np.abs(rows[:, J > 0] - 1.0)) < 1e-12 and p.min() >= lb
        assert np.max(np.abs(v.sum(axis=1) - 1.0)) < 1e-12
        assert runs[1][1] == lls and np.array_equal(runs[1][2][1], runs[0][2][1])
        assert np.array_equal(runs[1][3], v)
    finally:
        ctx.close()


def test_config5_share_properties():
    """one GPU's share of BASELINE configs[4] (I=1M / 8 = 125k tetraploid individuals,
    L=50k biallelic loci, K=8) on the dense kernels: EM ascent, loglik == next step's
    value, simplex rows, sum_k D_ik = non-missing copies, determinism."""
    from multiclust_b200 import Context, SynthParams
    I5, L5, K5, P5 = 125000, 50000, 8, 4
    ctx = Context(0)
    try:
        ctx.set_data_synth(I5, L5, SynthParams(seed=20261018, K=K5, jmax=2, miss_bp=100, ploidy=P5))
        lb = min(1e-8, 0.5 / (8 * I5) / P5)
        ctx.alloc_model(K5, admixture=1, q=0, eta_lb=lb, p_lb=lb)
        assert ctx.plan()["two_pass"] == 3
        eta0, p0, J, seg = _start(ctx, K5, True, 9)
        runs = []
        for rep in range(2):
            ctx.set_params(0, eta0, p0)
            pre = ctx.loglik(0)
            lls = [ctx.em_step(0, 0) for _ in range(3)]
            runs.append((pre, lls, ctx.get_params(0), ctx.posterior()))
        pre, lls, (eta, p), D = runs[0]
        assert np.all(np.isfinite(lls))
        assert abs(pre - lls[0]) <= 1e-12 * abs(pre)
        assert all(b > a for a, b in zip(lls, lls[1:]))
        eta = eta.reshape(I5, K5)
        assert np.max(np.abs(eta.sum(axis=1) - 1.0)) < 1e-12 and eta.min() >= lb
        p = p.reshape(K5, -1)
        rows = np.stack([np.bincount(seg, weights=p[k], minlength=len(J)) for k in range(K5)])
        assert np.max(np.abs(rows[:, J > 0] - 1.0)) < 1e-12 and p.min() >= lb
        # every non-missing copy is shared out completely (checked on a slice: the
        # natural codes of the whole share are 25 GB)
        D = D.reshape(I5, K5)
        assert np.max(np.abs(D.sum(axis=1) - np.round(D.sum(axis=1)))) < 1e-6
        assert D.sum(axis=1).max() <= L5 * P5 + 1e-6
        assert runs[1][1] == lls
        assert np.array_equal(runs[1][2][0], runs[0][2][0])
        assert np.array_equal(runs[1][2][1], runs[0][2][1])
    finally:
        ctx.close()


def test_mixture_multiallelic_digit_against_gather_at_scale():
    """the mixture model on 20k individuals x 10k loci with up to 20 alleles (T ~ 1.2e5 allele
    columns, 5 % missing): the digit-sliced integer kernels on column pairs against the
    two-pass gather kernel's mixture modes over three EM steps and a log-likelihood pass --
    several chunks and blocks per CTA, 32-bit accumulators far from their bound -- plus the
    usual size-independent properties"""
    from multiclust_b200 import Context, SynthParams
    I3, L3, K3 = 20000, 10000, 10
    out = []
    for kernel, family in ((0, 4), (2, 2)):
        ctx = Context(0)
        try:
            ctx.set_option(ctx.OPT_KERNEL, kernel)
            ctx.set_data_synth(I3, L3, SynthParams(seed=20261018, K=K3, jmax=20, miss_bp=500,
                                                   ploidy=2))
            lb = min(1e-8, 0.5 / I3 / 2)
            ctx.alloc_model(K3, admixture=0, q=0, eta_lb=lb, p_lb=lb)
            assert ctx.plan()["two_pass"] == family
            eta0, p0, J, seg = _start(ctx, K3, False, 21)
            ctx.set_params(0, eta0, p0)
            pre = ctx.loglik(0)
            lls = [ctx.em_step(0, 0) for _ in range(3)]
            out.append((pre, lls, ctx.get_params(0), ctx.posterior(), ctx.loglik(0)))
        finally:
            ctx.close()
    (pre, lls, (eta, p), v, post_ll), (pre2, lls2, (eta2, p2), v2, post_ll2) = out
    assert np.all(np.isfinite(lls)) and abs(pre - lls[0]) <= 1e-12 * abs(pre)
    assert all(b > a for a, b in zip(lls, lls[1:])) and post_ll > lls[-1]
    assert abs(eta.sum() - 1.0) < 1e-12 and np.max(np.abs(v.sum(axis=1) - 1.0)) < 1e-12
    assert abs(pre - pre2) <= 1e-12 * abs(pre) and abs(post_ll - post_ll2) <= 1e-12 * abs(pre)
    assert np.max(np.abs(np.array(lls) - np.array(lls2)) / np.abs(lls)) < 1e-12
    assert np.max(np.abs(eta - eta2)) < 1e-11 and np.max(np.abs(p - p2)) < 1e-11
    assert np.max(np.abs(v - v2)) < 1e-9
