"""GPU parity of every C-ABI entry point against the CPU oracle on the same
seeded inputs (sizes the oracle finishes in well under a second).

Tolerances are the north star's: log likelihood within 1e-9 relative,
parameters / posterior sums within 1e-7 absolute.  Single calls are expected
to agree far more tightly; the tight bounds asserted here (1e-12 / 1e-11)
document how much slack the multi-iteration tests have."""
import numpy as np
import pytest

from common import gen_data, random_params

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9
PAR_ATOL = 1e-7
TIGHT = 1e-11

SHAPES = [
    # I, L, K, jmax, miss_bp, P
    (60, 40, 3, 5, 300, 2),
    (37, 133, 10, 20, 500, 2),     # config-3 shaped: K=10, <=20 alleles, 5 % missing
    (45, 70, 8, 2, 0, 4),          # config-5 shaped: biallelic tetraploid
    (33, 50, 5, 2, 200, 2),
    (9, 7, 2, 3, 1000, 1),         # haploid, fewer individuals than a block
    (20, 300, 17, 6, 100, 3),      # K > one lane's share, odd ploidy
    (130, 20, 40, 4, 0, 2),        # large K (k_split > 1 with padding)
]


@pytest.fixture(scope="module")
def ctx():
    from multiclust_b200 import Context
    c = Context(0)
    yield c
    c.close()


def setup_pair(orc, ctx, tmp_path, shape, admixture, eta_constrained=0, q=0,
               accel=0, proj=1):
    I, L, K, jmax, miss, P = shape
    d = gen_data(tmp_path, I, L, K=min(K, 6), jmax=jmax, miss=miss, P=P)
    fit = orc.Fit(d["J"], d["codes"], admixture=admixture,
                  eta_constrained=eta_constrained, accel=accel, do_projection=proj)
    fit.alloc(K)
    ctx.set_data(d["J"], d["codes"])
    lb = fit.lower_bound
    ctx.alloc_model(K, admixture=admixture, eta_constrained=eta_constrained,
                    q=q, eta_lb=lb, p_lb=lb, do_projection=proj)
    rng = np.random.default_rng(1234 + I + L)
    per_indiv = bool(admixture and not eta_constrained)
    eta, p = random_params(rng, I, K, d["J"], per_indiv)
    fit.set_params(0, eta, p)
    ctx.set_params(0, eta, p)
    return d, fit, eta, p


def check_step(orc, ctx, fit, frm, to):
    fit.set_indices(0, frm, to)
    ll_o = fit.e_step()
    fit.m_step()
    ll_g = ctx.em_step(frm, to)
    assert abs(ll_g - ll_o) <= LL_RTOL * abs(ll_o)
    assert abs(ll_g - ll_o) <= 1e-12 * abs(ll_o) + 1e-12
    eo, po = fit.get_params(to)
    eg, pg = ctx.get_params(to)
    assert np.max(np.abs(eg - eo)) < TIGHT
    assert np.max(np.abs(pg - po)) < TIGHT
    assert np.max(np.abs(ctx.posterior() - fit.posterior())) < 1e-9
    return ll_o


@pytest.mark.parametrize("shape", SHAPES)
def test_admixture_em_step_and_loglik(orc, ctx, tmp_path, shape):
    d, fit, eta, p = setup_pair(orc, ctx, tmp_path, shape, admixture=1)
    ll_o = fit.log_likelihood(0)
    ll_g = ctx.loglik(0)
    assert abs(ll_g - ll_o) <= 1e-12 * abs(ll_o)
    check_step(orc, ctx, fit, 0, 1)       # out of place
    check_step(orc, ctx, fit, 1, 1)       # in place, like unaccelerated EM
    check_step(orc, ctx, fit, 1, 2)
    # the log likelihood of the new slot, and the posterior untouched by it
    post = ctx.posterior()
    assert abs(ctx.loglik(2) - fit.log_likelihood(2)) <= 1e-12 * abs(ll_o)
    assert np.array_equal(post, ctx.posterior())


@pytest.mark.parametrize("shape", SHAPES[:4])
def test_admixture_pooled_eta(orc, ctx, tmp_path, shape):
    d, fit, eta, p = setup_pair(orc, ctx, tmp_path, shape, admixture=1, eta_constrained=1)
    check_step(orc, ctx, fit, 0, 0)
    check_step(orc, ctx, fit, 0, 1)


@pytest.mark.parametrize("shape", SHAPES)
def test_mixture_em_step_and_loglik(orc, ctx, tmp_path, shape):
    d, fit, eta, p = setup_pair(orc, ctx, tmp_path, shape, admixture=0)
    ll_o = fit.log_likelihood(0)
    assert abs(ctx.loglik(0) - ll_o) <= 1e-12 * abs(ll_o)
    check_step(orc, ctx, fit, 0, 1)
    check_step(orc, ctx, fit, 1, 1)
    ik, cnt = ctx.partition()
    assert np.array_equal(ik, np.argmax(fit.posterior(), axis=1))
    assert cnt.sum() == d["I"]


def test_no_projection(orc, ctx, tmp_path):
    d, fit, eta, p = setup_pair(orc, ctx, tmp_path, SHAPES[0], admixture=1, proj=0)
    check_step(orc, ctx, fit, 0, 0)


def test_projection_kernel(orc, ctx, tmp_path):
    """simplex_project_eta / simplex_project_pklm on rows that need clamping"""
    d, fit, eta, p = setup_pair(orc, ctx, tmp_path, SHAPES[0], admixture=1)
    rng = np.random.default_rng(7)
    K, I, J = 3, d["I"], d["J"]
    eta2 = eta + rng.normal(0, 0.4, eta.size)
    p2 = p + rng.normal(0, 0.4, p.size)
    ctx.set_params(1, eta2, p2)
    ctx.project(1)
    eg, pg = ctx.get_params(1)
    lb = fit.lower_bound
    eo = np.concatenate([orc.project(eta2[i * K:(i + 1) * K], lb) for i in range(I)])
    T = int(J.sum())
    off = np.concatenate([[0], np.cumsum(J)])
    po = p2.copy()
    for k in range(K):
        for l in range(len(J)):
            a, b = k * T + off[l], k * T + off[l + 1]
            po[a:b] = orc.project(p2[a:b], lb)
    assert np.array_equal(eg, eo)       # same operations in the same order
    assert np.array_equal(pg, po)


def test_secant_pairs_and_updates(orc, ctx, tmp_path):
    """mc_delta, mc_step_dots, mc_qn_dots, mc_accel_update, mc_qn_update"""
    shape = SHAPES[1]
    I, L, K = shape[0], shape[1], shape[2]
    d, fit, eta, p = setup_pair(orc, ctx, tmp_path, shape, admixture=1, q=2, accel=5)
    rng = np.random.default_rng(99)
    T = fit.T
    xs = []
    for s in range(3):
        e, pp = random_params(rng, I, K, d["J"], True)
        ctx.set_params(s, e, pp)
        xs.append((e, pp))
    ctx.delta(0, 0, 1, 0)   # u0 = x1 - x0
    ctx.delta(1, 0, 2, 1)   # v0 = x2 - x1
    ctx.delta(0, 1, 2, 0)   # u1 = x2 - x0
    ctx.delta(1, 1, 0, 1)   # v1 = x0 - x1
    ue0, up0 = xs[1][0] - xs[0][0], xs[1][1] - xs[0][1]
    ve0, vp0 = xs[2][0] - xs[1][0], xs[2][1] - xs[1][1]
    ue1, up1 = xs[2][0] - xs[0][0], xs[2][1] - xs[0][1]
    ve1, vp1 = xs[0][0] - xs[1][0], xs[0][1] - xs[1][1]
    e3, p3 = ctx.step_dots(0)
    ref_e = [ue0 @ ue0, ue0 @ (ve0 - ue0), (ve0 - ue0) @ (ve0 - ue0)]
    ref_p = [up0 @ up0, up0 @ (vp0 - up0), (vp0 - up0) @ (vp0 - up0)]
    assert np.allclose(e3, ref_e, rtol=1e-12, atol=0)
    assert np.allclose(p3, ref_p, rtol=1e-12, atol=0)
    e2, p2 = ctx.qn_dots(0, 1)
    assert np.allclose(e2, [ue0 @ ue1, ue0 @ ve1], rtol=1e-11, atol=1e-14)
    assert np.allclose(p2, [up0 @ up1, up0 @ vp1], rtol=1e-11, atol=1e-14)
    # SQUAREM and QN1 extrapolations (projection applied afterwards)
    lb = fit.lower_bound
    off = np.concatenate([[0], np.cumsum(d["J"])])

    def project_all(e, pp):
        e = np.concatenate([orc.project(e[i * K:(i + 1) * K], lb) for i in range(I)])
        pp = pp.copy()
        for k in range(K):
            for l in range(L):
                a, b = k * T + off[l], k * T + off[l + 1]
                pp[a:b] = orc.project(pp[a:b], lb)
        return e, pp

    s = -1.7
    for qn1 in (0, 1):
        ctx.accel_update(qn1, 2, 0, 0, s)
        eg, pg = ctx.get_params(2)
        if qn1:
            er = xs[0][0] + ue0 + s * ve0
            pr = xs[0][1] + up0 + s * vp0
        else:
            er = xs[0][0] - 2 * s * ue0 + s * s * (ve0 - ue0)
            pr = xs[0][1] - 2 * s * up0 + s * s * (vp0 - up0)
        er, pr = project_all(er, pr)
        assert np.array_equal(eg, er)
        assert np.array_equal(pg, pr)
    # QN q=2 update: x[p] + u[uindex] + sum_jn v[(delta+j)%q] * Ainv[j][n] * cutu[n]
    Ainv = np.array([[0.3, -0.2], [0.15, 0.4]])
    cutu = np.array([0.7, -0.1])
    ctx.set_params(2, xs[2][0], xs[2][1])
    ctx.qn_update(1, 0, 1, 1, Ainv, cutu)
    eg, pg = ctx.get_params(1)
    er, pr = xs[0][0] + ue1, xs[0][1] + up1
    vs_e, vs_p = [ve1, ve0], [vp1, vp0]       # rows start at delta_index = 1
    for j in range(2):
        for n in range(2):
            er = er + vs_e[j] * Ainv[j, n] * cutu[n]
            pr = pr + vs_p[j] * Ainv[j, n] * cutu[n]
    er, pr = project_all(er, pr)
    assert np.array_equal(eg, er)
    assert np.array_equal(pg, pr)


def test_determinism(orc, ctx, tmp_path):
    """bitwise repeatable: fixed reduction order, no float atomics"""
    d, fit, eta, p = setup_pair(orc, ctx, tmp_path, SHAPES[1], admixture=1)
    out = []
    for rep in range(3):
        ctx.set_params(0, eta, p)
        lls = [ctx.em_step(0, 0) for _ in range(4)]
        e, pp = ctx.get_params(0)
        out.append((lls, e.copy(), pp.copy()))
    for rep in (1, 2):
        assert out[rep][0] == out[0][0]
        assert np.array_equal(out[rep][1], out[0][1])
        assert np.array_equal(out[rep][2], out[0][2])


def test_split_step_equals_whole_step(orc, ctx, tmp_path):
    """mc_em_step == mc_em_step_local + (exchange) + mc_em_step_finish"""
    d, fit, eta, p = setup_pair(orc, ctx, tmp_path, SHAPES[0], admixture=1)
    ll1 = ctx.em_step(0, 1)
    e1, p1 = ctx.get_params(1)
    ctx.em_step_local(0, 2)
    ptr, n = ctx.exchange_buffer()
    assert n == 3 * fit.T + 1 + 3
    ll2 = ctx.em_step_finish(2)
    e2, p2 = ctx.get_params(2)
    assert ll1 == ll2 and np.array_equal(e1, e2) and np.array_equal(p1, p2)


def test_synthetic_generator_matches_host(orc, ctx, tmp_path):
    """mc_set_data_synth (device fill) == mc_gen (host C) byte for byte"""
    from multiclust_b200 import SynthParams
    I, L = 70, 90
    d = gen_data(tmp_path, I, L, K=4, jmax=7, miss=400, P=2)
    sp = SynthParams(seed=20261018, K=4, jmax=7, miss_bp=400, ploidy=2)
    ctx.set_data_synth(I, L, sp)
    assert np.array_equal(ctx.get_J(), d["J"])
    assert np.array_equal(ctx.get_codes(), d["codes"])


# ---- the dense DMMA kernels (mc_dense.cuh) and the planner's kernel choice ----

DENSE_SHAPES = [
    # I, L, K, jmax, miss_bp, P: biallelic everywhere
    (45, 70, 8, 2, 0, 4),          # config-5 shaped
    (300, 37, 5, 2, 300, 2),       # more than one tile of individuals, ragged locus tile, missing
    (70, 530, 3, 2, 100, 1),       # haploid, many locus tiles
    (64, 48, 12, 2, 200, 3),       # K > 8: two blocks of clusters, odd ploidy
    (33, 40, 8, 2, 500, 6),        # hexaploid (3-bit counts)
]


@pytest.mark.parametrize("shape", DENSE_SHAPES)
@pytest.mark.parametrize("admixture", [1, 0])
def test_dense_kernels(orc, tmp_path, shape, admixture):
    from multiclust_b200 import Context
    c = Context(0)
    try:
        c.set_option(c.OPT_KERNEL, c.KERNEL_DENSE)
        d, fit, eta, p = setup_pair(orc, c, tmp_path, shape, admixture=admixture)
        assert c.plan()["two_pass"] == 3
        ll_o = fit.log_likelihood(0)
        assert abs(c.loglik(0) - ll_o) <= 1e-12 * abs(ll_o)
        check_step(orc, c, fit, 0, 1)
        check_step(orc, c, fit, 1, 1)
        post = c.posterior()
        assert abs(c.loglik(1) - fit.log_likelihood(1)) <= 1e-12 * abs(ll_o)
        assert np.array_equal(post, c.posterior())
    finally:
        c.close()


# ---- the digit-sliced integer kernels of the mixture model (mc_digit.cuh) ----

DIGIT_SHAPES = DENSE_SHAPES + [
    (530, 300, 5, 2, 300, 2),      # several blocks of loci (64) and of individuals (128)
    (129, 65, 1, 2, 0, 2),         # K = 1, one element past a block in both directions
    (200, 150, 16, 2, 200, 2),     # K = 16 (one m-tile per warp)
    (90, 130, 7, 2, 100, 15),      # the largest ploidy of the packed counts
    (150, 100, 10, 2, 0, 2),       # BASELINE K
]


@pytest.mark.parametrize("shape", DIGIT_SHAPES)
def test_digit_kernels(orc, tmp_path, shape):
    """mixture model on biallelic data: the planner picks the digit-sliced kernels; E-step,
    M-step and log likelihood against the oracle, the posterior survives mc_loglik"""
    from multiclust_b200 import Context
    c = Context(0)
    try:
        d, fit, eta, p = setup_pair(orc, c, tmp_path, shape, admixture=0)
        assert c.plan()["two_pass"] == 4
        ll_o = fit.log_likelihood(0)
        assert abs(c.loglik(0) - ll_o) <= 1e-12 * abs(ll_o)
        check_step(orc, c, fit, 0, 1)
        check_step(orc, c, fit, 1, 1)
        check_step(orc, c, fit, 1, 2)
        post = c.posterior()
        assert abs(c.loglik(2) - fit.log_likelihood(2)) <= 1e-12 * abs(ll_o)
        assert np.array_equal(post, c.posterior())
    finally:
        c.close()


@pytest.mark.parametrize("K", [2, 3, 4, 6, 9, 11, 12, 13, 14, 15])
def test_digit_kernels_every_k(orc, tmp_path, K):
    from multiclust_b200 import Context
    c = Context(0)
    try:
        d, fit, eta, p = setup_pair(orc, c, tmp_path, (70, 90, K, 2, 300, 2), admixture=0)
        assert c.plan()["two_pass"] == 4
        ll_o = fit.log_likelihood(0)
        assert abs(c.loglik(0) - ll_o) <= 1e-12 * abs(ll_o)
        check_step(orc, c, fit, 0, 1)
    finally:
        c.close()


def test_digit_kernels_agree_with_dmma(orc, tmp_path):
    """the integer path against the FP64 tensor path on the same start: the sums are exact
    on one side and rounded on the other, so they agree to a few ulp of the sums"""
    from multiclust_b200 import Context
    out = []
    for kernel in (3, 4):
        c = Context(0)
        try:
            c.set_option(c.OPT_KERNEL, kernel)
            d, fit, eta, p = setup_pair(orc, c, tmp_path, DIGIT_SHAPES[5], admixture=0)
            assert c.plan()["two_pass"] == kernel
            lls = [c.em_step(0, 0) for _ in range(5)]
            out.append((lls, c.get_params(0), c.posterior(), c.loglik(0)))
        finally:
            c.close()
    (l3, (e3, p3), v3, f3), (l4, (e4, p4), v4, f4) = out
    assert np.max(np.abs(np.array(l3) - np.array(l4)) / np.abs(l3)) < 1e-13
    assert abs(f3 - f4) <= 1e-13 * abs(f3)
    assert np.max(np.abs(e3 - e4)) < 1e-12 and np.max(np.abs(p3 - p4)) < 1e-12
    assert np.max(np.abs(v3 - v4)) < 1e-10


DIGIT_GENERAL_SHAPES = [
    # I, L, K, jmax, miss_bp, P: more than two alleles -> the column-pair form
    (300, 37, 5, 6, 300, 2),
    (70, 130, 3, 20, 100, 1),      # haploid, up to 20 alleles
    (129, 65, 16, 5, 0, 2),        # K = 16
    (90, 50, 7, 4, 100, 4),        # tetraploid
    (530, 300, 5, 12, 300, 2),     # several blocks of column pairs and of individuals
    (40, 30, 1, 3, 500, 3),        # K = 1, odd ploidy
]


@pytest.mark.parametrize("shape", DIGIT_GENERAL_SHAPES)
def test_digit_kernels_multiallelic(orc, tmp_path, shape):
    """mixture model on multi-allelic data: the digit-sliced kernels on column pairs"""
    from multiclust_b200 import Context
    c = Context(0)
    try:
        d, fit, eta, p = setup_pair(orc, c, tmp_path, shape, admixture=0)
        assert int(d["J"].max()) > 2 and c.plan()["two_pass"] == 4
        ll_o = fit.log_likelihood(0)
        assert abs(c.loglik(0) - ll_o) <= 1e-12 * abs(ll_o)
        check_step(orc, c, fit, 0, 1)
        check_step(orc, c, fit, 1, 1)
        check_step(orc, c, fit, 1, 2)
        post = c.posterior()
        assert abs(c.loglik(2) - fit.log_likelihood(2)) <= 1e-12 * abs(ll_o)
        assert np.array_equal(post, c.posterior())
    finally:
        c.close()


@pytest.mark.parametrize("shape", [(300, 120, 4, 7, 300, 2), (60, 40, 3, 5, 300, 12)])
def test_digit_multiallelic_log_zero_falls_back(orc, tmp_path, shape):
    """the log-likelihood pass with p == 0 on multi-allelic data: the FP64 kernel of the plan
    underneath -- the two-pass gather kernel, or the one-pass kernel at ploidy 12 -- runs
    behind the declined digit pass, on the device"""
    from multiclust_b200 import Context
    I, L, K = shape[:3]
    c = Context(0)
    try:
        d, fit, eta, p = setup_pair(orc, c, tmp_path, shape, admixture=0, proj=0)
        assert c.plan()["two_pass"] == 4
        J = d["J"]
        off = np.concatenate([[0], np.cumsum(J)])
        p = p.reshape(K, -1).copy()
        for l in (0, 7, L - 1):
            if J[l] >= 2:               # class 2 cannot carry allele 0 of these loci
                s_ = p[2, off[l]] + p[2, off[l] + 1]
                p[2, off[l]] = 0.0
                p[2, off[l] + 1] = s_
        p = p.ravel()
        fit.set_params(0, eta, p)
        c.set_params(0, eta, p)
        ll_o = fit.log_likelihood(0)
        assert np.isfinite(ll_o)
        for rep in range(3):            # eager, captured and replayed launch sequence
            assert abs(c.loglik(0) - ll_o) <= 1e-12 * abs(ll_o)
        check_step(orc, c, fit, 0, 1)   # E-step: p == 0 contributes nothing
        ll_1 = fit.log_likelihood(1)
        assert abs(c.loglik(1) - ll_1) <= 1e-12 * abs(ll_1)
    finally:
        c.close()


def test_digit_nan_parameter_is_reported(orc, tmp_path):
    """a NaN in p has no fixed-point image: the E-step on the digit-sliced kernels must not
    hide it (the FP64 kernels propagate it by arithmetic) -- the step's log likelihood is NaN,
    and the flag is gone for the next call"""
    from multiclust_b200 import Context
    c = Context(0)
    try:
        d, fit, eta, p = setup_pair(orc, c, tmp_path, DIGIT_SHAPES[1], admixture=0)
        assert c.plan()["two_pass"] == 4
        bad = p.copy()
        bad[3] = np.nan
        c.set_params(1, eta, bad)
        assert np.isnan(c.em_step(1, 2))
        assert np.isnan(c.loglik(1))            # falls back: the FP64 kernels propagate it
        ll_o = fit.log_likelihood(0)
        assert abs(c.loglik(0) - ll_o) <= 1e-12 * abs(ll_o)
        check_step(orc, c, fit, 0, 1)
    finally:
        c.close()


def test_digit_log_zero_falls_back(orc, tmp_path):
    """log 0 has no fixed-point form: without the projection p reaches 0, the log-likelihood
    pass falls back to the FP64 kernels on the device (no host round trip) and the next pass
    is back on the integer kernels"""
    from multiclust_b200 import Context
    c = Context(0)
    try:
        shape = (300, 200, 4, 2, 300, 2)
        d, fit, eta, p = setup_pair(orc, c, tmp_path, shape, admixture=0, proj=0)
        assert c.plan()["two_pass"] == 4
        J = d["J"]
        off = np.concatenate([[0], np.cumsum(J)])
        T = int(J.sum())
        p = p.reshape(4, T).copy()
        for l in (0, 7, 150):               # class 2 cannot carry allele 0 of these loci
            if J[l] >= 2:
                p[2, off[l]] = 0.0
                p[2, off[l] + 1] = 1.0 if J[l] == 2 else p[2, off[l] + 1]
        p = p.ravel()
        fit.set_params(0, eta, p)
        c.set_params(0, eta, p)
        ll_o = fit.log_likelihood(0)
        assert np.isfinite(ll_o)
        assert abs(c.loglik(0) - ll_o) <= 1e-12 * abs(ll_o)
        check_step(orc, c, fit, 0, 1)       # E-step: p == 0 contributes nothing
        ll_1 = fit.log_likelihood(1)
        assert abs(c.loglik(1) - ll_1) <= 1e-12 * abs(ll_1)
    finally:
        c.close()


def test_dense_pooled_eta(orc, tmp_path):
    from multiclust_b200 import Context
    c = Context(0)
    try:
        d, fit, eta, p = setup_pair(orc, c, tmp_path, DENSE_SHAPES[1], admixture=1,
                                    eta_constrained=1)
        assert c.plan()["two_pass"] == 3
        check_step(orc, c, fit, 0, 0)
        check_step(orc, c, fit, 0, 1)
    finally:
        c.close()


@pytest.mark.parametrize("kernel,two_pass", [(1, 0), (2, 2), (3, 3), (0, 3)])
def test_kernel_option(orc, tmp_path, kernel, two_pass):
    """mc_set_option(MC_OPT_KERNEL): every kernel family gives the oracle's step
    on biallelic data; the planner reports which one ran"""
    from multiclust_b200 import Context
    c = Context(0)
    try:
        c.set_option(c.OPT_KERNEL, kernel)
        d, fit, eta, p = setup_pair(orc, c, tmp_path, DENSE_SHAPES[0], admixture=1)
        assert c.plan()["two_pass"] == two_pass
        check_step(orc, c, fit, 0, 1)
    finally:
        c.close()


def test_multiallelic_data_never_dense(orc, tmp_path):
    from multiclust_b200 import Context
    c = Context(0)
    try:
        c.set_option(c.OPT_KERNEL, c.KERNEL_DENSE)
        d, fit, eta, p = setup_pair(orc, c, tmp_path, SHAPES[0], admixture=1)
        assert c.plan()["two_pass"] == 0      # falls back to the one-pass kernel
        check_step(orc, c, fit, 0, 1)
    finally:
        c.close()


def test_locale_sums_and_mixture_init(orc, tmp_path):
    """mc_locale_sums (popq numerators) and mc_init_mixture (nearest-centre start)
    against numpy restatements of write_file.c:446-459 / rnd_init.c:220-338"""
    from multiclust_b200 import Context
    I, L, K, P = 75, 33, 4, 2
    d = gen_data(tmp_path, I, L, K=3, jmax=5, miss=400, P=P)
    codes, J = d["codes"], d["J"]
    off = np.concatenate([[0], np.cumsum(J)])
    T = int(J.sum())
    c = Context(0)
    try:
        c.set_data(J, codes)
        c.alloc_model(K, admixture=0, q=0)
        centers = np.array([17, 3, 60, 41], dtype=np.int32)
        c.init_mixture(0, centers, codes[centers])
        eta, p = c.get_params(0)
        # numpy restatement
        cnt = np.zeros((I, T))
        for i in range(I):
            for l in range(L):
                for a in range(P):
                    if codes[i, l, a] != 255:
                        cnt[i, off[l] + codes[i, l, a]] += 1
        part = np.zeros(I, dtype=int)
        for i in range(I):
            if i == centers[0]:
                continue
            best = np.inf
            for k in range(K):
                if i == centers[k]:
                    part[i] = k
                    break
                dist = np.abs(cnt[i] - cnt[centers[k]]).sum()
                if dist < best:
                    best, part[i] = dist, k
        eta_r = (1.0 + np.bincount(part, minlength=K)) / (I + K)
        p_r = np.zeros((K, T))
        for k in range(K):
            S = cnt[part == k].sum(axis=0)
            row = 1.0 + (K - k) * S
            for l in range(L):
                a, b = off[l], off[l + 1]
                if b > a:
                    p_r[k, a:b] = row[a:b] / row[a:b].sum()
        assert np.array_equal(eta, eta_r)
        assert np.max(np.abs(p - p_r.ravel())) < 1e-15
        # locale sums of the posterior after one step
        c.em_step(0, 0)
        post = c.posterior()
        locale = (np.arange(I) * 7 % 5).astype(np.int32)
        got = c.locale_sums(locale, 5)
        ref = np.stack([post[locale == n].sum(axis=0) for n in range(5)])
        assert np.max(np.abs(got - ref)) < 1e-12
    finally:
        c.close()


@pytest.mark.parametrize("P", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("admixture", [1, 0])
def test_every_gather_kernel_instantiation(orc, tmp_path, P, admixture):
    """every (K pair count, padded ploidy) instantiation of the two-pass gather kernel
    -- admixture E+M / log likelihood and the mixture E / M modes -- runs one EM step
    and one log-likelihood pass against the oracle (a miscompiled instantiation shows up
    as a launch failure or a wrong number, not as an untested path)"""
    from multiclust_b200 import Context
    I, L = 40, 24
    d = gen_data(tmp_path, I, L, K=3, jmax=6, miss=300, P=P)
    c = Context(0)
    try:
        c.set_option(c.OPT_KERNEL, c.KERNEL_ADMIX3)
        c.set_data(d["J"], d["codes"])
        for K in range(1, 17):
            fit = orc.Fit(d["J"], d["codes"], admixture=admixture)
            fit.alloc(K)
            lb = fit.lower_bound
            c.alloc_model(K, admixture=admixture, q=0, eta_lb=lb, p_lb=lb)
            assert c.plan()["two_pass"] == 2
            eta, p = random_params(np.random.default_rng(100 * K + P), I, K, d["J"], bool(admixture))
            fit.set_params(0, eta, p)
            c.set_params(0, eta, p)
            ll_o = fit.log_likelihood(0)
            assert abs(c.loglik(0) - ll_o) <= 1e-12 * abs(ll_o), (K, P)
            check_step(orc, c, fit, 0, 1)
    finally:
        c.close()


@pytest.mark.parametrize("admixture", [1, 0])
@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_degenerate_loci(orc, admixture, kernel):
    """hand-made biallelic data with a monomorphic locus (J = 1), a locus where everybody is
    missing (J = 0, read_file.c:525-526), a locus with one observed allele plus the phantom
    slot (J = 2) and an individual without any data at several loci: every kernel family
    against the oracle (kernel 0 picks the dense DMMA path here)"""
    from multiclust_b200 import Context
    rng = np.random.default_rng(42)
    I, L, P, K = 37, 21, 2, 3
    codes = rng.integers(0, 2, size=(I, L, P)).astype(np.uint8)
    miss = rng.random((I, L, P)) < 0.1
    codes[miss] = 255
    codes[5, :10, :] = 255                  # an individual with little data
    codes[:, 3, :] = 0                      # monomorphic, nobody missing: J = 1
    codes[:, 7, :] = 255                    # all missing: J = 0
    codes[:, 11, :] = 0
    codes[::5, 11, 0] = 255                 # one allele + phantom slot: J = 2
    J = np.zeros(L, dtype=np.int32)
    for l in range(L):
        obs = codes[:, l, :][codes[:, l, :] != 255]
        nreal = int(obs.max()) + 1 if obs.size else 0
        # the allele codes must be dense: relabel a locus that only shows allele 1
        if nreal == 2 and not (obs == 0).any():
            codes[:, l, :][codes[:, l, :] == 1] = 0
            nreal = 1
        J[l] = nreal + (1 if nreal and (codes[:, l, :] == 255).any() else 0)
    assert J[3] == 1 and J[7] == 0 and J[11] == 2
    fit = orc.Fit(J, codes, admixture=admixture)
    fit.alloc(K)
    lb = fit.lower_bound
    c = Context(0)
    try:
        c.set_option(c.OPT_KERNEL, kernel)
        c.set_data(J, codes)
        c.alloc_model(K, admixture=admixture, q=0, eta_lb=lb, p_lb=lb)
        # kernel 0: dense DMMA (admixture) / digit-sliced integer kernels (mixture)
        assert c.plan()["two_pass"] == {0: 3 if admixture else 4, 1: 0, 2: 2}[kernel]
        eta, p = random_params(np.random.default_rng(5), I, K, J, bool(admixture))
        fit.set_params(0, eta, p)
        c.set_params(0, eta, p)
        ll_o = fit.log_likelihood(0)
        assert abs(c.loglik(0) - ll_o) <= 1e-12 * abs(ll_o)
        check_step(orc, c, fit, 0, 1)
        check_step(orc, c, fit, 1, 1)
    finally:
        c.close()
