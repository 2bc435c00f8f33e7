"""Pin the CPU oracle (oracle/mc_oracle.c) on the golden vectors written by the
unmodified reference (tests/golden/make_golden.py): initial parameters, every
log likelihood handed to stop(), the final parameters and posterior sums must
be BIT-identical.  This is what lets the GPU tests use the oracle as the
checker at sizes and option mixes the fixtures do not cover."""
import numpy as np
import pytest

from common import golden_names, load_golden


def make_fit(orc, g):
    o = g["meta"]["options"]
    return orc.Fit(g["J"], g["codes"], admixture=o["admixture"],
                   eta_constrained=o["eta_constrained"], accel=o["accel"],
                   do_projection=o["do_projection"], max_iter=o["max_iter"],
                   abs_error=o["abs_error"], rel_error=o["rel_error"])


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference(orc, name):
    g = load_golden(name)
    o = g["meta"]["options"]
    # glibc rand(): the reference never seeds unless -r is given
    # (multiclust.c:1592-1596); seed 1 is the libc default state
    orc.seed(o["seed"] if o["seed"] >= 0 else 1)
    fit = make_fit(orc, g)
    assert fit.lower_bound == g["meta"]["bound"]
    last_K = None
    for rec in g["meta"]["fits"]:
        K, init = rec["K"], rec["init"]
        if K != last_K:
            fit.alloc(K)
            last_K = K
        key = "K%d_i%d_" % (K, init)
        fit.reset_trace()
        fit.initialize()
        eta, p = fit.get_params(0)
        assert np.array_equal(eta, g[key + "start_eta"]), "initial eta"
        assert np.array_equal(p, g[key + "start_p"]), "initial p"
        fit.em()
        st = fit.state()
        assert np.array_equal(fit.trace(), g[key + "ll"]), "log likelihood trace"
        assert st["logL"] == rec["logL"]
        assert st["n_iter"] == rec["n_iter"]
        assert st["converged"] == rec["converged"]
        assert st["iter_stop"] == rec["iter_stop"]
        assert st["pindex"] == rec["pindex"]
        eta, p = fit.get_params(st["pindex"])
        assert np.array_equal(eta, g[key + "final_eta"])
        assert np.array_equal(p, g[key + "final_p"])
        assert np.array_equal(fit.posterior(), g[key + "final_post"])


def test_projection_matches_hand_case(orc):
    # simplex.c:109-143: csum includes clamped coordinates, n shrinks inside
    # the sweep while the shift stays that sweep's value
    x = orc.project([0.7, 0.5, -0.1, 0.2], 1e-8)
    assert abs(x.sum() - 1.0) < 1e-12
    assert x.min() >= 1e-8
    y = orc.project([0.25, 0.25, 0.25, 0.25], 1e-8)
    assert np.array_equal(y, np.array([0.25] * 4))
