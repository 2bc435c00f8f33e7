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
