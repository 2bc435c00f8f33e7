"""The N > 1 path on CPU: two gloo ranks, each holding half of the individuals,
run the individual-sharded EM step of multiclust_b200/sharding.py with the
oracle standing in for the device (the CUDA context offers the same four
calls).  The result must equal the unsharded oracle step: eta rows exactly
(they are local), p and the log likelihood to rounding (the sums over
individuals are split in two)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import gen_data, random_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShard:
    """adapter: the four calls of a shard on top of the CPU oracle"""

    def __init__(self, orc, J, codes, admixture, eta_constrained, K, eta, p, slices=True):
        self.fit = orc.Fit(J, codes, admixture=admixture, eta_constrained=eta_constrained)
        self.fit.alloc(K)
        self.fit.set_params(0, eta, p)
        self.K, self.T = K, self.fit.T
        self.n = K * self.T + 1 + K
        self.buf = torch.zeros(self.n + 64, dtype=torch.float64)     # like mc_exchange_buffer
        self.slices = slices

    def em_step_local(self, frm, to):
        self.fit.set_indices(0, frm, to)
        ll = self.fit.e_step()
        N, S = self.fit.sums()
        self.buf[:self.K * self.T] = torch.from_numpy(N)
        self.buf[self.K * self.T] = ll
        self.buf[self.K * self.T + 1:self.n] = torch.from_numpy(S)

    def exchange_tensor(self):
        return self.buf

    def exchange_len(self):
        return self.n

    def __getattr__(self, name):
        # the slice path is offered only when asked for (both paths are tested)
        if name == "sum_slices" and self.__dict__.get("slices"):
            return self._sum_slices
        raise AttributeError(name)

    def _sum_slices(self, parts, world, first, count):
        total = parts[:count].clone()
        for r in range(1, world):          # rank order: deterministic
            total += parts[r * count:(r + 1) * count]
        self.buf[first:first + count] = total

    def sum_gathered(self, gathered, world):
        n = self.n
        total = gathered[:n].clone()
        for r in range(1, world):          # rank order: deterministic
            total += gathered[r * n:(r + 1) * n]
        self.buf[:n] = total

    def em_step_finish(self, to):
        kt = self.K * self.T
        self.fit.m_step_from_sums(self.buf[:kt].numpy(), self.buf[kt + 1:self.n].numpy())
        return float(self.buf[kt])


def _worker(rank, world, port, mcb_path, admixture, eta_constrained, K, out, slices=True):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import orc
    from multiclust_b200.sharding import shard_bounds, sharded_em_step
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = orc.read_mcb(mcb_path)
    I = d["I"]
    per_indiv = bool(admixture and not eta_constrained)
    eta, p = random_params(np.random.default_rng(5), I, K, d["J"], per_indiv)
    lo, hi = shard_bounds(I, world)[rank]
    eta_local = eta.reshape(I, K)[lo:hi].ravel() if per_indiv else eta
    shard = OracleShard(orc, d["J"], d["codes"][lo:hi], admixture, eta_constrained, K,
                        eta_local, p, slices=slices)
    gathered = torch.zeros(world * shard.buf.numel(), dtype=torch.float64)
    lls = []
    for it in range(3):
        lls.append(sharded_em_step(shard, dist, world, 0, 0, gathered))
    e, pp = shard.fit.get_params(0)
    np.savez(os.path.join(out, "rank%d.npz" % rank), ll=np.array(lls), eta=e, p=pp,
             lo=lo, hi=hi)
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("admixture,eta_constrained,slices,world",
                         [(1, 0, True, 2), (1, 1, True, 2), (0, 0, True, 2), (1, 0, False, 2),
                          (1, 0, True, 3)])
def test_two_rank_step_equals_single(orc, tmp_path, admixture, eta_constrained, slices, world):
    K = 4
    d = gen_data(tmp_path, 41, 30, K=3, jmax=6, miss=300, P=2)
    mp.spawn(_worker, args=(world, free_port(), d["mcb_path"], admixture, eta_constrained, K,
                            str(tmp_path), slices), nprocs=world, join=True)
    # the unsharded oracle
    I = d["I"]
    per_indiv = bool(admixture and not eta_constrained)
    eta, p = random_params(np.random.default_rng(5), I, K, d["J"], per_indiv)
    fit = orc.Fit(d["J"], d["codes"], admixture=admixture, eta_constrained=eta_constrained)
    fit.alloc(K)
    fit.set_params(0, eta, p)
    fit.set_indices(0, 0, 0)
    ref_ll = []
    for it in range(3):
        ref_ll.append(fit.e_step())
        fit.m_step()
    e_ref, p_ref = fit.get_params(0)
    ranks = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    # every rank ends with bit-identical replicated parameters
    assert np.array_equal(ranks[0]["p"], ranks[1]["p"])
    assert np.array_equal(ranks[0]["ll"], ranks[1]["ll"])
    assert np.allclose(ranks[0]["ll"], ref_ll, rtol=1e-12, atol=0)
    assert np.max(np.abs(ranks[0]["p"] - p_ref)) < 1e-12
    if per_indiv:
        got = np.concatenate([r["eta"] for r in ranks])
        assert np.max(np.abs(got - e_ref)) < 1e-12
    else:
        assert np.array_equal(ranks[0]["eta"], ranks[1]["eta"])
        assert np.max(np.abs(ranks[0]["eta"] - e_ref)) < 1e-12


def test_shard_bounds():
    from multiclust_b200.sharding import shard_bounds
    b = shard_bounds(10, 4)
    assert b[0][0] == 0 and b[-1][1] == 10
    assert all(b[i][1] == b[i + 1][0] for i in range(3))
    assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
