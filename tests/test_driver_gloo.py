"""multiclust_b200/em_driver.py (the torchrun mirror of host/em_driver.c) on CPU: gloo ranks,
each holding a slice of the individuals of a golden case, drive the oracle through the same
em() / em_2_steps() / accelerated_em_step() control flow -- SQUAREM, QN q=1 and q=2, plain EM,
mixture -- and must reproduce the log-likelihood trajectory the UNMODIFIED reference produced
(tests/golden).  The eta parts of the acceleration dot products and the log likelihoods are
summed over ranks, the p parts are replicated: exactly what bench.py's parity check does on the
GPUs."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleAccelShard:
    """the shard interface of em_driver.Driver on top of the CPU oracle; the vector
    operations of the acceleration schemes are numpy restatements of accel_em.c"""

    def __init__(self, orc, J, codes, o, K, bound, eta, p, n_total):
        self.orc = orc
        self.fit = orc.Fit(J, codes, admixture=o["admixture"], eta_constrained=o["eta_constrained"],
                           do_projection=o["do_projection"], lower_bound=bound)
        self.fit.alloc(K)
        self.K, self.T, self.J = K, self.fit.T, np.asarray(J)
        self.I = codes.shape[0]
        self.per_indiv = bool(o["admixture"] and not o["eta_constrained"])
        self.proj, self.lb = o["do_projection"], bound
        self.x = [(eta.copy(), p.copy()) for _ in range(3)]
        self.u, self.v = {}, {}
        self.n = K * self.T + 1 + K
        self.buf = torch.zeros(self.n + 64, dtype=torch.float64)

    # -- sharded EM step (as in test_sharded_gloo.OracleShard)
    def em_step_local(self, frm, to):
        self.fit.set_params(0, *self.x[frm])
        self.fit.set_indices(0, 0, 0)
        ll = self.fit.e_step()
        N, S = self.fit.sums()
        kt = self.K * self.T
        self.buf[:kt] = torch.from_numpy(N)
        self.buf[kt] = ll
        self.buf[kt + 1:self.n] = torch.from_numpy(S)
        self._to = to

    def exchange_tensor(self):
        return self.buf

    def exchange_len(self):
        return self.n

    def sum_slices(self, parts, world, first, count):
        total = parts[:count].clone()
        for r in range(1, world):
            total += parts[r * count:(r + 1) * count]
        self.buf[first:first + count] = total

    def em_step_finish(self, to):
        kt = self.K * self.T
        self.fit.m_step_from_sums(self.buf[:kt].numpy(), self.buf[kt + 1:self.n].numpy())
        self.x[to] = self.fit.get_params(0)
        return float(self.buf[kt])

    def loglik(self, slot):
        self.fit.set_params(1, *self.x[slot])
        return self.fit.log_likelihood(1)

    # -- acceleration plumbing (accel_em.c:142-184, 291-310, 364-402, 449-503)
    def delta(self, which, pair, t, f):
        d = (self.x[t][0] - self.x[f][0], self.x[t][1] - self.x[f][1])
        (self.v if which else self.u)[pair] = d

    def step_dots(self, pair):
        out = []
        for part in (0, 1):
            u, v = self.u[pair][part], self.v[pair][part]
            r = v - u
            out.append(np.array([u @ u, u @ r, r @ r]))
        return out[0], out[1]

    def qn_dots(self, q1, q2):
        out = []
        for part in (0, 1):
            out.append(np.array([self.u[q1][part] @ self.u[q2][part],
                                 self.u[q1][part] @ self.v[q2][part]]))
        return out[0], out[1]

    def _project(self, eta, p):
        if not self.proj:
            return eta, p
        K, T = self.K, self.T
        if self.per_indiv:
            eta = np.concatenate([self.orc.project(eta[i * K:(i + 1) * K], self.lb)
                                  for i in range(self.I)]) if self.I else eta
        else:
            eta = self.orc.project(eta, self.lb)
        p = p.copy()
        off = np.concatenate([[0], np.cumsum(self.J)])
        for k in range(K):
            for l in range(len(self.J)):
                a, b = k * T + off[l], k * T + off[l + 1]
                if b > a:
                    p[a:b] = self.orc.project(p[a:b], self.lb)
        return eta, p

    def accel_update(self, qn1, t, p_, pair, s):
        out = []
        for part in (0, 1):
            x, u, v = self.x[p_][part], self.u[pair][part], self.v[pair][part]
            out.append(x + u + s * v if qn1 else x - 2 * s * u + s * s * (v - u))
        self.x[t] = self._project(*out)

    def qn_update(self, t, p_, uindex, delta_index, Ainv, cutu):
        q = len(cutu)
        out = []
        for part in (0, 1):
            acc = self.x[p_][part] + self.u[uindex][part]
            for j in range(q):
                v = self.v[(delta_index + j) % q][part]
                for n in range(q):
                    acc = acc + v * Ainv[j * q + n] * cutu[n]
            out.append(acc)
        self.x[t] = self._project(*out)


def _worker(rank, world, port, name, out):
    sys.path.insert(0, ROOT)
    from oracle import orc
    from multiclust_b200.em_driver import Driver, accel_q
    from multiclust_b200.sharding import shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    meta = json.loads(str(z["meta"]))
    o, fit = meta["options"], meta["fits"][0]
    K, key = fit["K"], "K%d_i%d_" % (fit["K"], fit["init"])
    codes, J = z["codes"], z["J"]
    I = codes.shape[0]
    per_indiv = bool(o["admixture"] and not o["eta_constrained"])
    lo, hi = shard_bounds(I, world)[rank]
    eta0 = z[key + "start_eta"]
    eta_l = eta0.reshape(I, K)[lo:hi].ravel() if per_indiv else eta0
    shard = OracleAccelShard(orc, J, codes[lo:hi], o, K, meta["bound"], eta_l, z[key + "start_p"], I)
    drv = Driver(shard, dist, world, admixture=o["admixture"], eta_constrained=o["eta_constrained"],
                 accel=o["accel"], max_iter=o["max_iter"], abs_error=o["abs_error"],
                 rel_error=o["rel_error"],
                 gathered=torch.zeros(world * (shard.n + 64), dtype=torch.float64))
    assert drv.q == max(accel_q(o["accel"]), 1)
    drv.em(K)
    eta, p = shard.x[drv.pindex]
    np.savez(os.path.join(out, "%s_rank%d.npz" % (name, rank)), ll=np.array(drv.trace), eta=eta,
             p=p, lo=lo, hi=hi, n_iter=drv.n_iter)
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("name", ["admix_em", "admix_s1", "admix_s3", "admix_s4", "admix_s5",
                                  "mix_s1"])
def test_sharded_driver_reproduces_reference_trajectory(orc, tmp_path, name):
    world = 2
    mp.spawn(_worker, args=(world, free_port(), name, str(tmp_path)), nprocs=world, join=True)
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    meta = json.loads(str(z["meta"]))
    fit = meta["fits"][0]
    K, key = fit["K"], "K%d_i%d_" % (fit["K"], fit["init"])
    ranks = [np.load(os.path.join(str(tmp_path), "%s_rank%d.npz" % (name, r))) for r in range(world)]
    ref_ll = z[key + "ll"]
    for r in ranks:
        assert r["ll"].shape == ref_ll.shape
        assert np.max(np.abs(r["ll"] - ref_ll) / np.abs(ref_ll)) <= 1e-9
        assert int(r["n_iter"]) == fit["n_iter"]
    assert np.array_equal(ranks[0]["ll"], ranks[1]["ll"])
    assert np.array_equal(ranks[0]["p"], ranks[1]["p"])
    assert np.max(np.abs(ranks[0]["p"] - z[key + "final_p"])) <= 1e-7
    if meta["options"]["admixture"] and not meta["options"]["eta_constrained"]:
        eta = np.concatenate([r["eta"] for r in ranks])
    else:
        eta = ranks[0]["eta"]
    assert np.max(np.abs(eta - z[key + "final_eta"])) <= 1e-7
