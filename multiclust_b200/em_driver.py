"""Python mirror of host/em_driver.c for one-process-per-GPU runs (torchrun):
bench.py's parity check and other-configuration timings, and the tests.

No compute lives here.  The control flow is the reference's em() /
em_step() / em_2_steps() / accelerated_em_step() (em_alg.c:44-207, 1072-1171;
accel_em.c:35-419) with every pass over data or parameters delegated to a
`shard` -- multiclust_b200.Context (CUDA) or the oracle adapter of the tests.
With world > 1 the individuals are sharded over the ranks: the exchange step
of sharding.py sums the sufficient statistics, the eta parts of the
acceleration dot products and the log likelihoods are added over ranks in rank
order (all_gather + ordered sum), the p parts are replicated.
"""
import math

import numpy as np

from .sharding import sharded_em_step

QN = 4      # accel_scheme: 0 EM, 1-3 SQUAREM, 4 QN q=1, 5-6 QN q=2,3 (multiclust.c:818-820)


def accel_q(accel):
    return 0 if accel == 0 else (accel - 3 if accel > QN else 1)


class Driver:
    def __init__(self, shard, dist=None, world=1, admixture=1, eta_constrained=0, accel=0,
                 max_iter=0, abs_error=1e-4, rel_error=0.0, n_init_iter=0, adjust_step=0,
                 gathered=None):
        self.s, self.dist, self.world = shard, dist, world
        self.accel, self.q = accel, max(accel_q(accel), 1)
        self.eta_sharded = bool(admixture and not eta_constrained)
        self.max_iter, self.abs_error, self.rel_error = max_iter, abs_error, rel_error
        self.n_init_iter, self.adjust_step = n_init_iter, adjust_step
        self.gathered = gathered
        self.pindex = self.findex = self.tindex = 0
        self.delta_index = 0
        self.n_iter = 0
        self.logL = -math.inf
        self.converged = self.stopped = self.iter_stop = self.accel_step = 0
        self.trace = []

    # ---- sums over ranks, in rank order --------------------------------
    def _rank_sum(self, vals):
        """vals: list of floats local to this rank -> their sums over ranks"""
        if self.world == 1:
            return list(vals)
        import torch
        t = torch.tensor(vals, dtype=torch.float64, device=self._device())
        out = torch.empty(self.world * len(vals), dtype=torch.float64, device=t.device)
        self.dist.all_gather_into_tensor(out, t)
        out = out.cpu().numpy().reshape(self.world, len(vals))
        tot = out[0].copy()
        for r in range(1, self.world):
            tot += out[r]
        return [float(x) for x in tot]

    def _device(self):
        return "cuda" if self.dist.get_backend() == "nccl" else "cpu"

    # ---- em_alg.c:101-182 ------------------------------------------------
    def stop(self, ll):
        self.n_iter += 1
        self.trace.append(ll)
        if math.isnan(ll):
            raise FloatingPointError("nan log likelihood")
        stopped = 0
        if self.max_iter and self.n_iter > self.max_iter:
            self.iter_stop = stopped = 1
        else:
            abs_diff = abs(ll - self.logL) if self.abs_error else 0.0
            rel_diff = abs_diff / abs(self.logL) if self.rel_error else 0.0
            done = not ((self.abs_error and abs_diff > self.abs_error)
                        or (self.rel_error and rel_diff > self.rel_error))
            if done:
                self.converged = stopped = 1
        self.stopped = stopped
        if ll < self.logL and not stopped:
            raise ArithmeticError("log likelihood decrease (%r < %r)" % (ll, self.logL))
        self.accel_step = 0
        self.logL = ll
        return stopped

    def em_step(self):
        ll = sharded_em_step(self.s, self.dist, self.world, self.findex, self.tindex,
                             self.gathered)
        return self.stop(ll)

    def log_likelihood(self, which):
        return self._rank_sum([self.s.loglik(which)])[0]

    # ---- em_alg.c:1072-1171 --------------------------------------------
    def em_2_steps(self):
        self.findex = self.pindex
        self.tindex = (self.findex + 1) % 3
        for j in range(2):
            if self.em_step():
                return 1
            self.s.delta(j, self.delta_index, self.tindex, self.findex)
            self.findex = self.tindex
            self.tindex = (self.findex + 1) % 3
            if self.tindex == self.pindex:
                self.tindex = (self.tindex + 1) % 3
        self.delta_index = (self.delta_index + 1) % self.q
        return 0

    def _dots(self, e, p):
        """eta parts are sharded (added over ranks) unless eta is pooled"""
        e = self._rank_sum(list(e)) if self.eta_sharded else list(e)
        return [a + b for a, b in zip(e, p)]

    # ---- accel_em.c:130-243 ----------------------------------------------
    def step_size(self):
        e, p = self.s.step_dots(self.delta_index)
        utu, utvu, vutvu = self._dots(e, p)
        try:
            if self.accel == 1:
                s = utu / utvu
            elif self.accel == 2:
                s = utvu / vutvu
            elif self.accel == 3:
                if math.sqrt(utu) < 1e-8:
                    return math.nan
                s = -math.sqrt(utu / vutvu)
            else:
                s = -utu / utvu
        except ZeroDivisionError:
            return math.nan
        if self.accel < QN and s > -1:
            s = -1.0
        return s

    def accelerated_update(self, s):
        self.delta_index = self.delta_index - 1 if self.delta_index else self.q - 1
        self.s.accel_update(self.accel == QN, self.tindex, self.pindex, self.delta_index, s)
        ll = self.log_likelihood(self.tindex)
        self.delta_index = (self.delta_index + 1) % self.q
        return ll

    # ---- accel_em.c:262-419 ----------------------------------------------
    def qn_accelerated_update(self):
        q = self.q
        vindex = self.delta_index - 1 if self.delta_index else q - 1
        uindex = vindex - 1 if vindex else q - 1
        A = np.zeros(q * q)
        cutu = np.zeros(q)
        q1, j = self.delta_index, 0
        while True:
            q2, n = self.delta_index, 0
            while True:
                e, p = self.s.qn_dots(q1, q2)
                utu, utv = self._dots(e, p)
                cutu[n] = utu
                A[j * q + n] = utu - utv
                n += 1
                q2 = (q2 + 1) % q
                if q2 == self.delta_index:
                    break
            q1 = (q1 + 1) % q
            j += 1
            if q1 == self.delta_index:
                break
        Ai = np.zeros(q * q)
        if q == 1:
            Ai[0] = 1 / A[0]
        elif q == 2:
            det = A[0] * A[3] - A[1] * A[2]
            Ai[0], Ai[3], Ai[1], Ai[2] = A[3] / det, A[0] / det, -A[1] / det, -A[2] / det
        else:
            c00 = A[4] * A[8] - A[5] * A[7]
            c01 = A[8] * A[3] - A[5] * A[6]
            c02 = A[3] * A[7] - A[4] * A[6]
            det = A[0] * c00 - A[1] * c01 + A[2] * c02
            Ai[:] = [c00 / det, (A[2] * A[7] - A[1] * A[8]) / det, (A[1] * A[5] - A[2] * A[4]) / det,
                     (A[5] * A[6] - A[3] * A[8]) / det, (A[0] * A[8] - A[2] * A[6]) / det,
                     (A[2] * A[3] - A[0] * A[5]) / det,
                     c02 / det, (A[1] * A[6] - A[0] * A[7]) / det, (A[0] * A[4] - A[1] * A[3]) / det]
        self.s.qn_update(self.tindex, self.pindex, uindex, self.delta_index, Ai, cutu)
        return self.log_likelihood(self.tindex)

    # ---- accel_em.c:35-114 -----------------------------------------------
    def accelerated_em_step(self):
        self.em_2_steps()
        if self.stopped:
            return 1
        emll = self.log_likelihood(self.findex)
        s, keep_em = 0.0, False
        if self.accel <= QN:
            s = self.step_size()
            keep_em = math.isnan(s) or math.isinf(s)
        if not keep_em:
            n_adjust = 0
            while True:
                ll = self.accelerated_update(s) if self.accel <= QN else self.qn_accelerated_update()
                if self.adjust_step and ll < emll:
                    s = (s - 1) / 2
                go = n_adjust < self.adjust_step and ll < emll and s < -1
                n_adjust += 1
                if not go:
                    break
            if ll > emll:
                self.pindex = self.tindex
                self.accel_step = 1
                return 0
        self.pindex = self.findex
        return 0

    # ---- em_alg.c:44-90 -----------------------------------------------------
    def em(self, K):
        halt = 0
        if K == 1:
            self.em_step()
            self.logL = self.log_likelihood(self.tindex)
            return
        while self.n_iter < self.n_init_iter and not halt:
            halt = self.em_step()
        for _ in range(1, self.q):
            self.em_2_steps()
            self.pindex = self.findex
        if self.converged:
            return
        while True:     # do ... while (!halt)
            halt = self.accelerated_em_step() if self.accel else self.em_step()
            if halt:
                break
