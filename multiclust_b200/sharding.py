"""Individual-sharded EM step: plumbing shared by bench.py (NCCL, one rank per
GPU) and the CPU tests (gloo).

A single large fit shards individuals: rank r owns the rows
[r*I/N, (r+1)*I/N) of the genotype matrix, of eta and of the posterior sums; p
is replicated.  One EM step is

    local   E-step + everything that needs only this rank's individuals
            (mc_em_step_local); leaves [K*T allele-count sums | ll | K pooled
            sums] in the rank's exchange buffer
    gather  ONE all-gather of the exchange buffers
    sum     the N buffers are added in rank order (mc_exchange_sum), so every
            rank holds bit-identical totals whatever algorithm the collective
            library used
    finish  normalise + project p (and pooled eta) into the target slot,
            identically on every rank (mc_em_step_finish)

No compute lives here: `shard` is any object with em_step_local / exchange /
sum_gathered / em_step_finish (multiclust_b200.Context for CUDA, an oracle
adapter in the tests).
"""


def shard_bounds(n_individuals, world):
    """[(first, last)) rows of each rank; sizes differ by at most one"""
    return [(r * n_individuals // world, (r + 1) * n_individuals // world)
            for r in range(world)]


def sharded_em_step(shard, dist, world, frm, to, gathered):
    """one EM step of an individual-sharded fit; returns the global log
    likelihood.  `gathered` is a preallocated tensor of world * len(exchange)."""
    shard.em_step_local(frm, to)
    if world == 1:
        return shard.em_step_finish(to)
    dist.all_gather_into_tensor(gathered, shard.exchange_tensor())
    shard.sum_gathered(gathered, world)
    return shard.em_step_finish(to)
