"""Individual-sharded EM step: plumbing shared by bench.py (NCCL, one rank per
GPU) and the CPU tests (gloo).

A single large fit shards individuals: rank r owns the rows
[r*I/N, (r+1)*I/N) of the genotype matrix, of eta and of the posterior sums; p
is replicated.  One EM step is

    local   E-step + everything that needs only this rank's individuals
            (mc_em_step_local); leaves [K*T allele-count sums | ll | K pooled
            sums] in the rank's exchange buffer.  The eta side of an admixture
            step keeps running on a second stream during the exchange.
    sum     a deterministic reduce-scatter + all-gather: the buffer is cut into
            N equal slices, slice j goes to rank j (all-to-all), rank r adds
            the N copies of slice r in rank order (mc_exchange_sum_slice) and
            the totals are all-gathered back in place -- 2 * (N-1)/N buffers
            over the links per rank whatever N is, and every rank ends up with
            bit-identical totals whatever algorithm the collective library
            used.  (Shards without sum_slices, or N > 64, fall back to one
            all-gather of the whole buffers + mc_exchange_sum.)
    finish  normalise + project p (and pooled eta) into the target slot,
            identically on every rank (mc_em_step_finish)

No compute lives here: `shard` is any object with em_step_local /
exchange_tensor / sum_slices (or sum_gathered) / em_step_finish
(multiclust_b200.Context for CUDA, an oracle adapter in the tests).
"""


def shard_bounds(n_individuals, world):
    """[(first, last)) rows of each rank; sizes differ by at most one"""
    return [(r * n_individuals // world, (r + 1) * n_individuals // world)
            for r in range(world)]


def _all_to_all(dist, recv, send, m, world, rank):
    """equal-split all-to-all of m elements per peer (gloo has no all_to_all:
    one scatter per source rank there)"""
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(recv, send)
        return
    for src in range(world):
        pieces = [send[j * m:(j + 1) * m].clone() for j in range(world)] if rank == src else None
        dist.scatter(recv[src * m:(src + 1) * m], pieces, src=src)


def exchange_sum(shard, dist, world, scratch):
    """sum the shards' exchange buffers over ranks, deterministically"""
    x = shard.exchange_tensor()             # capacity: logical length + 64
    n = shard.exchange_len()
    if hasattr(shard, "sum_slices") and world <= 64 and scratch.numel() >= n + 64:
        rank = dist.get_rank()
        m = -(-n // world)
        xin, recv = x[:m * world], scratch[:m * world]
        _all_to_all(dist, recv, xin, m, world, rank)
        shard.sum_slices(recv, world, rank * m, m)
        own = recv[:m]                      # the parts are summed: reuse their space
        own.copy_(xin[rank * m:(rank + 1) * m])
        dist.all_gather_into_tensor(xin, own)
    else:
        gathered = scratch[:world * n]
        dist.all_gather_into_tensor(gathered, x[:n])
        shard.sum_gathered(gathered, world)


def sharded_em_step(shard, dist, world, frm, to, scratch):
    """one EM step of an individual-sharded fit; returns the global log
    likelihood.  `scratch` is a preallocated tensor of at least
    world * len(exchange) elements (len + 64 suffice for the slice path)."""
    shard.em_step_local(frm, to)
    if world > 1:
        exchange_sum(shard, dist, world, scratch)
    return shard.em_step_finish(to)
