#!/usr/bin/env python
"""Python restatement of how k3_build_csc (csrc/mc_admix3_build.cuh) schedules the pass-2 entry
lists of the gather kernel, on tiles of the bench's own generator (include/mc_synth.h), with the
cost model the kernel is built around: the eta rows of individuals with different i % 8 start in
different shared-memory bank groups, so one LDS.128 of a quarter warp (8 lanes) costs as many
wavefronts as the most frequent residue class among its 8 entries.

    python tools/list_schedule_sim.py [n_tiles]

prints wavefronts per quarter-warp step for
  fixed    entries dealt to fixed (lane + step) % 8 slots, contiguous lanes per column
           (the first round-2 builder)
  sched    lanes of a column spread over the quarter warps, carriers dealt round the lanes
           class by class, per step the lanes take the best-stocked free class with one
           augmenting move (the builder as it is)
and for an optimal schedule of the same dealt lists (bipartite edge colouring), which reaches
the bound max(longest lane, largest class total) steps per quarter warp without a conflict.
tests/test_list_schedule.py checks the restatement's invariants and the ordering of the three numbers on one tile; the
device-side counts are in profiles/r02_ncu_admix3_kernel.txt (LDS.128 wavefronts against ideal).
"""
import sys
from collections import defaultdict

M64 = (1 << 64) - 1
IT, LT, NQ = 256, 8, 32         # individuals and loci per tile (diploid), quarter warps


# ---- include/mc_synth.h ------------------------------------------------------------------
def _mix(x):
    x = (x + 0x9e3779b97f4a7c15) & M64
    x = ((x ^ (x >> 30)) * 0xbf58476d1ce4e5b9) & M64
    x = ((x ^ (x >> 27)) * 0x94d049bb133111eb) & M64
    return x ^ (x >> 31)


def _hash(seed, tag, a, b):
    h = _mix(seed ^ ((tag * 0xd6e8feb86659fd93) & M64))
    h = _mix(h ^ a)
    return _mix(h ^ ((b + 0x632be59bd9b4e019) & M64))


class Synth:
    def __init__(self, seed=20261018, K=10, jmax=20, miss_bp=500):
        self.seed, self.K, self.jmax, self.miss = seed, K, jmax, miss_bp

    def nalleles(self, l):
        return 2 if self.jmax <= 2 else 2 + _hash(self.seed, 1, l, 0) % (self.jmax - 1)

    def code(self, i, l, a):
        cell = (i * 0x100000001b3 + l) & M64
        if self.miss and _hash(self.seed, 4, cell, a) % 10000 < self.miss:
            return 255
        w = [(1 + (_hash(self.seed, 3, i, k) >> 48)) ** 3 for k in range(self.K)]
        r, acc, kk = _hash(self.seed, 5, cell, a) % sum(w), 0, 0
        for k in range(self.K):
            acc += w[k]
            if r < acc:
                kk = k
                break
        n = self.nalleles(l)
        w = [(1 + (_hash(self.seed, 2, l * 4096 + kk, j) >> 44)) ** 2 for j in range(n)]
        r, acc = _hash(self.seed, 6, cell, a) % sum(w), 0
        for j in range(n):
            acc += w[j]
            if r < acc:
                return j
        return 0


def tile_columns(g, it, lt):
    """the allele columns of tile (it, lt): per column the individuals (0..255) carrying it"""
    cols = []
    for l in range(lt * LT, lt * LT + LT):
        codes = [(g.code(i, l, 0), g.code(i, l, 1)) for i in range(it * IT, it * IT + IT)]
        for j in range(g.nalleles(l)):
            carriers = [ii for ii, c in enumerate(codes) if j in c]
            if carriers:
                cols.append(carriers)
    return cols


# ---- the builder -------------------------------------------------------------------------
def lanes_per_column(cols):
    """smallest list length q with sum ceil(n_c / q) <= 256; column c gets ceil(n_c / q) lanes"""
    cnt = [len(c) for c in cols]
    q = max(1, (sum(cnt) + IT - 1) // IT)
    while sum((n + q - 1) // q for n in cnt) > IT:
        q += 1
    return [(n + q - 1) // q for n in cnt]


def deal(cols):
    """class-major cyclic dealing: lane -> {class: [individuals]} (logical lanes, contiguous per
    column)"""
    S = lanes_per_column(cols)
    sub, lane0 = [defaultdict(list) for _ in range(sum(S))], 0
    for c, carriers in enumerate(cols):
        for k, ii in enumerate(sorted(carriers, key=lambda ii: (ii & 7, ii))):
            sub[lane0 + k % S[c]][ii & 7].append(ii)
        lane0 += S[c]
    return sub


def schedule(sub):
    """per quarter warp (logical lanes qw, 32 + qw, ...) and step: the active lanes, fewest
    classes left first, take the best-stocked class nobody took; else one augmenting move; else
    the best-stocked class (a conflict).  Returns {(quarter, slot, step): individual}"""
    out = {}
    for qw in range(NQ):
        lanes = [ln for ln in range(len(sub)) if ln % NQ == qw]
        left = {t: sum(len(v) for v in sub[t].values()) for t in lanes}
        st = 0
        while any(left.values()):
            act = [t for t in lanes if left[t]]
            cnt = {t: [len(sub[t][r]) for r in range(8)] for t in act}
            avail = {t: {r for r in range(8) if cnt[t][r]} for t in act}
            used, holder, choice = set(), {}, {}
            for t in sorted(act, key=lambda t: (len(avail[t]), lanes.index(t))):
                free = avail[t] - used
                if free:
                    r = max(free, key=lambda r: (cnt[t][r], -r))
                    holder[r] = t
                    used.add(r)
                    choice[t] = r
                    continue
                for r in sorted(avail[t], key=lambda r: (-cnt[t][r], r)):
                    h = holder.get(r)
                    alt = avail[h] - used if h is not None else None
                    if alt:
                        r2 = max(alt, key=lambda x: (cnt[h][x], -x))
                        holder[r2], choice[h] = h, r2
                        used.add(r2)
                        holder[r], choice[t] = t, r
                        break
                else:
                    choice[t] = max(avail[t], key=lambda r: (cnt[t][r], -r))
            for t in act:
                out[(qw, lanes.index(t), st)] = sub[t][choice[t]].pop(0)
                left[t] -= 1
            st += 1
    return out


def schedule_optimal(sub):
    """an optimal schedule of the same dealt lists: the entries of a quarter warp are the edges
    of a bipartite multigraph lanes x residue classes; by Koenig's theorem they can be coloured
    with D = max degree colours (alternating-path recolouring), i.e. scheduled into
    D = max(longest lane, largest class total) steps without any conflict -- at the price of
    idle slots inside a lane's list.  Returns {(quarter, slot, step): individual}"""
    out = {}
    for qw in range(NQ):
        lanes = [ln for ln in range(len(sub)) if ln % NQ == qw]
        if not lanes:
            continue
        at_lane = [dict() for _ in lanes]       # slot -> {colour: class}
        at_class = [dict() for _ in range(8)]   # class -> {colour: slot}
        D = max(max(sum(len(v) for v in sub[t].values()) for t in lanes),
                max(sum(len(sub[t][r]) for t in lanes) for r in range(8)))
        for b, t in enumerate(lanes):
            for r in range(8):
                for _ in sub[t][r]:
                    fa = next(c for c in range(D) if c not in at_lane[b])
                    fb = next(c for c in range(D) if c not in at_class[r])
                    if fa in at_class[r]:
                        # free colour fa at class r: swap fa / fb along the path that
                        # starts at class r with its fa-edge (it cannot reach lane b)
                        path, node, colour, is_class = [], r, fa, True
                        while True:
                            tab = at_class[node] if is_class else at_lane[node]
                            if colour not in tab:
                                break
                            nxt = tab[colour]
                            path.append((node, nxt, colour) if is_class else (nxt, node, colour))
                            node, is_class = nxt, not is_class
                            colour = fb if colour == fa else fa
                        for cls, slot, colour in path:
                            del at_class[cls][colour]
                            del at_lane[slot][colour]
                        for cls, slot, colour in path:
                            other = fb if colour == fa else fa
                            at_class[cls][other] = slot
                            at_lane[slot][other] = cls
                    at_lane[b][fa] = r
                    at_class[r][fa] = b
        pools = {t: {r: list(v) for r, v in sub[t].items()} for t in lanes}
        for b, t in enumerate(lanes):
            for colour, r in sorted(at_lane[b].items()):
                out[(qw, b, colour)] = pools[t][r].pop(0)
    return out


def fixed_slots(cols):
    """the first round-2 builder: contiguous lanes per column in one quarter warp after the
    other, entry dealt to a slot with (lane + step) % 8 == i % 8 where one is free"""
    S, out, lane0 = lanes_per_column(cols), {}, 0
    for c, carriers in enumerate(cols):
        n, s_ = len(carriers), S[c]
        q, rem = n // s_, n % s_
        slots, rest = [None] * n, []
        nxt = {r: [0, (r - lane0) & 7] for r in range(8)}
        for ii in carriers:
            r = ii & 7
            st, seg = nxt[r]
            while st <= q and (seg >= s_ or (st == q and seg >= rem)):
                st += 1
                seg = (r - lane0 - st) & 7
            if st < q or (st == q and seg < rem):
                slots[st * s_ + seg] = ii
                nxt[r] = [st, seg + 8]
            else:
                nxt[r] = [q + 1, 0]
                rest.append(ii)
        for ii in rest:
            slots[slots.index(None)] = ii
        for x, ii in enumerate(slots):
            lane = lane0 + x % s_
            out[(lane // 8, lane % 8, x // s_)] = ii
        lane0 += s_
    return out


def wavefronts(out):
    """(wavefronts, quarter-warp steps) of one LDS.128 per entry"""
    steps = defaultdict(list)
    for (qw, _, st), ii in out.items():
        steps[(qw, st)].append(ii & 7)
    return sum(max(v.count(r) for r in range(8)) for v in steps.values()), len(steps)


def bound(sub):
    b = 0
    for qw in range(NQ):
        lanes = [ln for ln in range(len(sub)) if ln % NQ == qw]
        if lanes:
            b += max(max(sum(len(v) for v in sub[t].values()) for t in lanes),
                     max(sum(len(sub[t][r]) for t in lanes) for r in range(8)))
    return b


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    g = Synth()
    tot = defaultdict(int)
    for k in range(n):
        cols = tile_columns(g, 3 + 37 * k, 5 + 211 * k)
        w, s = wavefronts(fixed_slots(cols))
        tot["fixed_w"] += w
        tot["fixed_s"] += s
        sub = deal(cols)
        tot["bound"] += bound(sub)
        w, s = wavefronts(schedule_optimal(sub))
        tot["opt_w"] += w
        tot["opt_s"] += s
        w, s = wavefronts(schedule(sub))         # consumes the dealt lists
        tot["sched_w"] += w
        tot["sched_s"] += s
        tot["entries"] += sum(len(c) for c in cols)
    print("%d tiles, %d entries" % (n, tot["entries"]))
    print("fixed  %.3f wavefronts per quarter-warp step (%d steps)" % (
        tot["fixed_w"] / tot["fixed_s"], tot["fixed_s"]))
    print("sched  %.3f wavefronts per quarter-warp step (%d steps)" % (
        tot["sched_w"] / tot["sched_s"], tot["sched_s"]))
    print("optimal (edge colouring, idle slots inside the lists): %d wavefronts in %d steps = the "
          "bound %d; sched needs %d wavefronts" % (tot["opt_w"], tot["opt_s"], tot["bound"],
                                                   tot["sched_w"]))


if __name__ == "__main__":
    main()
