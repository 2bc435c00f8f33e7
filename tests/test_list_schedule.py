"""The list scheduling of the gather kernel's pass 2 (k3_build_csc, csrc/mc_admix3_build.cuh)
restated in Python (tools/list_schedule_sim.py), no GPU: the generator port produces the bytes
of include/mc_synth.h, the schedule places every entry exactly once in its own lane, and it
needs fewer shared-memory wavefronts than the fixed-slot dealing it replaced -- the ordering
DESIGN.md section 4.1 quotes.  (The kernel itself is checked against the oracle by the GPU
tests; the device-side wavefront counts are in profiles/r02_ncu_admix3_kernel.txt.)"""
import os
import sys

import numpy as np

from common import ROOT, gen_data

sys.path.insert(0, os.path.join(ROOT, "tools"))
import list_schedule_sim as sim  # noqa: E402


def test_generator_port_matches_mc_synth(tmp_path):
    I, L = 48, 12
    d = gen_data(tmp_path, I, L, K=10, jmax=20, miss=500, P=2, seed=20261018)
    g = sim.Synth(seed=20261018, K=10, jmax=20, miss_bp=500)
    lab_off = np.concatenate([[0], np.cumsum(d["nreal"])])
    for i in range(I):
        for l in range(L):
            for a in range(2):
                c = int(d["codes"][i, l, a])
                want = g.code(i, l, a)
                if c == 255:
                    assert want == 255
                else:       # mc_gen labels allele j of locus l 101 + 3 j + l % 4; codes are ranks
                    assert int(d["labels"][lab_off[l] + c]) == 101 + 3 * want + l % 4


def test_schedule_places_every_entry_once_and_saves_wavefronts():
    g = sim.Synth()
    cols = sim.tile_columns(g, 3, 5)
    S = sim.lanes_per_column(cols)
    assert sum(S) <= sim.IT
    sub = sim.deal(cols)
    lane_of = {}
    lane0 = 0
    for c, carriers in enumerate(cols):
        sizes = [sum(len(v) for v in sub[lane0 + s].values()) for s in range(S[c])]
        assert sum(sizes) == len(carriers) and max(sizes) - min(sizes) <= 1
        for s in range(S[c]):
            for r, v in sub[lane0 + s].items():
                assert all(ii & 7 == r for ii in v)
                for ii in v:
                    lane_of[(c, ii)] = lane0 + s
        lane0 += S[c]
    assert len(lane_of) == sum(len(c) for c in cols)
    best = sim.bound(sub)
    per_lane = {}
    for ln, classes in enumerate(sub):
        per_lane[ln] = sorted(ii for v in classes.values() for ii in v)
    # an optimal schedule of the same lists (edge colouring): conflict-free in `best` steps
    opt = sim.schedule_optimal(sub)
    w_opt, s_opt = sim.wavefronts(opt)
    assert w_opt == s_opt == best
    got = {}
    for (qw, slot, st), ii in opt.items():
        got.setdefault(slot * sim.NQ + qw, []).append(ii)
    assert all(sorted(got.get(ln, [])) == want for ln, want in per_lane.items())
    out = sim.schedule(sub)
    got = {}
    for (qw, slot, st), ii in out.items():
        got.setdefault(slot * sim.NQ + qw, []).append((st, ii))
    for ln, want in per_lane.items():
        steps = sorted(got.get(ln, []))
        assert [st for st, _ in steps] == list(range(len(want)))    # dense, no holes
        assert sorted(ii for _, ii in steps) == want
    w_new, s_new = sim.wavefronts(out)
    w_old, s_old = sim.wavefronts(sim.fixed_slots(cols))
    assert best <= w_new < w_old
    assert w_new / s_new < 1.3 < w_old / s_old
