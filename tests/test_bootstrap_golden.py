"""The bootstrap sampler restated in Python (tests/common.py: the glibc rand() stream and the
loops of bootstrap.c:77-175 in the default parse mode) against samples made by the reference's
OWN parametric_bootstrap() from dumped parameters and a re-seeded generator
(tests/golden/bootsample_*.npz, oracle/ref_harness.c --bootstrap-sample).  This pins the
restatement the GPU test of mc_bootstrap_data is checked with; no GPU needed."""
import glob
import os

import numpy as np
import pytest

from common import ROOT, bootstrap_numpy, codes_to_counts, glibc_stream

NAMES = sorted(os.path.basename(f)[len("bootsample_"):-len(".npz")]
               for f in glob.glob(os.path.join(ROOT, "tests", "golden", "bootsample_*.npz")))


def test_golden_samples_exist():
    assert {"admix", "admix_pooled", "mix", "admix_tetra"} <= set(NAMES)


@pytest.mark.parametrize("name", NAMES)
def test_restated_sampler_equals_reference_sample(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "bootsample_%s.npz" % name))
    J, codes, K = g["J"], g["codes"], int(g["K"])
    I, L, P = codes.shape
    admixture, per_indiv = int(g["admixture"]), int(g["per_indiv"])
    off = np.concatenate([[0], np.cumsum(J)]).astype(np.int64)
    n = I * (2 * L * P if admixture else 1 + L * P)
    x = glibc_stream(int(g["seed"]), n)
    draws = (x[31:] >> np.uint64(1)).astype(np.float64)
    sample = bootstrap_numpy(draws, I, L, P, K, J, off, g["eta"], g["p"], admixture, per_indiv)
    counts = g["counts"].astype(np.int64)
    assert np.array_equal(codes_to_counts(sample, J), counts)
    # the reference's default parse mode draws every copy: missing data are filled in
    live = J > 0
    per_locus = np.add.reduceat(counts, off[:-1][live], axis=1) if live.all() else None
    if per_locus is not None:
        assert np.all(per_locus == P)
    assert (codes == 255).any()                 # although the observed data had missing copies
