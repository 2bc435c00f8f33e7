"""Bit-exact parsing / allele recoding of the C host (multiclust_b200/host/
read_data.c) against the reference parser: through the golden fixtures (which
hold the reference's own parse of the same generated text) and, where the
prebuilt reference harness is available, live on hand-written edge cases
(interleaved rows, --missing remap, all-missing locus, tetraploid, the
inter-marker distance line, -R).  No GPU is involved: --parse-only stops
before any device call."""
import os
import subprocess

import numpy as np
import pytest

from common import ROOT, ensure_mc_gen, golden_names, load_golden

CLI = os.path.join(ROOT, "multiclust_b200", "host", "multiclust")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(CLI):
        from multiclust_b200 import build
        build.build_host()


def parse_with_cli(path, out, extra=()):
    r = subprocess.run([CLI, "-f", path, "--parse-only", out] + list(extra),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    from oracle import orc
    return orc.read_mcb(out)


def regenerate(tmp_path, gen, interleaved=False):
    stru = str(tmp_path / "d.stru")
    cmd = [ensure_mc_gen(), "--I", str(gen["I"]), "--L", str(gen["L"]), "--K", str(gen["K"]),
           "--jmax", str(gen["jmax"]), "--miss", str(gen["miss"]), "--P", str(gen["P"]),
           "--stru", stru]
    if interleaved:
        cmd.append("--interleaved")
    subprocess.check_call(cmd)
    return stru


@pytest.mark.parametrize("name", ["admix_em", "admix_tetra", "admix_k10",
                                  "admix_biallelic", "admix_sweep", "mix_biallelic_k5"])
@pytest.mark.parametrize("interleaved", [False, True])
def test_parse_matches_reference_golden(tmp_path, name, interleaved):
    g = load_golden(name)
    gen = g["meta"]["gen"]
    stru = regenerate(tmp_path, gen, interleaved)
    d = parse_with_cli(stru, str(tmp_path / "o.mcb"), ["-p", str(gen["P"])])
    for key in ("J", "nreal", "labels", "codes", "locale"):
        assert np.array_equal(d[key], g[key]), key


def write(path, text):
    with open(path, "w") as fp:
        fp.write(text)
    return str(path)


EDGE = {
    # stacked diploid, one all-missing locus, one locus with a missing copy
    "all_missing_locus": ("l1 l2 l3\n"
                          "a p1 3 -9 7\n" "a p1 5 -9 7\n"
                          "b p2 3 -9 -9\n" "b p2 3 -9 9\n"
                          "c p1 5 -9 8\n" "c p1 4 -9 7\n", []),
    # user-defined missing marker
    "missing_remap": ("l1 l2\n"
                      "a p1 1 0\n" "a p1 2 4\n"
                      "b p1 0 4\n" "b p1 2 6\n", ["--missing", "0"]),
    # interleaved tetraploid, every column named
    "tetra_interleaved": ("l1a l1b l1c l1d l2a l2b l2c l2d\n"
                          "a p1 1 2 2 3 9 9 8 -9\n"
                          "b p2 3 3 3 3 8 7 7 7\n"
                          "c p2 1 1 -9 2 9 9 9 9\n", ["-p", "4"]),
    # interleaved diploid, each locus named once
    "interleaved_named_once": ("l1 l2 l3\n"
                               "a p1 11 12 5 5 7 -9\n"
                               "b p2 12 12 6 5 7 7\n"
                               "c p3 13 11 5 6 -9 -9\n", []),
    # inter-marker distance line (the reference then counts one line short)
    "distance_line": ("l1 l2\n" "-1 10 20\n"
                      "a p1 1 2\n" "a p1 2 2\n"
                      "b p1 1 1\n" "b p1 3 2\n"
                      "c p2 1 2\n" "c p2 1 3\n"
                      "d p2 3 3\n" "d p2 1 2\n" "e p2 1 1\n", []),
    # R-formatted header (two extra column names)
    "r_format": ("id pop l1 l2\n"
                 "a p1 1 2\n" "a p1 2 2\n"
                 "b p1 1 1\n" "b p1 3 2\n", ["-R"]),
}


@pytest.mark.parametrize("case", sorted(EDGE))
def test_edge_cases_against_live_reference(tmp_path, case):
    from oracle import orc
    if not orc.have_ref():
        pytest.skip("oracle/_ref/ref_harness not built")
    text, extra = EDGE[case]
    stru = write(tmp_path / "e.stru", text)
    mine = parse_with_cli(stru, str(tmp_path / "mine.mcb"), extra)
    pre = str(tmp_path / "ref")
    r = orc.run_ref(["-f", stru, "-a", "-k", "2"] + extra, dump=pre, parse_only=True)
    assert r.returncode == 0, r.stderr
    ref = orc.read_mcb(pre + ".parse.mcb")
    for key in ("I", "L", "P", "npops"):
        assert mine[key] == ref[key], key
    for key in ("J", "nreal", "labels", "codes", "locale"):
        assert np.array_equal(mine[key], ref[key]), key


def test_distance_line_quirk_same_error(tmp_path):
    """with an even number of haplotype rows after a distance line the
    reference miscounts and refuses the file; so does the drop-in (same code)"""
    text = EDGE["distance_line"][0].replace("e p2 1 1\n", "")
    stru = write(tmp_path / "e.stru", text)
    r = subprocess.run([CLI, "-f", stru, "--parse-only", str(tmp_path / "o.mcb")],
                       capture_output=True, text=True)
    assert r.returncode == 7 and "is not a multiple of ploidy" in r.stderr


def test_unsupported_options_are_refused(tmp_path):
    stru = write(tmp_path / "e.stru", EDGE["r_format"][0])
    for flag in (["-x"], ["--simulate", "q", "p"], ["-I"], ["-u", "1"]):
        r = subprocess.run([CLI, "-f", stru] + flag, capture_output=True, text=True)
        assert r.returncode != 0
        assert "outside the EM path" in r.stderr


def test_bootstrap_option_checks(tmp_path):
    """-b n (reference multiclust.c:869-877, 1427-1434): K must exceed 1; not with --shard-fits"""
    stru = write(tmp_path / "e.stru", EDGE["r_format"][0])
    r = subprocess.run([CLI, "-f", stru, "-R", "-k", "1", "-b", "2"], capture_output=True, text=True)
    assert r.returncode == 11 and "must exceed 1" in r.stderr     # INVALID_USER_SETUP
    r = subprocess.run([CLI, "-f", stru, "-R", "-k", "2", "-b", "2", "--gpus", "2",
                        "--shard-fits"], capture_output=True, text=True)
    assert r.returncode == 11 and "one model at a time" in r.stderr
    r = subprocess.run([CLI, "-f", stru, "-R", "-k", "2", "-b", "-1"], capture_output=True, text=True)
    assert r.returncode == 10                                      # INVALID_CMD_ARGUMENT


def test_missing_file_and_bad_option(tmp_path):
    r = subprocess.run([CLI, "-f", str(tmp_path / "nope.stru"), "--parse-only", "x"],
                       capture_output=True, text=True)
    assert r.returncode == 5 and "could not open file" in r.stderr    # FILE_OPEN_ERROR
    r = subprocess.run([CLI, "-Z"], capture_output=True, text=True)
    assert r.returncode == 9                                          # INVALID_CMD_OPTION
    r = subprocess.run([CLI, "-k", "3"], capture_output=True, text=True)
    assert r.returncode == 8                                          # INVALID_CMDLINE


def test_timed_repetition_option_checks(tmp_path):
    """-w n <r> [t <min>] [m <min>] (reference multiclust.c:1683-1714): the number of
    repetitions must be positive, an unknown sub-option is an invalid option, and the
    bootstrap cannot be timed"""
    stru = write(tmp_path / "e.stru", EDGE["r_format"][0])
    r = subprocess.run([CLI, "-f", stru, "-R", "-k", "2", "-w", "n", "0"], capture_output=True, text=True)
    assert r.returncode == 10                                      # INVALID_CMD_ARGUMENT
    r = subprocess.run([CLI, "-f", stru, "-R", "-k", "2", "-w", "x", "3"], capture_output=True, text=True)
    assert r.returncode == 9                                       # INVALID_CMD_OPTION
    r = subprocess.run([CLI, "-f", stru, "-R", "-k", "2", "-w", "n", "2", "t"], capture_output=True, text=True)
    assert r.returncode == 10
    r = subprocess.run([CLI, "-f", stru, "-R", "-k", "2", "-w", "n", "2", "-b", "2"],
                       capture_output=True, text=True)
    assert r.returncode == 11 and "cannot be timed" in r.stderr   # INVALID_USER_SETUP
