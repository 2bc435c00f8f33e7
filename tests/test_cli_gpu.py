"""End-to-end parity of the product binary (multiclust_b200/host/multiclust:
C host + libmc_cuda.so) with the golden vectors written by the unmodified
reference: same generated STRUCTURE text, same command line (plus --trace /
--dump for full precision), compared per fit:
  - initial parameters after initialize_model (same rand() stream)   1e-12
  - every log likelihood handed to stop()                  1e-9 relative
  - final parameters and posterior sums                    1e-7 absolute
  - n_iter, converged, iter_stop                           exact
(the slot index pindex is not compared: whenever SQUAREM's step length is
clipped to -1 the extrapolated point equals the plain EM point F(F(x)) up to
rounding, accel_em.c:236-237, so accept/reject -- and with it which of two
numerically identical slots is kept -- is decided by the last bit)
and the seven result-file kinds against the same numbers at %f."""
import os
import subprocess

import numpy as np
import pytest

from common import ROOT, ensure_mc_gen, golden_names, load_golden

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "multiclust_b200", "host", "multiclust")
LL_RTOL = 1e-9
PAR_ATOL = 1e-7


def run_cli(tmp_path, g, extra=()):
    from oracle import orc
    gen = g["meta"]["gen"]
    stru = str(tmp_path / "d.stru")
    subprocess.check_call([ensure_mc_gen(), "--I", str(gen["I"]), "--L", str(gen["L"]),
                           "--K", str(gen["K"]), "--jmax", str(gen["jmax"]),
                           "--miss", str(gen["miss"]), "--P", str(gen["P"]), "--stru", stru])
    out = tmp_path / "out"
    out.mkdir()
    trace = str(tmp_path / "trace.txt")
    pre = str(tmp_path / "dump")
    cmd = [CLI, "-f", stru] + g["meta"]["cmd"].split() + [
        "-d", str(out), "--trace", trace, "--dump", pre] + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ll, fit, cur = {}, {}, None
    for line in open(trace):
        w = line.split()
        if w[0] == "init":
            cur = (int(w[1]), int(w[2]))
            ll[cur] = []
        elif w[0] == "ll":
            ll[cur].append(float(w[3]))
        elif w[0] == "fit":
            d = {}
            for kv in w[3:]:
                k, v = kv.split("=")
                d[k] = float(v) if k == "logL" else int(v)
            fit[(int(w[1]), int(w[2]))] = d
    states = {}
    for key in fit:
        for tag in ("start", "final"):
            states[key + (tag,)] = orc.read_state("%s.K%d.init%d.%s.bin" % (pre, key[0], key[1], tag))
    return r, ll, fit, states, out


@pytest.mark.parametrize("name", golden_names())
def test_cli_matches_reference(tmp_path, name):
    g = load_golden(name)
    r, ll, fit, states, out = run_cli(tmp_path, g)
    assert len(fit) == len(g["meta"]["fits"])
    for rec in g["meta"]["fits"]:
        K, init = rec["K"], rec["init"]
        key = "K%d_i%d_" % (K, init)
        st = states[(K, init, "start")]
        assert np.max(np.abs(st["eta"] - g[key + "start_eta"])) < 1e-12
        assert np.max(np.abs(st["p"] - g[key + "start_p"])) < 1e-12
        ref_ll = g[key + "ll"]
        got = np.array(ll[(K, init)])
        assert got.size == ref_ll.size, "number of EM steps"
        assert np.all(np.abs(got - ref_ll) <= LL_RTOL * np.abs(ref_ll))
        f = fit[(K, init)]
        assert abs(f["logL"] - rec["logL"]) <= LL_RTOL * abs(rec["logL"])
        for field in ("n_iter", "converged", "iter_stop"):
            assert f[field] == rec[field], field
        fi = states[(K, init, "final")]
        assert np.max(np.abs(fi["eta"] - g[key + "final_eta"])) < PAR_ATOL
        assert np.max(np.abs(fi["p"] - g[key + "final_p"])) < PAR_ATOL
        assert np.max(np.abs(fi["posterior"] - g[key + "final_post"])) < PAR_ATOL * 10


def test_result_files(tmp_path):
    """names and contents of the result files of an admixture and a mixture fit"""
    for name, kind in (("admix_em", "admix"), ("mix_em", "mix")):
        g = load_golden(name)
        sub = tmp_path / name
        sub.mkdir()
        r, ll, fit, states, out = run_cli(sub, g)
        K = g["meta"]["fits"][0]["K"]
        # the best of the fits is what is left on disk
        best = max(g["meta"]["fits"], key=lambda f: f["logL"])
        key = "K%d_i%d_" % (K, best["init"])
        files = sorted(os.listdir(out))
        base = "d.stru"
        expect = ["%s.%s.K=%d.out.txt" % (base, kind, K), "%s.%s.K=%d.pklm.txt" % (base, kind, K)]
        if kind == "admix":
            expect += ["%s.admix.K=%d.etaik.txt" % (base, K), "%s_admix_popq_%d.popq" % (base, K),
                       "%s_admix_indivq_%d.indivq" % (base, K)]
        else:
            expect += ["%s.mix.K=%d.etak.txt" % (base, K), "%s_mix_popq.popq" % base,
                       "%s.mix.K=%d.indivq" % (base, K)]
        assert files == sorted(expect)
        txt = open(out / expect[0]).read().split("\n")
        assert txt[0].startswith("logL = %f" % best["logL"])
        rows = [l.split("\t") for l in open(out / expect[1]).read().strip().split("\n")[1:]]
        p = np.array([float(x[3]) for x in rows])
        assert np.max(np.abs(p - g[key + "final_p"])) < 1e-6
        # per-initialisation line on stdout (multiclust.c:618-627)
        assert "K = %d, initialization = 0: %f" % (K, g["meta"]["fits"][0]["logL"]) in r.stdout


def test_dash_C_is_dash_T(tmp_path):
    g = load_golden("admix_em")
    g["meta"]["cmd"] = g["meta"]["cmd"].replace("-T", "-C")
    r, ll, fit, states, out = run_cli(tmp_path, g)
    assert fit[(3, 0)]["n_iter"] == g["meta"]["fits"][0]["n_iter"]


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("name", ["admix_em", "admix_s3", "admix_s5", "admix_c_em", "mix_em",
                                  "mix_s1", "admix_k10", "admix_tetra", "admix_biallelic",
                                  "mix_biallelic_k5"])
def test_cli_two_gpus_matches_reference(tmp_path, name):
    """--gpus 2: individuals sharded over two devices, NCCL exchange of the
    sufficient statistics (libmc_comm.so); same parity bars as one device"""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    g = load_golden(name)
    r, ll, fit, states, out = run_cli(tmp_path, g, extra=["--gpus", "2"])
    for rec in g["meta"]["fits"]:
        K, init = rec["K"], rec["init"]
        key = "K%d_i%d_" % (K, init)
        st = states[(K, init, "start")]
        assert np.max(np.abs(st["eta"] - g[key + "start_eta"])) < 1e-12
        assert np.max(np.abs(st["p"] - g[key + "start_p"])) < 1e-12
        ref_ll = g[key + "ll"]
        got = np.array(ll[(K, init)])
        assert got.size == ref_ll.size
        assert np.all(np.abs(got - ref_ll) <= LL_RTOL * np.abs(ref_ll))
        fi = states[(K, init, "final")]
        assert np.max(np.abs(fi["eta"] - g[key + "final_eta"])) < PAR_ATOL
        assert np.max(np.abs(fi["p"] - g[key + "final_p"])) < PAR_ATOL
        assert np.max(np.abs(fi["posterior"] - g[key + "final_post"])) < PAR_ATOL * 10


@pytest.mark.parametrize("name", ["admix_sweep", "mix_sweep", "admix_s1", "mix_em", "admix_c_em"])
def test_cli_shard_fits_matches_reference(tmp_path, name):
    """--gpus 2 --shard-fits: whole fits (K, initialisation) dealt to two
    devices (host/shard_fits.c).  Every fit starts from the reference's rand()
    stream position, so initial parameters, traces and final parameters equal
    the sequential reference's; stdout lines and result files equal those of
    the one-device run of this program."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    g = load_golden(name)
    a = tmp_path / "sharded"
    b = tmp_path / "sequential"
    a.mkdir(), b.mkdir()
    r, ll, fit, states, out = run_cli(a, g, extra=["--gpus", "2", "--shard-fits"])
    assert len(fit) == len(g["meta"]["fits"])
    for rec in g["meta"]["fits"]:
        K, init = rec["K"], rec["init"]
        key = "K%d_i%d_" % (K, init)
        st = states[(K, init, "start")]
        assert np.max(np.abs(st["eta"] - g[key + "start_eta"])) < 1e-12
        assert np.max(np.abs(st["p"] - g[key + "start_p"])) < 1e-12
        ref_ll = g[key + "ll"]
        got = np.array(ll[(K, init)])
        assert got.size == ref_ll.size
        assert np.all(np.abs(got - ref_ll) <= LL_RTOL * np.abs(ref_ll))
        f = fit[(K, init)]
        for field in ("n_iter", "converged", "iter_stop"):
            assert f[field] == rec[field], field
        fi = states[(K, init, "final")]
        assert np.max(np.abs(fi["eta"] - g[key + "final_eta"])) < PAR_ATOL
        assert np.max(np.abs(fi["p"] - g[key + "final_p"])) < PAR_ATOL
    r1, ll1, fit1, states1, out1 = run_cli(b, g)
    # the replayed bookkeeping: same lines in the same order (times blanked)
    import re
    blank = lambda s: re.sub(r"\d\d:\d\d:\d\d", "hh:mm:ss", s).replace(str(a), "").replace(str(b), "")
    assert blank(r.stdout) == blank(r1.stdout)
    assert sorted(os.listdir(out)) == sorted(os.listdir(out1))
    for fn in os.listdir(out):
        assert open(out / fn).read() == open(out1 / fn).read(), fn


# ---- parametric bootstrap (-b) against the stock reference's stdout ----

def _bootstrap_cases():
    import glob
    return sorted(os.path.basename(f)[len("bootstrap_"):-len(".json")]
                  for f in glob.glob(os.path.join(ROOT, "tests", "golden", "bootstrap_*.json")))


def _same_text(got, want):
    """line by line; tokens that are numbers agree to 2e-6 (the reference prints %f), the
    hh:mm:ss fields are not compared"""
    import re
    # (NCCL prints its version banner on stdout when NCCL_DEBUG asks for it)
    g = [x for x in got.strip().splitlines() if not x.startswith("NCCL version")]
    w = want.strip().splitlines()
    assert len(g) == len(w), (len(g), len(w), got[-600:])
    for a, b in zip(g, w):
        ta = re.sub(r"\d\d:\d\d:\d\d", "T", a).replace(",", " ").replace(";", " ")
        tb = re.sub(r"\d\d:\d\d:\d\d", "T", b).replace(",", " ").replace(";", " ")
        ta = ta.replace("(", " ").replace(")", " ").replace("=", " ").split()
        tb = tb.replace("(", " ").replace(")", " ").replace("=", " ").split()
        assert len(ta) == len(tb), (a, b)
        for x, y in zip(ta, tb):
            try:
                fx, fy = float(x), float(y)
            except ValueError:
                assert x == y, (a, b)
                continue
            if fx != fx or fy != fy:    # the reference's 0/0 statistics print as nan
                assert fx != fx and fy != fy, (a, b)
                continue
            assert abs(fx - fy) <= 2e-6 + 1e-9 * abs(fy), (a, b)


@pytest.mark.parametrize("name", ["admix", "mix", "admix_tetra", "mix_biallelic_s1"])
def test_cli_bootstrap_two_gpus_matches_reference(tmp_path, name):
    """-b with --gpus 2: every device draws the sample of its rows of individuals from its
    part of the rand() stream"""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    _run_bootstrap_case(tmp_path, name, ["--gpus", "2"])


@pytest.mark.parametrize("name", _bootstrap_cases())
def test_cli_bootstrap_matches_reference(tmp_path, name):
    """-b n: the observed fits under H0 and Ha, every bootstrap sample drawn from the H0
    estimates with the reference's rand() stream and loops, its two fits, the test statistics
    and the p-value line -- the product binary's stdout against the stock reference's"""
    _run_bootstrap_case(tmp_path, name, [])


def _run_bootstrap_case(tmp_path, name, extra):
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "bootstrap_%s.json" % name)))
    gen = g["gen"]
    subprocess.check_call([ensure_mc_gen(), "--I", str(gen["I"]), "--L", str(gen["L"]),
                           "--K", str(gen["K"]), "--jmax", str(gen["jmax"]),
                           "--miss", str(gen["miss"]), "--P", str(gen["P"]),
                           "--stru", str(tmp_path / "d.stru")], stdout=subprocess.DEVNULL)
    (tmp_path / "out").mkdir()
    r = subprocess.run([CLI, "-f", "d.stru"] + g["cmd"].split() + ["-d", "out/"] + extra,
                       cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    _same_text(r.stdout, g["stdout"])


def _timed_cases():
    import glob
    return sorted(os.path.basename(f)[len("timed_"):-len(".json")]
                  for f in glob.glob(os.path.join(ROOT, "tests", "golden", "timed_*.json")))


def _mask_clock(text):
    """stdout of a `-w` run without its clock readings: the two elapsed-seconds fields of
    every per-repetition line (they follow hh:mm:ss and the four counters) and the
    `Average time` line"""
    out = []
    for line in text.strip().splitlines():
        if line.startswith("Average time:"):
            continue
        w = line.split()
        if len(w) > 20 and w[12] == "ND" and ":" in w[14]:
            w[19] = w[20] = "0"
            line = " ".join(w)
        out.append(line.replace("-nan", "nan"))
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("name", _timed_cases())
def test_cli_timed_matches_reference(tmp_path, name):
    """-w n <r>: the estimation repeated r times on one rand() stream, a compact line per
    repetition and the statistics block of multiclust.c:296-344 (averages, +/- deviations,
    first hit of the maximum, AIC / BIC choice of K in a sweep) -- the product binary's
    stdout against the stock reference's, clock readings aside"""
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "timed_%s.json" % name)))
    gen = g["gen"]
    subprocess.check_call([ensure_mc_gen(), "--I", str(gen["I"]), "--L", str(gen["L"]),
                           "--K", str(gen["K"]), "--jmax", str(gen["jmax"]),
                           "--miss", str(gen["miss"]), "--P", str(gen["P"]),
                           "--stru", str(tmp_path / "d.stru")], stdout=subprocess.DEVNULL)
    r = subprocess.run([CLI, "-f", "d.stru"] + g["cmd"].split(), cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert not list(tmp_path.glob("*.txt")) and not list(tmp_path.glob("*q")), \
        "-w must not write result files"
    _same_text(_mask_clock(r.stdout), _mask_clock(g["stdout"]))
