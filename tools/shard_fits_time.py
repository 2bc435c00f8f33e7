"""Wall clock of a K sweep x multi-start run (BASELINE config 4: -n 64, K = 2..12, fixed
-C 100 iterations per fit, on mc_gen data I=2000 L=1000) on one device, with several fits
in flight per device (--fits-per-gpu) and with the fits dealt to N devices (--shard-fits).
Writes one JSON line with the timings (profiles/r02_c4_shard_fits.json).
  C4_GPUS="1 2 4 8"  C4_FPG="1 4 8"  C4_N=64  C4_K2=12"""
import json, os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gen = os.path.join(ROOT, "multiclust_b200", "host", "mc_gen")
cli = os.path.join(ROOT, "multiclust_b200", "host", "multiclust")
tmp = tempfile.mkdtemp(prefix="c4_")
stru = os.path.join(tmp, "d.stru")
I, L = int(os.environ.get("C4_I", 2000)), int(os.environ.get("C4_L", 1000))
subprocess.check_call([gen, "--I", str(I), "--L", str(L), "--K", "4", "--jmax", "6",
                       "--miss", "200", "--P", "2", "--stru", stru])
gpus = [int(x) for x in os.environ.get("C4_GPUS", "1 2").split()]
fpgs = [int(x) for x in os.environ.get("C4_FPG", "1 4").split()]
n_init, k2, iters = os.environ.get("C4_N", "64"), os.environ.get("C4_K2", "12"), 100
base = [cli, "-f", stru, "-a", "-1", "2", "-2", k2, "-n", n_init,
        "-C", str(iters), "-E", "1e-30", "--timing"]
runs = [("one device, sequential", [])]
for g in gpus:
    for f in fpgs:
        if g == 1 and f == 1:
            continue
        runs.append(("--gpus %d --shard-fits --fits-per-gpu %d" % (g, f),
                     ["--gpus", str(g), "--shard-fits", "--fits-per-gpu", str(f)]))
blank = lambda s, d: re.sub(r"\d\d:\d\d:\d\d", "", s).replace(d, "")
res, ref = [], None
for name, extra in runs:
    d = os.path.join(tmp, "out%d" % len(res)); os.makedirs(d)
    t0 = time.perf_counter()
    r = subprocess.run(base + ["-d", d] + extra, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    fits = r.stdout.count("initialization =")
    files = {f: open(os.path.join(d, f)).read() for f in sorted(os.listdir(d))}
    out = (blank(r.stdout, d), files)
    if ref is None:
        ref = out
    row = {"run": name, "rc": r.returncode, "wall_s": round(dt, 3), "fits": fits,
           "us_per_em_iteration": round(dt / max(fits * (iters + 1), 1) * 1e6, 2),
           "identical_to_sequential": out == ref,
           "timing": (r.stderr.strip().splitlines() or [""])[-1][:300]}
    res.append(row)
    print(json.dumps(row), flush=True)
print(json.dumps({"config": "C4: I=%d L=%d admixture K=2..%s x -n %s, -C %d" % (I, L, k2, n_init, iters),
                  "runs": res}))
