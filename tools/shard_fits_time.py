"""Wall clock of a K sweep x multi-start run (BASELINE config 4 shape, reduced)
on one device and with the fits dealt to N devices (--shard-fits)."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gen = os.path.join(ROOT, "multiclust_b200", "host", "mc_gen")
cli = os.path.join(ROOT, "multiclust_b200", "host", "multiclust")
tmp = tempfile.mkdtemp(prefix="c4_")
stru = os.path.join(tmp, "d.stru")
I, L = int(os.environ.get("C4_I", 2000)), int(os.environ.get("C4_L", 1000))
subprocess.check_call([gen, "--I", str(I), "--L", str(L), "--K", "4", "--jmax", "6",
                       "--miss", "200", "--P", "2", "--stru", stru])
n_gpus = int(os.environ.get("C4_GPUS", "2"))
base = [cli, "-f", stru, "-a", "-1", "2", "-2", "9", "-n", os.environ.get("C4_N", "16"),
        "-C", "100", "-E", "1e-30", "--timing"]
outs = []
for extra in ([], ["--gpus", str(n_gpus), "--shard-fits"]):
    d = os.path.join(tmp, "out%d" % len(outs)); os.makedirs(d)
    t0 = time.perf_counter()
    r = subprocess.run(base + ["-d", d] + extra, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    outs.append((r.stdout, d))
    print(r.stderr.strip().splitlines()[-1][:220] if r.stderr.strip() else "")
    print("%-28s rc %d  wall %.2f s  (%d fits)" % (" ".join(extra) or "one device", r.returncode, dt,
                                                r.stdout.count("initialization =")))
import re
blank = lambda s, d: re.sub(r"\d\d:\d\d:\d\d", "", s).replace(d, "")
print("stdout identical:", blank(outs[0][0], outs[0][1]) == blank(outs[1][0], outs[1][1]))
same = all(open(os.path.join(outs[0][1], f)).read() == open(os.path.join(outs[1][1], f)).read()
           for f in os.listdir(outs[0][1]))
print("result files identical:", same, sorted(os.listdir(outs[0][1]))[:3])
