"""The parametric bootstrap (-b) of the product binary next to the stock reference binary
(oracle/_ref/multiclust, one host core) on the same mc_gen data and command line: wall clock of
both and a line-by-line comparison of their stdout (numbers to 2e-6: the reference prints %f).
  BOOT_I=400 BOOT_L=300 BOOT_ARGS="-a -k 3 -n 3 -b 5 -T 30 -E 1e-30" python tools/bootstrap_time.py"""
import os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gen = os.path.join(ROOT, "multiclust_b200", "host", "mc_gen")
cli = os.path.join(ROOT, "multiclust_b200", "host", "multiclust")
ref = os.path.join(ROOT, "oracle", "_ref", "multiclust")
I, L = os.environ.get("BOOT_I", "400"), os.environ.get("BOOT_L", "300")
args = os.environ.get("BOOT_ARGS", "-a -k 3 -n 3 -b 5 -T 30 -E 1e-30").split()
tmp = tempfile.mkdtemp(prefix="boot_")
subprocess.check_call([gen, "--I", I, "--L", L, "--K", "3", "--jmax", "6", "--miss", "300",
                       "--P", "2", "--stru", os.path.join(tmp, "d.stru")], stdout=subprocess.DEVNULL)
os.mkdir(os.path.join(tmp, "out"))
res = {}
for name, exe in (("reference (1 host core)", ref), ("B200", cli)):
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-f", "d.stru"] + args + ["-d", "out/"], cwd=tmp,
                       capture_output=True, text=True)
    res[name] = (time.perf_counter() - t0, r.returncode, r.stdout)
    print("%-24s %8.2f s  rc=%d  %s" % (name, res[name][0], r.returncode,
                                         r.stdout.strip().splitlines()[-1]), flush=True)


def tokens(line):
    line = re.sub(r"\d\d:\d\d:\d\d", "T", line)
    for ch in ",;()=":
        line = line.replace(ch, " ")
    return line.split()


a = [x for x in res["B200"][2].strip().splitlines() if not x.startswith("NCCL version")]
b = res["reference (1 host core)"][2].strip().splitlines()
bad = int(len(a) != len(b))
worst = 0.0
for x, y in zip(a, b):
    tx, ty = tokens(x), tokens(y)
    if len(tx) != len(ty):
        bad += 1
        continue
    for u, v in zip(tx, ty):
        try:
            fu, fv = float(u), float(v)
        except ValueError:
            bad += u != v
            continue
        worst = max(worst, abs(fu - fv))
        bad += abs(fu - fv) > 2e-6 + 1e-9 * abs(fv)
print("I=%s L=%s %s: %d lines, %d mismatches, largest difference of a printed number %.1e; "
      "speed-up %.0fx" % (I, L, " ".join(args), len(b), bad, worst,
                          res["reference (1 host core)"][0] / res["B200"][0]))
sys.exit(1 if bad else 0)
