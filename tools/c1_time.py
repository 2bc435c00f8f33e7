"""BASELINE config 1 (the reference's own CPU-runnable case: I=200, L=100, <=5
alleles, K=3, -C 500): this program on the GPU against the unmodified reference
binary on one host core, same file, same command line."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gen = os.path.join(ROOT, "multiclust_b200", "host", "mc_gen")
cli = os.path.join(ROOT, "multiclust_b200", "host", "multiclust")
ref = os.path.join(ROOT, "oracle", "_ref", "multiclust")
tmp = tempfile.mkdtemp(prefix="c1_")
stru = os.path.join(tmp, "d.stru")
subprocess.check_call([gen, "--I", "200", "--L", "100", "--K", "3", "--jmax", "5",
                       "--miss", "300", "--P", "2", "--stru", stru])
args = ["-f", stru, "-a", "-k", "3", "-T", "500", "-E", "1e-30", "-n", "1"]
for name, exe, extra in (("b200", cli, ["--timing"]), ("reference (1 core)", ref, [])):
    if not os.path.exists(exe):
        print(name, "not built"); continue
    d = os.path.join(tmp, name.split()[0]); os.makedirs(d)
    t0 = time.perf_counter()
    r = subprocess.run([exe] + args + ["-d", d] + extra, capture_output=True, text=True)
    print("%-20s rc %d wall %.3f s | %s" % (name, r.returncode, time.perf_counter() - t0,
          (r.stderr.strip().splitlines() or [""])[-1][:200]))
    print("    ", (r.stdout.strip().splitlines() or [""])[0][:150])
