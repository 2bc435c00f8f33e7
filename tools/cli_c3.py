"""End-to-end wall clock of the drop-in command line at BASELINE config 3:
the synthetic genotypes are made on the device (same generator as bench.py),
written as an MCB1 file, and `multiclust -a -k 10 -C <n> --timing` is run on it."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiclust_b200 import Context, SynthParams
from oracle import orc

I, L, K = int(os.environ.get("C3_I", 100000)), int(os.environ.get("C3_L", 10000)), 10
ctx = Context(0)
ctx.set_data_synth(I, L, SynthParams(seed=20261018, K=K, jmax=20, miss_bp=500, ploidy=2))
J, codes = ctx.get_J(), ctx.get_codes()
ctx.close()
has_missing = (codes == 255).any(axis=(0, 2))
nreal = J - has_missing.astype(J.dtype)
labels = np.concatenate([np.arange(1, n + 1) for n in nreal]).astype(np.int32)
tmp = tempfile.mkdtemp(prefix="c3_")
path = os.path.join(tmp, "c3.mcb")
orc.write_mcb(path, J, nreal, labels, np.zeros(I, np.int32), codes, npops=1)
del codes
cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "multiclust_b200", "host", "multiclust")
for extra in (["-C", "20"], ["-C", "200"], ["-C", "200", "--gpus", os.environ.get("C3_GPUS", "1")]):
    if extra[-2] == "--gpus" and extra[-1] == "1":
        continue
    t0 = time.perf_counter()
    r = subprocess.run([cli, "-f", path, "-a", "-k", str(K), "-E", "1e-30", "-n", "1",
                        "-d", tmp, "--timing"] + extra, capture_output=True, text=True)
    print(" ".join(extra), "-> rc", r.returncode, "wall %.2f s" % (time.perf_counter() - t0))
    print("   ", " | ".join(r.stderr.strip().splitlines()[-2:]))
    print("   ", r.stdout.strip().splitlines()[-1][:160] if r.stdout.strip() else "")
