#!/usr/bin/env python
"""bench.py -- admixture EM iterations/s at BASELINE.json's headline config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[2], the one the metric is quoted on):
admixture -a, I=100k individuals, L=10k loci, <=20 alleles per locus, K=10,
diploid, 5 % missing, unaccelerated EM.  Genotypes come from the repo's
counter-based generator (include/mc_synth.h; the reference's --simulate cannot
make this workload, SURVEY.md finding 4) and are created directly in HBM.

A "step" is one EM iteration (mc_em_step: E-step + M-step + log likelihood)
over the whole genotype matrix; the log likelihood is read back every step like
the reference's stop() needs it (em_alg.c:195-207).

  value      iterations/s with genotypes and parameters resident in HBM
  e2e        iterations/s of a whole fit driven through the C ABI from HOST
             buffers (a fresh context; the resident-data context is closed
             first): mc_set_data from pinned host memory (H2D of the genotype
             codes) + mc_alloc_model + mc_set_params + `--e2e-iters` (200, the
             configuration's fixed -C 200) x mc_em_step (8-byte D2H each) +
             mc_get_params + mc_get_posterior
  roofline   algorithmic bytes (I*L*P + 16*I*K + 16*K*T) of the
             genotype-streaming kernel / its CUDA-event duration, against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json.  Because the
             FP64 step can never exceed ~10 % of the HBM roofline,
             the block also carries `fp64` (algorithmic FMAs of the launch / the
             FP64 pipe's 64 FMA/clk/SM, from the same live event time) and the
             pipe utilisations `fp64_pipe_frac` / `lsu_wavefront_frac` / the
             DRAM `traffic` of the committed ncu capture -- only when that
             capture was taken from the csrc/ that is running (sha-256 key).
  cpu_baseline  the unmodified reference (oracle/_ref/ref_harness calling the
             reference's own em_step) on ONE host core -- the reference has no
             threads -- on the first `--cpu-indiv` individuals of the same
             workload, extrapolated linearly in I (cost is linear in I,
             em_alg.c:325,650,717; the reference cannot allocate its
             I*K*T-double scratch at full size, SURVEY.md finding 5)
  parity_check  OUTSIDE the timed region, for every N: six golden cases
             generated from the unmodified reference (tests/golden: admix_k10,
             admix_s5 = QN q=2, admix_tetra, mix_s1 = SQUAREM on the gather
             kernels; admix_biallelic on the DMMA kernels, mix_biallelic_k5 on
             the digit-sliced integer kernels) are re-fitted with
             their individuals sharded over the N ranks through the same
             sharded step / acceleration calls, and the log-likelihood
             trajectory (<= 1e-9 relative) and final parameters (<= 1e-7
             absolute) are compared with the reference's.  A failure exits
             non-zero.
  other_configs  the other BASELINE configurations, ms per step: C1 (the
             reference's own CPU-runnable case) and C2 (mixture, biallelic, SQUAREM),
             each with the reference's own em_step on one host core beside it
             (`reference_cpu`: C1 whole, C2 on 500 of its 10k individuals),
             one GPU's share of C5 -- and with --gpus 8 the WHOLE C5 (I=1M
             tetraploid, K=8, QN q=2) sharded over the 8 ranks -- and C4
             (K = 2..12 x 64 initialisations x 100 iterations = 704 whole fits)
             through the drop-in command line, the fits dealt to the N devices
             (`--gpus N --shard-fits --fits-per-gpu 4`, no collective) while the
             ranks' own devices are idle: seconds for the fits phase.

N > 1 (torchrun, one rank per GPU): individuals are sharded.  `value` is weak
scaling (every rank holds I individuals, global I = N * 100k; value = N *
iterations/s, i.e. iterations/s normalised to 100k individuals); `strong` is
the same fit with the headline I = 100k split over the N ranks.  Per iteration
the K x T allele-count sums, the log likelihood and the pooled-eta sums are
summed over ranks with a deterministic reduce-scatter + all-gather
(multiclust_b200/sharding.py) so every rank gets bit-identical parameters.
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "admixture EM iters/s at I=100k,L=10k,K=10; fraction of B200 HBM peak"
UNIT = "iterations/s"
SEED = 20261018
FP64_FMA_PER_CLK_SM = 64        # measured: tools/lds_probe3.cu, tools/dmma_probe.cu


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--I", type=int, default=100000)
    ap.add_argument("--L", type=int, default=10000)
    ap.add_argument("--K", type=int, default=10)
    ap.add_argument("--jmax", type=int, default=20)
    ap.add_argument("--miss-bp", type=int, default=500)
    ap.add_argument("--ploidy", type=int, default=2)
    ap.add_argument("--cpu-indiv", type=int, default=320,
                    help="individuals in the CPU baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-iters", type=int, default=200,
                    help="iterations of the end-to-end fit (BASELINE configs[2]: fixed -C 200)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the other_configs block")
    return ap.parse_args()


def workload_name(a):
    return ("admixture -a I=%d L=%d K=%d <=%d alleles/locus ploidy=%d %.1f%% missing, "
            "unaccelerated EM" % (a.I, a.L, a.K, a.jmax, a.ploidy, a.miss_bp / 100.0))


def config_block(a):
    """identical in both arms (the driver compares them)"""
    return {"workload": workload_name(a), "I": a.I, "L": a.L, "K": a.K, "jmax": a.jmax,
            "ploidy": a.ploidy, "missing_bp": a.miss_bp, "seed": SEED,
            "l2": "inputs (%.2f GB of genotype codes per GPU) exceed the 126 MB L2; no flush"
                  % (a.I * a.L * a.ploidy / 1e9)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def csrc_sha16():
    """key of the committed ncu metrics: the kernels' sources as they are now"""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "multiclust_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as fp:
                h.update(name.encode() + b"\0" + fp.read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([w.strip() for w in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------- CPU baseline

def make_sample_mcb(a, n):
    """first n individuals of the workload (same seed, counter-based generator),
    recoded the way the reference parser would recode them"""
    from multiclust_b200 import build as mcbuild
    gen = os.path.join(ROOT, "multiclust_b200", "host", "mc_gen")
    if not os.path.exists(gen):
        mcbuild.build_host()
    tmp = tempfile.mkdtemp(prefix="mcbench_")
    path = os.path.join(tmp, "sample.mcb")
    subprocess.check_call([gen, "--I", str(n), "--L", str(a.L), "--K", str(a.K),
                           "--jmax", str(a.jmax), "--miss", str(a.miss_bp),
                           "--P", str(a.ploidy), "--seed", str(SEED), "--npops", "1",
                           "--mcb", path])
    return path


def run_reference_sample(a, path, steps):
    """time the reference's own em_step on the sample (1 core: the reference
    is single-threaded); falls back to the C port when oracle/_ref is absent"""
    from oracle import orc
    if orc.have_ref():
        r = orc.run_ref(["-f", "sample", "-a", "-k", str(a.K), "-p", str(a.ploidy),
                         "-n", "1", "-E", "1e-30"], mcb=path, time_steps=steps,
                        timeout=1800)
        if r.returncode != 0 or not r.stdout.strip():
            raise RuntimeError("reference harness failed: " + r.stderr[-400:])
        res = json.loads(r.stdout.strip().splitlines()[-1])
        return res["sec_per_step"], "reference"
    d = orc.read_mcb(path)
    fit = orc.Fit(d["J"], d["codes"], admixture=1, max_iter=steps + 2, abs_error=1e-30)
    fit.alloc(a.K)
    orc.seed(1)
    fit.initialize()
    fit.set_indices(0, 0, 0)
    fit.e_step(); fit.m_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        fit.e_step(); fit.m_step()
    return (time.perf_counter() - t0) / steps, "port"


def reference_step_ms(I, I_sample, L, K, P, jmax, miss, admixture, steps=3):
    """ms per em_step of the unmodified reference (one host core) on the first I_sample of I
    individuals of a configuration, extrapolated linearly in I -- for `other_configs`"""
    from oracle import orc
    if not orc.have_ref():
        raise RuntimeError("oracle/_ref is not built")
    a = argparse.Namespace(L=L, K=K, jmax=jmax, miss_bp=miss, ploidy=P)
    path = make_sample_mcb(a, I_sample)
    try:
        r = orc.run_ref(["-f", "sample", "-k", str(K), "-p", str(P), "-n", "1", "-E", "1e-30"]
                        + (["-a"] if admixture else []), mcb=path, time_steps=steps, timeout=600)
    finally:
        try:
            os.remove(path)
        except OSError:
            pass
    if r.returncode != 0 or not r.stdout.strip():
        raise RuntimeError("reference harness failed: " + r.stderr[-300:])
    sec = json.loads(r.stdout.strip().splitlines()[-1])["sec_per_step"]
    return {"ms_per_em_step": sec * 1e3 * (I / float(I_sample)), "cores": 1, "kind": "reference",
            "sample": "%d of %d individuals, %d timed em_step calls%s" % (
                I_sample, I, steps, "" if I_sample == I else ", extrapolated linearly in I")}


def cpu_baseline(a, steps):
    n = min(a.cpu_indiv, a.I)
    path = make_sample_mcb(a, n)
    try:
        sec, kind = run_reference_sample(a, path, steps)
    finally:
        try:
            os.remove(path)
        except OSError:
            pass
    full = sec * (a.I / float(n))
    return {"value": 1.0 / full, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %d of %d individuals (recoded on the sample), all %d loci, "
                      "%d timed em_step calls (%.3f s each, single thread: the reference has "
                      "no threading), extrapolated linearly in I" % (n, a.I, a.L, steps, sec),
            "sec_per_step_sample": sec}


# ------------------------------------------------------------------ our arm

def random_start(ctx, K, per_indiv, rng_seed):
    """a valid starting point: normalised eta rows and p rows (random initial
    values do not change the per-iteration cost)"""
    import numpy as np
    rng = np.random.default_rng(rng_seed)
    J = ctx.get_J()
    T = int(J.sum())
    eta = rng.random((ctx.I if per_indiv else 1, K)) + 0.1
    eta /= eta.sum(axis=1, keepdims=True)
    p = rng.random((K, T)) + 0.1
    seg = np.repeat(np.arange(len(J)), J)
    sums = np.zeros((K, len(J)))
    for k in range(K):
        sums[k] = np.bincount(seg, weights=p[k], minlength=len(J))
    p /= sums[:, seg]
    return eta.ravel(), p.ravel()


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version
    there when NCCL_DEBUG is set) write to fd 1 as well, so fd 1 is pointed at
    stderr for the rest of the run and the line goes to a private duplicate."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


class Env:
    """rank / device / collective plumbing of one bench process"""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            sys.exit("bench.py: no CUDA device (the EM path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
            # a wait that keeps the devices idle (an NCCL barrier spins on the GPU)
            self.cpu_group = dist.new_group(backend="gloo")
        # one torch stream carries the context's kernels, the NCCL collectives
        # and the timing events (the legacy default stream has handle 0, which
        # mc_set_stream reads as "use the context's own stream")
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def context(self):
        from multiclust_b200 import Context
        ctx = Context(self.local)
        ctx.set_stream(self.stream.cuda_stream)
        return ctx

    def scratch(self, ctx):
        if self.world == 1:
            return None
        n = ctx.exchange_len()
        return self.torch.empty(n + 64, dtype=self.torch.float64, device="cuda")


def timed_steps(env, ctx, scratch, steps, warmup, sample_clocks=False):
    """W warm-up + K timed sharded EM steps; CUDA events on the launch stream,
    barrier + synchronize on both sides, max over ranks"""
    from multiclust_b200.sharding import sharded_em_step
    import numpy as np
    torch = env.torch
    lls = []
    for _ in range(warmup):
        lls.append(sharded_em_step(ctx, env.dist, env.world, 0, 0, scratch))
    ctx.profile_read()
    sampler = ClockSampler(env.local) if sample_clocks and env.rank == 0 else None
    env.barrier()
    if sampler:
        sampler.start()
    ctx.profile_enable(True)
    n0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        lls.append(sharded_em_step(ctx, env.dist, env.world, 0, 0, scratch))
    e1.record()
    env.barrier()
    ms = env.max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count() - n0
    ctx.profile_enable(False)
    nk, kms = ctx.profile_read()
    clocks = sampler.stop() if sampler else None
    if not all(np.isfinite(lls)) or any(b < a_ - 1e-9 * abs(a_) for a_, b in zip(lls, lls[1:])):
        sys.exit("bench.py: log likelihood trajectory is not monotone/finite: %r" % lls[:6])
    return {"ms_per_step": ms / steps, "launches": launches, "kernel_ms": kms / max(nk, 1),
            "kernel_launches": nk, "clocks": clocks, "lls": lls}


def setup_c3(env, a, n_indiv, i_first, seed):
    """synthetic headline workload for this rank: individuals [i_first, i_first + n)"""
    from multiclust_b200 import SynthParams
    ctx = env.context()
    sp = SynthParams(seed=SEED, K=a.K, jmax=a.jmax, miss_bp=a.miss_bp, ploidy=a.ploidy)
    ctx.set_data_synth(n_indiv, a.L, sp, i_first=i_first)
    if env.world > 1:
        # allele slots must agree on every rank: recode on the union
        Jall = [None] * env.world
        env.dist.all_gather_object(Jall, ctx.get_J().tolist())
        if any(j != Jall[0] for j in Jall):
            sys.exit("bench.py: ranks disagree on allele slots; use a larger --I")
    return ctx


def start_params(env, ctx, a, lb, seed):
    ctx.alloc_model(a.K, admixture=1, q=0, eta_lb=lb, p_lb=lb)
    eta0, p0 = random_start(ctx, a.K, True, seed + env.rank)
    if env.world > 1:       # p is replicated: take rank 0's
        pt = env.torch.from_numpy(p0).cuda()
        env.dist.broadcast(pt, 0)
        p0 = pt.cpu().numpy()
    ctx.set_params(0, eta0, p0)
    return eta0, p0


# ------------------------------------------------------------ parity check

PARITY_CASES = ["admix_k10", "admix_s5", "admix_tetra", "mix_s1", "admix_biallelic",
                "mix_biallelic_k5"]


def parity_check(env):
    """golden cases of the unmodified reference, re-fitted with the individuals
    sharded over the ranks of this run"""
    import numpy as np
    from multiclust_b200.em_driver import Driver
    from multiclust_b200.sharding import shard_bounds
    gdir = os.path.join(ROOT, "tests", "golden")
    out = {"ok": True, "max_rel_ll": 0.0, "max_abs_param": 0.0, "cases": {},
           "n_ranks": env.world, "tolerance": {"ll_rel": 1e-9, "param_abs": 1e-7}}
    for name in PARITY_CASES:
        z = np.load(os.path.join(gdir, name + ".npz"))
        meta = json.loads(str(z["meta"]))
        o, fit = meta["options"], meta["fits"][0]
        K, key = fit["K"], "K%d_i%d_" % (fit["K"], fit["init"])
        codes, J = z["codes"], z["J"]
        I = codes.shape[0]
        per_indiv = bool(o["admixture"] and not o["eta_constrained"])
        lo, hi = shard_bounds(I, env.world)[env.rank]
        ctx = env.context()
        ctx.set_data(J, codes[lo:hi])
        q = 0 if not o["accel"] else (o["accel"] - 3 if o["accel"] > 4 else 1)
        ctx.alloc_model(K, admixture=o["admixture"], eta_constrained=o["eta_constrained"], q=q,
                        eta_lb=meta["bound"], p_lb=meta["bound"],
                        do_projection=o["do_projection"])
        eta0 = z[key + "start_eta"]
        ctx.set_params(0, eta0.reshape(I, K)[lo:hi].ravel() if per_indiv else eta0,
                       z[key + "start_p"])
        drv = Driver(ctx, env.dist, env.world, admixture=o["admixture"],
                     eta_constrained=o["eta_constrained"], accel=o["accel"],
                     max_iter=o["max_iter"], abs_error=o["abs_error"], rel_error=o["rel_error"],
                     gathered=env.scratch(ctx))
        drv.em(K)
        ref_ll = z[key + "ll"]
        ll = np.array(drv.trace)
        rel = float(np.max(np.abs(ll - ref_ll) / np.abs(ref_ll))) if ll.shape == ref_ll.shape \
            else float("inf")
        eta, p = ctx.get_params(drv.pindex)
        ref_eta = z[key + "final_eta"]
        if per_indiv:
            ref_eta = ref_eta.reshape(I, K)[lo:hi].ravel()
        dpar = max(float(np.max(np.abs(eta - ref_eta))) if eta.size else 0.0,
                   float(np.max(np.abs(p - z[key + "final_p"]))))
        dpar = env.max_over_ranks(dpar)
        ok = rel <= 1e-9 and dpar <= 1e-7 and drv.n_iter == fit["n_iter"]
        out["cases"][name] = {"ok": bool(ok), "rel_ll": rel, "abs_param": dpar,
                              "n_iter": drv.n_iter, "kernel": ctx.plan()["two_pass"]}
        out["ok"] = bool(out["ok"] and ok)
        out["max_rel_ll"] = max(out["max_rel_ll"], rel)
        out["max_abs_param"] = max(out["max_abs_param"], dpar)
        ctx.close()
    return out


# ----------------------------------------------------- other configurations

def other_configs(env, a):
    """ms per step of the other BASELINE configurations (rank 0 alone for the
    single-GPU ones; every rank for the sharded C5)"""
    import numpy as np
    from multiclust_b200 import SynthParams
    from multiclust_b200.em_driver import Driver
    from multiclust_b200.sharding import sharded_em_step
    torch = env.torch
    res = {}

    def fit_ms(ctx, n, fn):
        for _ in range(3):      # the launch graph is captured on the second call
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def single(name, I, L, K, P, jmax, miss, admixture, accel, steps, ref_sample=0):
        ctx = env.context()
        sp = SynthParams(seed=SEED, K=min(K, 8), jmax=jmax, miss_bp=miss, ploidy=P)
        ctx.set_data_synth(I, L, sp)
        lb = min(1e-8, 0.5 / I / P)
        q = 0 if not accel else (accel - 3 if accel > 4 else 1)
        ctx.alloc_model(K, admixture=admixture, q=q, eta_lb=lb, p_lb=lb)
        eta, p = random_start(ctx, K, bool(admixture), 11)
        ctx.set_params(0, eta, p)
        out = {"I": I, "L": L, "K": K, "ploidy": P, "kernel": ctx.plan()["two_pass"],
               "ms_per_em_step": fit_ms(ctx, steps, lambda: ctx.em_step(0, 0)),
               "ms_per_loglik": fit_ms(ctx, steps, lambda: ctx.loglik(0))}
        if accel:
            ctx.set_params(0, eta, p)
            drv = Driver(ctx, None, 1, admixture=admixture, accel=accel, max_iter=10 ** 9,
                         abs_error=1e-300)      # never converges: every call is a full cycle
            for _ in range(1, drv.q):
                drv.em_2_steps()
                drv.pindex = drv.findex
            out["ms_per_accelerated_cycle"] = fit_ms(ctx, max(steps // 4, 3),
                                                     drv.accelerated_em_step)
            out["accel"] = accel
        ctx.close()
        if ref_sample and not a.no_cpu:
            # the reference's own em_step beside it (reported, never required)
            try:
                out["reference_cpu"] = reference_step_ms(I, min(I, ref_sample), L, K, P, jmax,
                                                         miss, admixture)
            except Exception as exc:
                out["reference_cpu"] = {"error": repr(exc)[:200]}
        res[name] = out

    if env.rank == 0:
        single("C1 admixture I=200 L=100 K=3 <=5 alleles (-C 500 fit: x500)",
               200, 100, 3, 2, 5, 300, 1, 0, 50, ref_sample=200)
        single("C2 mixture I=10k L=5k K=5 biallelic diploid, SQUAREM -s 1",
               10000, 5000, 5, 2, 2, 0, 0, 1, 20, ref_sample=500)
        if env.world != 8:      # with 8 ranks the whole configuration runs below
            single("C5 share: admixture I=125k (1M / 8) L=50k K=8 biallelic tetraploid, QN -s 5",
                   125000, 50000, 8, 4, 2, 0, 1, 5, 5)
    if env.world == 8:
        # the whole configuration 5: one million tetraploid individuals over 8 ranks
        I, L, K, P = 125000, 50000, 8, 4
        ctx = env.context()
        sp = SynthParams(seed=SEED, K=8, jmax=2, miss_bp=0, ploidy=P)
        ctx.set_data_synth(I, L, sp, i_first=env.rank * I)
        lb = min(1e-8, 0.5 / (I * env.world) / P)
        ctx.alloc_model(K, admixture=1, q=2, eta_lb=lb, p_lb=lb)
        eta, p = random_start(ctx, K, True, 13 + env.rank)
        pt = torch.from_numpy(p).cuda()
        env.dist.broadcast(pt, 0)
        ctx.set_params(0, eta, pt.cpu().numpy())
        scratch = env.scratch(ctx)
        env.barrier()
        em = env.max_over_ranks(fit_ms(ctx, 5, lambda: sharded_em_step(
            ctx, env.dist, env.world, 0, 0, scratch)))
        drv = Driver(ctx, env.dist, env.world, admixture=1, accel=5, max_iter=10 ** 9,
                     abs_error=1e-300, gathered=scratch)
        drv.em_2_steps()
        drv.pindex = drv.findex
        env.barrier()
        cyc = env.max_over_ranks(fit_ms(ctx, 3, drv.accelerated_em_step))
        res["C5 whole: admixture I=1M L=50k K=8 biallelic tetraploid, QN -s 5, 8 ranks"] = {
            "I": I * env.world, "L": L, "K": K, "ploidy": P, "kernel": ctx.plan()["two_pass"],
            "ms_per_em_step": em, "ms_per_accelerated_cycle": cyc, "accel": 5,
            "accelerated_steps_accepted": int(drv.accel_step), "logL": drv.logL}
        ctx.close()
    # configuration 4 through the command line, on every device of the job; the ranks
    # wait on the CPU so that their devices are idle meanwhile
    env.cpu_barrier()
    if env.rank == 0:
        try:
            res["C4 multi-start: admixture I=2000 L=1000, K=2..12 x -n 64, -C 100, "
                "fits dealt to %d device(s)" % env.world] = c4_shard_fits(env.world)
        except Exception as exc:        # reported, never required
            res["C4 multi-start"] = {"error": repr(exc)[:300]}
    env.cpu_barrier()
    return res


def c4_shard_fits(n_gpus, n_init=64, k_min=2, k_max=12, iters=100, per_gpu=4):
    """BASELINE config 4 (multiclust.c:471-660 dealt to devices, host/shard_fits.c): the
    product binary on mc_gen data with whole fits dealt to `n_gpus` devices, `per_gpu` fits in
    flight on each; the binary's own --timing line gives the fits phase"""
    import re
    host = os.path.join(ROOT, "multiclust_b200", "host")
    tmp = tempfile.mkdtemp(prefix="c4_")
    stru = os.path.join(tmp, "d.stru")
    subprocess.check_call([os.path.join(host, "mc_gen"), "--I", "2000", "--L", "1000", "--K", "4",
                           "--jmax", "6", "--miss", "200", "--P", "2", "--stru", stru],
                          stdout=subprocess.DEVNULL)
    cmd = [os.path.join(host, "multiclust"), "-f", stru, "-a", "-1", str(k_min), "-2", str(k_max),
           "-n", str(n_init), "-C", str(iters), "-E", "1e-30", "--timing", "-d", tmp + "/",
           "--gpus", str(n_gpus), "--shard-fits", "--fits-per-gpu", str(per_gpu)]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    wall = time.perf_counter() - t0
    fits = r.stdout.count("initialization =")
    m = re.search(r"(\d+) workers on (\d+) devices.*fits phase ([0-9.]+)", r.stderr)
    if r.returncode != 0 or not m or fits != (k_max - k_min + 1) * n_init:
        raise RuntimeError("multiclust --shard-fits: rc %d, %d fits: %s"
                           % (r.returncode, fits, r.stderr[-200:]))
    phase = float(m.group(3))
    return {"fits": fits, "em_iterations_per_fit": iters, "workers": int(m.group(1)),
            "devices": int(m.group(2)), "fits_phase_s": phase,
            "fits_per_s": fits / phase if phase > 0 else None,
            "us_per_em_iteration": phase / (fits * (iters + 1)) * 1e6,
            "wall_s_with_cuda_startup": round(wall, 3),
            "command": "multiclust " + " ".join(c for c in cmd[3:] if c not in ("-d", tmp + "/"))}


def ncu_metrics(a):
    """pipe utilisations / DRAM bytes of the dominant kernel from the committed
    ncu capture -- only if it was taken from the sources that are running"""
    path = os.path.join(ROOT, "profiles", "ncu_metrics.json")
    try:
        m = json.load(open(path))
    except Exception:
        return None, "no profiles/ncu_metrics.json"
    if m.get("csrc_sha16") != csrc_sha16():
        return None, "profiles/ncu_metrics.json is from other kernel sources (%s, running %s)" % (
            m.get("csrc_sha16"), csrc_sha16())
    if (m.get("I"), m.get("L"), m.get("K")) != (a.I, a.L, a.K):
        return None, "profiles/ncu_metrics.json is for another workload"
    return m, "profiles/ncu_metrics.json (%s)" % m.get("capture", "")


def main():
    a = parse_args()
    out = _claim_stdout()
    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        return reference_arm(a, out)

    import numpy as np
    env = Env()
    torch, world, rank = env.torch, env.world, env.rank

    # ---- the headline workload: weak scaling, I individuals per rank ----
    ctx = setup_c3(env, a, a.I, rank * a.I, 7)
    lb = min(1e-8, 0.5 / (a.I * world) / a.ploidy)
    eta0, p0 = start_params(env, ctx, a, lb, 7)
    plan = ctx.plan()
    scratch = env.scratch(ctx)
    W = max(a.warmup, 3)
    tw = timed_steps(env, ctx, scratch, a.steps, W, sample_clocks=True)
    ms_per_step = tw["ms_per_step"]
    value = world * 1000.0 / ms_per_step

    # ---- end to end through the C ABI from host buffers (rank-local fit) ----
    e2e = None
    if not a.no_e2e:
        from multiclust_b200 import Context
        from multiclust_b200.sharding import sharded_em_step
        codes = np.empty((ctx.I, ctx.L, ctx.P), dtype=np.uint8)
        pinned = torch.from_numpy(codes).pin_memory()
        codes_p = pinned.numpy()
        ctx.lib.mc_get_codes(ctx.h, ctypes.c_void_p(codes_p.ctypes.data))
        J = ctx.get_J()
        eta_h = torch.from_numpy(eta0).pin_memory().numpy()
        p_h = torch.from_numpy(p0).pin_memory().numpy()
        eta_o = torch.empty(eta0.size, dtype=torch.float64).pin_memory().numpy()
        p_o = torch.empty(p0.size, dtype=torch.float64).pin_memory().numpy()
        post_o = torch.empty(eta0.size, dtype=torch.float64).pin_memory().numpy()
        # the resident-data context is done; its device memory goes back to the
        # library's pool, as it would between two fits of one process
        ctx.close()
        ctx2 = Context(env.local)
        env.barrier()
        t0 = time.perf_counter()
        ctx2.set_data(J, codes_p)
        t1 = time.perf_counter()
        ctx2.alloc_model(a.K, admixture=1, q=0, eta_lb=lb, p_lb=lb)
        ctx2.set_params(0, eta_h, p_h)
        ctx2.set_stream(env.stream.cuda_stream)
        t2 = time.perf_counter()
        n_it = max(a.e2e_iters, 1)
        for _ in range(n_it):
            sharded_em_step(ctx2, env.dist, world, 0, 0, scratch)
        t3 = time.perf_counter()
        ctx2.lib.mc_get_params(ctx2.h, 0, ctypes.c_void_p(eta_o.ctypes.data),
                               ctypes.c_void_p(p_o.ctypes.data))
        ctx2.lib.mc_get_posterior(ctx2.h, ctypes.c_void_p(post_o.ctypes.data))
        env.barrier()
        t4 = time.perf_counter()
        sec = env.max_over_ranks(t4 - t0)
        phases = {"set_data_s": t1 - t0, "alloc_model_plan_set_params_s": t2 - t1,
                  "em_steps_s": t3 - t2, "get_results_s": t4 - t3}
        ctx2.close()
        h2d = codes.nbytes + J.nbytes + eta0.nbytes + p0.nbytes
        d2h = 8 * n_it + eta0.nbytes + p0.nbytes + post_o.nbytes
        e2e = {"value": world * n_it / sec, "unit": UNIT, "iterations": n_it,
               "h2d_bytes_per_step": h2d // n_it, "d2h_bytes_per_step": d2h // n_it,
               "phases_rank0": phases,
               "what": "whole fit of %d iterations (the configuration's fixed -C %d) from pinned "
                       "host buffers: mc_set_data + mc_alloc_model + mc_set_params + mc_em_step "
                       "x%d + mc_get_params + mc_get_posterior; bytes amortised per iteration"
                       % (n_it, n_it, n_it)}
        del codes, pinned, codes_p
    else:
        ctx.close()

    # ---- strong scaling: the headline I split over the ranks ----
    strong = None
    if world > 1:
        n_loc = a.I // world
        ctx3 = setup_c3(env, a, n_loc, rank * n_loc, 7)
        start_params(env, ctx3, a, min(1e-8, 0.5 / a.I / a.ploidy), 7)
        ts = timed_steps(env, ctx3, env.scratch(ctx3), a.steps, W)
        strong = {"ms_per_step": ts["ms_per_step"], "value": 1000.0 / ts["ms_per_step"],
                  "unit": UNIT, "individuals_per_gpu": n_loc, "global_individuals": n_loc * world,
                  "kernel_ms": ts["kernel_ms"]}
        ctx3.close()

    # ---- parity: a wrong kernel must not get a number (run after the timed
    # sections so that its small contexts do not disturb the memory pool the
    # end-to-end fit allocates from; nothing is printed before it passes) ----
    parity = None
    if not a.no_parity:
        parity = parity_check(env)
        if not parity["ok"]:
            sys.stderr.write("bench.py: parity check failed: %s\n" % json.dumps(parity))
            if rank == 0:
                out.write(json.dumps({"metric": METRIC, "value": None, "unit": UNIT,
                                      "n_gpus": world, "parity_check": parity,
                                      "error": "parity check failed"}) + "\n")
                out.flush()
            return 3

    others = None
    if not a.no_other:
        others = other_configs(env, a)

    if rank != 0:
        if world > 1:
            env.dist.barrier()
            env.dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    alg = plan["algorithmic_bytes_em"]
    k_ms, nk = tw["kernel_ms"], tw["kernel_launches"]
    achieved = alg / (k_ms * 1e-3) / 1e9 if nk else None
    ncu, ncu_src = ncu_metrics(a)
    sm_mhz = (tw["clocks"] or {}).get("sm_mhz") or 1965.0
    n_sms = torch.cuda.get_device_properties(env.local).multi_processor_count
    # algorithmic FMAs of one launch: 3K per non-missing allele copy (tmp, A, G)
    copies = a.I * a.L * a.ploidy * (1.0 - a.miss_bp / 1e4)
    fma = copies * 3 * a.K
    fp64_peak = FP64_FMA_PER_CLK_SM * n_sms * sm_mhz * 1e6
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": ncu.get("dram_bytes_per_launch") if ncu else None,
                "peak_source": peak_src,
                "kernel": {3: "dense_kernel<ADMIX_EM> (DMMA)",
                           2: "admix3_kernel<MODE_EM> (two-pass gather)"}.get(
                               plan.get("two_pass"), "tile_kernel<MODE_ADMIX_EM>"),
                "kernel_ms": k_ms, "kernel_launches_timed": nk,
                "algorithmic_bytes_per_launch": alg,
                "kernel_share_of_step": (k_ms / ms_per_step) if nk else None,
                "fp64": {"algorithmic_fma_per_launch": fma,
                         "peak_fma_per_s": fp64_peak,
                         "frac": (fma / (k_ms * 1e-3) / fp64_peak) if nk else None,
                         "note": "3K FMAs per non-missing allele copy against %d FMA/clk/SM x %d "
                                 "SMs x %.0f MHz (measured pipe rate)"
                                 % (FP64_FMA_PER_CLK_SM, n_sms, sm_mhz)},
                "fp64_pipe_frac": (ncu.get("fp64_pipe_pct") / 100.0) if ncu else None,
                "lsu_wavefront_frac": (ncu.get("lsu_wavefront_pct") / 100.0) if ncu else None,
                "ncu_source": ncu_src}

    cpu = None
    if not a.no_cpu:
        try:
            cpu = cpu_baseline(a, a.cpu_steps)
        except Exception as exc:  # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": "failed: %s" % exc}

    lls = tw["lls"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (include/mc_synth.h, seed %d), random-init parameters" % SEED,
        "config": config_block(a),
        "details": {"individuals_per_gpu": a.I, "global_individuals": a.I * world,
                    "value_definition": "n_gpus x iterations/s (iterations/s normalised to %d "
                                        "individuals)" % a.I,
                    "plan": {k: plan[k] for k in ("two_pass", "k_split", "k_per_lane", "warps",
                                                  "groups", "n_tiles", "n_chunks", "grid",
                                                  "block", "smem_bytes")},
                    "csrc_sha16": csrc_sha16()},
        "clocks": tw["clocks"], "e2e": e2e, "gpu_launches": int(tw["launches"]),
        "roofline": roofline, "cpu_baseline": cpu, "parity_check": parity,
        "strong": strong, "other_configs": others,
        "logL_first": lls[0], "logL_last": lls[-1],
    }
    out.write(json.dumps(line) + "\n")
    out.flush()
    if world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return 0


def reference_arm(a, out):
    """--impl reference: the reference's own CPU implementation (oracle/_ref,
    compiled in place from the unmodified sources) on a bounded sample."""
    t0 = time.perf_counter()
    cpu = cpu_baseline(a, max(1, min(a.steps, a.cpu_steps)))
    value = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (include/mc_synth.h, seed %d)" % SEED,
        "config": config_block(a),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    out.write(json.dumps(line) + "\n")
    out.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
