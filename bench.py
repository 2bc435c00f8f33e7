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
             first, so its device memory is back in the library's pool): mc_set_data from pinned host memory (H2D of the genotype
             codes) + mc_alloc_model + mc_set_params + `steps` x mc_em_step
             (8-byte D2H each) + mc_get_params + mc_get_posterior
  roofline   algorithmic bytes (I*L*P + 16*I*K + 16*K*T) of the
             genotype-streaming kernel / its CUDA-event duration, against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference (oracle/_ref/ref_harness calling the
             reference's own em_step) on ONE host core -- the reference has no
             threads -- on the first `--cpu-indiv` individuals of the same
             workload, extrapolated linearly in I (cost is linear in I,
             em_alg.c:325,650,717; the reference cannot allocate its
             I*K*T-double scratch at full size, SURVEY.md finding 5)

N > 1 (torchrun, one rank per GPU): individuals are sharded, every rank holds
I individuals (weak scaling, global I = N * 100k); per iteration the K x T
allele-count sums, the log likelihood and the pooled-eta sums are exchanged with
one NCCL all-gather and added in rank order (mc_exchange_sum) so every rank
gets bit-identical parameters.  value = N * iterations/s, i.e. iterations/s
normalised to the 100k-individual configuration.
"""
import argparse
import ctypes
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
    return ap.parse_args()


def workload_name(a):
    return ("admixture -a I=%d L=%d K=%d <=%d alleles/locus ploidy=%d %.1f%% missing, "
            "unaccelerated EM" % (a.I, a.L, a.K, a.jmax, a.ploidy, a.miss_bp / 100.0))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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

def init_params(a, ctx, rng_seed):
    """a valid starting point: Dirichlet-like eta rows and p rows (random
    initial values do not change the per-iteration cost)"""
    import numpy as np
    rng = np.random.default_rng(rng_seed)
    J = ctx.get_J()
    T = int(J.sum())
    eta = rng.random((ctx.I, a.K)) + 0.1
    eta /= eta.sum(axis=1, keepdims=True)
    p = rng.random((a.K, T)) + 0.1
    off = np.concatenate([[0], np.cumsum(J)])
    seg = np.repeat(np.arange(len(J)), J)
    sums = np.zeros((a.K, len(J)))
    for k in range(a.K):
        sums[k] = np.bincount(seg, weights=p[k], minlength=len(J))
    p /= sums[:, seg]
    del off
    return eta.ravel(), p.ravel()


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version
    there when NCCL_DEBUG is set) write to fd 1 as well, so fd 1 is pointed at
    stderr for the rest of the run and the line goes to a private duplicate."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    a = parse_args()
    out = _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(a, out)

    import numpy as np
    import torch
    from multiclust_b200 import Context, SynthParams

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device (the EM path has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # one torch stream carries the context's kernels, the NCCL all-gather and
    # the timing events (the legacy default stream has handle 0, which
    # mc_set_stream reads as "use the context's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = Context(local)
    ctx.set_stream(stream.cuda_stream)
    sp = SynthParams(seed=SEED, K=a.K, jmax=a.jmax, miss_bp=a.miss_bp, ploidy=a.ploidy)
    # rank r owns individuals [r*I, (r+1)*I) of the global synthetic population
    ctx.set_data_synth(a.I, a.L, sp, i_first=rank * a.I)
    if world > 1:
        # allele slots must agree on every rank: recode on the union
        Jall = [None] * world
        dist.all_gather_object(Jall, ctx.get_J().tolist())
        if any(j != Jall[0] for j in Jall):
            sys.exit("bench.py: ranks disagree on allele slots; use a larger --I")
    lb = min(1e-8, 0.5 / (a.I * world) / a.ploidy)
    ctx.alloc_model(a.K, admixture=1, q=0, eta_lb=lb, p_lb=lb)
    plan = ctx.plan()
    eta0, p0 = init_params(a, ctx, 7 + rank)
    if world > 1:
        # p is replicated: take rank 0's
        pt = torch.from_numpy(p0).cuda()
        dist.broadcast(pt, 0)
        p0 = pt.cpu().numpy()
    ctx.set_params(0, eta0, p0)

    from multiclust_b200.sharding import sharded_em_step
    gathered = None
    if world > 1:
        gathered = torch.empty(world * ctx.exchange_buffer()[1], dtype=torch.float64,
                               device="cuda")

    def one_step():
        return sharded_em_step(ctx, dist, world, 0, 0, gathered)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lls = []
    for _ in range(max(a.warmup, 3)):
        lls.append(one_step())
    ctx.profile_read()

    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ctx.profile_enable(True)
    n0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        lls.append(one_step())
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - n0
    ctx.profile_enable(False)
    nk, kms = ctx.profile_read()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if not all(np.isfinite(lls)) or any(b < a_ - 1e-9 * abs(a_) for a_, b in zip(lls, lls[1:])):
        sys.exit("bench.py: log likelihood trajectory is not monotone/finite: %r" % lls[:6])

    ms_per_step = ms / a.steps
    value = world * 1000.0 / ms_per_step

    # ---- end to end through the C ABI from host buffers (rank-local fit) ----
    e2e = None
    if not a.no_e2e:
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
        ctx2 = Context(local)
        barrier()
        t0 = time.perf_counter()
        ctx2.set_data(J, codes_p)
        ctx2.alloc_model(a.K, admixture=1, q=0, eta_lb=lb, p_lb=lb)
        ctx2.set_params(0, eta_h, p_h)
        ctx2.set_stream(stream.cuda_stream)
        for _ in range(a.steps):
            sharded_em_step(ctx2, dist, world, 0, 0, gathered)
        ctx2.lib.mc_get_params(ctx2.h, 0, ctypes.c_void_p(eta_o.ctypes.data),
                               ctypes.c_void_p(p_o.ctypes.data))
        ctx2.lib.mc_get_posterior(ctx2.h, ctypes.c_void_p(post_o.ctypes.data))
        barrier()
        sec = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        ctx2.close()
        h2d = codes.nbytes + J.nbytes + eta0.nbytes + p0.nbytes
        d2h = 8 * a.steps + eta0.nbytes + p0.nbytes + post_o.nbytes
        e2e = {"value": world * a.steps / sec, "unit": UNIT,
               "h2d_bytes_per_step": h2d // a.steps, "d2h_bytes_per_step": d2h // a.steps,
               "what": "whole fit of %d iterations from pinned host buffers: mc_set_data + "
                       "mc_alloc_model + mc_set_params + mc_em_step x%d + mc_get_params + "
                       "mc_get_posterior; bytes amortised per iteration" % (a.steps, a.steps)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    alg = plan["algorithmic_bytes_em"]
    k_ms = kms / max(nk, 1)
    achieved = alg / (k_ms * 1e-3) / 1e9 if nk else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("I") == a.I and tj.get("L") == a.L and tj.get("K") == a.K:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": peak_src,
                "kernel": "admix3_kernel<MODE_EM> (two-pass, residue-matched gather)"
                if plan.get("two_pass") else "tile_kernel<MODE_ADMIX_EM>",
                "kernel_ms": k_ms, "kernel_launches_timed": nk,
                "algorithmic_bytes_per_launch": alg,
                "kernel_share_of_step": (k_ms / ms_per_step) if nk else None}

    cpu = None
    if not a.no_cpu:
        try:
            cpu = cpu_baseline(a, a.cpu_steps)
        except Exception as exc:  # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": "failed: %s" % exc}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (include/mc_synth.h, seed %d), random-init parameters" % SEED,
        "config": {"workload": workload_name(a),
                   "individuals_per_gpu": a.I, "global_individuals": a.I * world,
                   "value_definition": "n_gpus x iterations/s (iterations/s normalised to %d "
                                       "individuals)" % a.I,
                   "l2": "inputs (%.2f GB of genotype codes per GPU) exceed the 126 MB L2; no flush"
                         % (a.I * a.L * a.ploidy / 1e9),
                   "plan": {k: plan[k] for k in ("two_pass", "k_split", "k_per_lane", "warps",
                                                 "groups", "n_tiles", "n_chunks", "grid",
                                                 "block", "smem_bytes")}},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
        "logL_first": lls[0], "logL_last": lls[-1],
    }
    out.write(json.dumps(line) + "\n")
    out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
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
        "config": {"workload": workload_name(a)},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    out.write(json.dumps(line) + "\n")
    out.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
