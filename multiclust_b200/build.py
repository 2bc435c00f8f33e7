"""Build the native pieces of multiclust_b200 in-tree.

  libmc_cuda.so   hand-written sm_100a kernels + the C ABI (include/mc_cuda.h)
  host/multiclust the drop-in command line (C17) linked against libmc_cuda.so
  host/mc_gen     synthetic workload generator

nvcc cross-compiles for sm_100a without a GPU; the resulting files travel to
the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
INC = os.path.join(ROOT, "include")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB = os.path.join(PKG, "libmc_cuda.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-diag-suppress", "128",
              "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-Wall", "-I" + INC, "-I" + CSRC]
HOST_CFLAGS = ["-std=c17", "-O2", "-Wall", "-Wextra", "-D_GNU_SOURCE",
               "-ffp-contract=off", "-I" + INC, "-I" + HOST]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


OBJ = os.path.join(CSRC, "_obj")


def build_cuda(force=False, extra=()):
    """libmc_cuda.so from mc_cuda.cu and the mc_inst_*.cu instantiation units,
    compiled in parallel (one nvcc per unit) and linked"""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    units = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and f != "mc_comm.cu")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
            if f.endswith((".cu", ".cuh", ".h")) and f != "mc_comm.cu"]
    deps += [os.path.join(INC, f) for f in os.listdir(INC)]
    if not force and _newer(LIB, deps):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    procs = []
    for u in units:
        obj = os.path.join(OBJ, u[:-3] + ".o")
        cmd = [nvcc] + flags + list(extra) + ["-c", "-o", obj, os.path.join(CSRC, u)]
        print("+", " ".join(cmd), flush=True)
        procs.append((u, obj, subprocess.Popen(cmd)))
    objs = []
    for u, obj, pr in procs:
        if pr.wait() != 0:
            for _, _, other in procs:
                if other.poll() is None:
                    other.kill()
            raise subprocess.CalledProcessError(pr.returncode, "nvcc -c " + u)
        objs.append(obj)
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    return LIB


COMM = os.path.join(PKG, "libmc_comm.so")


def build_comm(force=False):
    """libmc_comm.so: the NCCL exchange used by the C host with --gpus N"""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = os.path.join(CSRC, "mc_comm.cu")
    deps = [src, LIB] + [os.path.join(INC, f) for f in os.listdir(INC)]
    if not force and _newer(COMM, deps):
        return COMM
    _run([nvcc] + NVCC_FLAGS + ["-o", COMM, src, "-L" + PKG, "-lmc_cuda",
                                "-Xlinker", "-rpath," + PKG, "-lnccl"])
    return COMM


def build_host(force=False):
    cc = shutil.which("gcc") or "gcc"
    out = []
    gen = os.path.join(HOST, "mc_gen")
    gsrc = [os.path.join(HOST, "mc_gen.c")] + [os.path.join(INC, f) for f in os.listdir(INC)]
    if force or not _newer(gen, gsrc):
        _run([cc] + HOST_CFLAGS + ["-o", gen, os.path.join(HOST, "mc_gen.c")])
    out.append(gen)
    cli_src = sorted(os.path.join(HOST, f) for f in os.listdir(HOST)
                     if f.endswith(".c") and f != "mc_gen.c")
    if cli_src:
        exe = os.path.join(HOST, "multiclust")
        deps = cli_src + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")]
        deps += [os.path.join(INC, f) for f in os.listdir(INC)] + [LIB, COMM]
        if force or not _newer(exe, deps):
            _run([cc] + HOST_CFLAGS + ["-o", exe] + cli_src
                 + ["-L" + PKG, "-lmc_comm", "-lmc_cuda", "-Wl,-rpath," + PKG, "-lm", "-lpthread"])
        out.append(exe)
    return out


def build_all(force=False):
    build_cuda(force)
    build_comm(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
