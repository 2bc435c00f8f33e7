"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module (see oracle/mc_oracle.h).
Also holds the readers for the dump files written by oracle/ref_harness.c
(the unmodified reference driven in place) and for MCB1 genotype files.
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_HARNESS = os.path.join(HERE, "_ref", "ref_harness")
REF_BINARY = os.path.join(HERE, "_ref", "multiclust")


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class Options(C.Structure):
    _fields_ = [("admixture", C.c_int), ("eta_constrained", C.c_int),
                ("accel_scheme", C.c_int), ("do_projection", C.c_int),
                ("n_init_iter", C.c_int), ("max_iter", C.c_int),
                ("adjust_step", C.c_int), ("abs_error", C.c_double),
                ("rel_error", C.c_double), ("lower_bound", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.POINTER(Options)]
        L.orc_destroy.argtypes = [C.c_void_p]
        for name in ("orc_T", "orc_q", "orc_eta_len", "orc_n_parameters",
                     "orc_trace_len", "orc_em_step", "orc_em_2_steps",
                     "orc_accelerated_em_step"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_lower_bound.restype = C.c_double
        L.orc_lower_bound.argtypes = [C.c_void_p]
        L.orc_alloc_model.argtypes = [C.c_void_p, C.c_int]
        L.orc_seed.argtypes = [C.c_long]
        L.orc_initialize.argtypes = [C.c_void_p]
        L.orc_get_params.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.orc_set_params.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.orc_e_step.restype = C.c_double
        L.orc_e_step.argtypes = [C.c_void_p]
        L.orc_m_step.argtypes = [C.c_void_p]
        L.orc_log_likelihood.restype = C.c_double
        L.orc_log_likelihood.argtypes = [C.c_void_p, C.c_int]
        L.orc_project.argtypes = [dp, C.c_int, C.c_double]
        L.orc_em.argtypes = [C.c_void_p]
        L.orc_set_indices.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_get_state.argtypes = [C.c_void_p, dp, ip, ip, ip, ip, ip, ip]
        L.orc_get_posterior.argtypes = [C.c_void_p, dp]
        L.orc_get_trace.argtypes = [C.c_void_p, dp]
        L.orc_get_sums.argtypes = [C.c_void_p, dp, dp]
        L.orc_m_step_from_sums.argtypes = [C.c_void_p, dp, dp]
        L.orc_reset_trace.argtypes = [C.c_void_p]
        for name in ("orc_aic", "orc_bic"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p, C.c_double]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Fit:
    """One (data, options) pair; mirrors the reference's options/data/model."""

    def __init__(self, J, codes, admixture=1, eta_constrained=0, accel=0,
                 do_projection=1, n_init_iter=0, max_iter=0, adjust_step=0,
                 abs_error=1e-4, rel_error=0.0, lower_bound=1e-8):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        assert codes.ndim == 3
        self.I, self.L, self.P = codes.shape
        self.J = np.ascontiguousarray(J, dtype=np.int32)
        self.codes = codes
        self.opt = Options(admixture, eta_constrained, accel, do_projection,
                           n_init_iter, max_iter, adjust_step, abs_error,
                           rel_error, lower_bound)
        self.admixture = admixture
        self.h = lib().orc_create(self.I, self.L, self.P, self.J.ctypes.data,
                                  codes.ctypes.data, C.byref(self.opt))
        self.T = lib().orc_T(self.h)
        self.K = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    @property
    def lower_bound(self):
        return lib().orc_lower_bound(self.h)

    @property
    def q(self):
        return lib().orc_q(self.h)

    def alloc(self, K):
        self.K = K
        lib().orc_alloc_model(self.h, K)
        self.neta = lib().orc_eta_len(self.h)

    def n_parameters(self):
        return lib().orc_n_parameters(self.h)

    def initialize(self):
        lib().orc_initialize(self.h)

    def get_params(self, slot=0):
        eta = np.empty(self.neta)
        p = np.empty(self.K * self.T)
        lib().orc_get_params(self.h, slot, _dp(eta), _dp(p))
        return eta, p

    def set_params(self, slot, eta, p):
        eta = np.ascontiguousarray(eta, dtype=np.float64).ravel()
        p = np.ascontiguousarray(p, dtype=np.float64).ravel()
        assert eta.size == self.neta and p.size == self.K * self.T
        lib().orc_set_params(self.h, slot, _dp(eta), _dp(p))

    def set_indices(self, p, f, t):
        lib().orc_set_indices(self.h, p, f, t)

    def e_step(self):
        return lib().orc_e_step(self.h)

    def m_step(self):
        lib().orc_m_step(self.h)

    def em_step(self):
        return lib().orc_em_step(self.h)

    def em_2_steps(self):
        return lib().orc_em_2_steps(self.h)

    def accelerated_em_step(self):
        return lib().orc_accelerated_em_step(self.h)

    def log_likelihood(self, slot):
        return lib().orc_log_likelihood(self.h, slot)

    def em(self):
        lib().orc_em(self.h)

    def state(self):
        ll = C.c_double()
        v = [C.c_int() for _ in range(6)]
        lib().orc_get_state(self.h, C.byref(ll), *[C.byref(x) for x in v])
        names = ("n_iter", "converged", "stopped", "iter_stop", "pindex",
                 "aborted")
        out = {n: x.value for n, x in zip(names, v)}
        out["logL"] = ll.value
        return out

    def posterior(self):
        out = np.empty(self.I * self.K)
        lib().orc_get_posterior(self.h, _dp(out))
        return out.reshape(self.I, self.K)

    def sums(self):
        N = np.empty(self.K * self.T)
        S = np.empty(self.K)
        lib().orc_get_sums(self.h, _dp(N), _dp(S))
        return N, S

    def m_step_from_sums(self, N, S):
        N = np.ascontiguousarray(N, dtype=np.float64)
        S = np.ascontiguousarray(S, dtype=np.float64)
        lib().orc_m_step_from_sums(self.h, _dp(N), _dp(S))

    def trace(self):
        n = lib().orc_trace_len(self.h)
        out = np.empty(n)
        if n:
            lib().orc_get_trace(self.h, _dp(out))
        return out

    def reset_trace(self):
        lib().orc_reset_trace(self.h)

    def aic(self, max_logL):
        return lib().orc_aic(self.h, max_logL)

    def bic(self, max_logL):
        return lib().orc_bic(self.h, max_logL)


def seed(s):
    lib().orc_seed(int(s))


def project(x, floor):
    x = np.array(x, dtype=np.float64)
    lib().orc_project(_dp(x), x.size, float(floor))
    return x


# ---------------------------------------------------------------- file readers

def read_mcb(path):
    """MCB1 container (include/mc_format.h) -> dict of numpy arrays."""
    with open(path, "rb") as fp:
        raw = fp.read()
    assert raw[:4] == b"MCB1", path
    I, L, P, npops = struct.unpack_from("<4i", raw, 4)
    pos = 20
    J = np.frombuffer(raw, "<i4", L, pos).copy(); pos += 4 * L
    nreal = np.frombuffer(raw, "<i4", L, pos).copy(); pos += 4 * L
    nlab = int(nreal.sum())
    labels = np.frombuffer(raw, "<i4", nlab, pos).copy(); pos += 4 * nlab
    locale = np.frombuffer(raw, "<i4", I, pos).copy(); pos += 4 * I
    codes = np.frombuffer(raw, "u1", I * L * P, pos).copy().reshape(I, L, P)
    return dict(I=I, L=L, P=P, npops=npops, J=J, nreal=nreal, labels=labels,
                locale=locale, codes=codes)


def write_mcb(path, J, nreal, labels, locale, codes, npops=None):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    I, L, P = codes.shape
    locale = np.ascontiguousarray(locale, dtype="<i4")
    if npops is None:
        npops = int(locale.max()) + 1
    with open(path, "wb") as fp:
        fp.write(b"MCB1")
        fp.write(struct.pack("<4i", I, L, P, npops))
        fp.write(np.ascontiguousarray(J, dtype="<i4").tobytes())
        fp.write(np.ascontiguousarray(nreal, dtype="<i4").tobytes())
        fp.write(np.ascontiguousarray(labels, dtype="<i4").tobytes())
        fp.write(locale.tobytes())
        fp.write(codes.tobytes())


def read_state(path):
    """PREFIX.<tag>.bin written by ref_harness.c dump_state()."""
    with open(path, "rb") as fp:
        raw = fp.read()
    K, I, T, per_indiv, admixture, n_iter = struct.unpack_from("<6i", raw, 0)
    logL, = struct.unpack_from("<d", raw, 24)
    pos = 32
    neta = I * K if per_indiv else K
    eta = np.frombuffer(raw, "<f8", neta, pos).copy(); pos += 8 * neta
    p = np.frombuffer(raw, "<f8", K * T, pos).copy(); pos += 8 * K * T
    post = np.frombuffer(raw, "<f8", I * K, pos).copy().reshape(I, K)
    return dict(K=K, I=I, T=T, per_indiv=per_indiv, admixture=admixture,
                n_iter=n_iter, logL=logL, eta=eta, p=p, posterior=post)


def read_trace(path):
    """PREFIX.trace.txt -> {"ll": {(K, init): [..]}, "fit": {(K, init): {..}}}."""
    ll, fit, steps = {}, {}, {}
    with open(path) as fp:
        for line in fp:
            w = line.split()
            if not w:
                continue
            if w[0] == "ll":
                ll.setdefault((int(w[1]), int(w[2])), []).append(float(w[4]))
            elif w[0] == "fit":
                d = {}
                for kv in w[3:]:
                    k, v = kv.split("=")
                    d[k] = float(v) if k == "logL" else int(v)
                fit[(int(w[1]), int(w[2]))] = d
            elif w[0] == "step":
                d = {}
                for kv in w[4:]:
                    k, v = kv.split("=")
                    d[k] = float(v) if k == "logL" else int(v)
                steps.setdefault((int(w[1]), int(w[2])), []).append(d)
    return dict(ll=ll, fit=fit, steps=steps)


def have_ref():
    return os.path.exists(REF_HARNESS)


def run_ref(args, dump=None, mcb=None, steps=False, time_steps=0,
            parse_only=False, timeout=600, bootstrap_seed=None):
    """Run the reference harness; `args` is the multiclust command line."""
    cmd = [REF_HARNESS]
    if dump:
        cmd += ["--dump", dump]
    if mcb:
        cmd += ["--mcb", mcb]
    if steps:
        cmd += ["--steps"]
    if time_steps:
        cmd += ["--time", str(time_steps)]
    if parse_only:
        cmd += ["--parse-only"]
    if bootstrap_seed is not None:
        cmd += ["--bootstrap-sample", str(bootstrap_seed)]
    cmd += ["--"] + [str(a) for a in args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
