"""ctypes view of include/mc_cuda.h (one method per entry point).

No compute lives here.  If libmc_cuda.so is missing the import of the library
fails loudly; there is no fallback path of any kind.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))

# every symbol include/mc_cuda.h declares (tests check the library exports all)
SYMBOLS = [
    "mc_last_error", "mc_abi_version", "mc_create", "mc_destroy",
    "mc_set_stream", "mc_sync", "mc_ctx_device", "mc_ctx_stream", "mc_set_data", "mc_set_data_synth",
    "mc_get_dims", "mc_get_J", "mc_get_codes", "mc_alloc_model", "mc_eta_len",
    "mc_set_params", "mc_get_params", "mc_init_admixture", "mc_init_admixture_local", "mc_init_admixture_rand", "mc_init_admixture_rand_local", "mc_init_mixture", "mc_init_mixture_local", "mc_init_mixture_finish", "mc_save_mle", "mc_bootstrap_data", "mc_restore_data", "mc_em_step", "mc_loglik", "mc_read_ll",
    "mc_get_posterior", "mc_partition", "mc_locale_sums", "mc_delta", "mc_step_dots",
    "mc_qn_dots", "mc_accel_update", "mc_qn_update", "mc_project",
    "mc_copy_slot", "mc_em_step_local", "mc_exchange_buffer",
    "mc_em_step_finish", "mc_exchange_sum", "mc_exchange_sum_slice", "mc_get_plan", "mc_launch_count",
    "mc_profile_enable", "mc_profile_read", "mc_set_option",
]


class McError(RuntimeError):
    pass


class SynthParams(C.Structure):
    """mcs_params of include/mc_synth.h"""
    _fields_ = [("seed", C.c_uint64), ("K", C.c_int32), ("jmax", C.c_int32),
                ("miss_bp", C.c_int32), ("ploidy", C.c_int32)]


class PlanInfo(C.Structure):
    _fields_ = [("K", C.c_int32), ("k_split", C.c_int32),
                ("k_per_lane", C.c_int32), ("loci_per_warp", C.c_int32),
                ("warps", C.c_int32), ("groups", C.c_int32),
                ("n_tiles", C.c_int32), ("n_chunks", C.c_int32),
                ("n_units", C.c_int32), ("grid", C.c_int32),
                ("block", C.c_int32), ("indiv_per_block", C.c_int32),
                ("ploidy_padded", C.c_int32), ("two_pass", C.c_int32),
                ("reserved", C.c_int32), ("smem_bytes", C.c_int64),
                ("algorithmic_bytes_em", C.c_int64),
                ("algorithmic_bytes_ll", C.c_int64)]


def lib_path():
    return os.path.join(PKG, "libmc_cuda.so")


_lib = None


def load_library():
    """dlopen libmc_cuda.so; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise McError(
            "%s is missing: build it with `python -m multiclust_b200.build` "
            "(there is no CPU fallback for the EM path)" % path)
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    L.mc_last_error.restype = C.c_char_p
    L.mc_last_error.argtypes = [vp]
    L.mc_abi_version.restype = C.c_int
    L.mc_create.argtypes = [C.POINTER(vp), C.c_int]
    L.mc_destroy.argtypes = [vp]
    L.mc_destroy.restype = None
    L.mc_set_stream.argtypes = [vp, vp]
    L.mc_sync.argtypes = [vp]
    L.mc_ctx_device.argtypes = [vp]
    L.mc_ctx_stream.argtypes = [vp]
    L.mc_ctx_stream.restype = vp
    L.mc_set_data.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, vp, vp]
    L.mc_set_data_synth.argtypes = [vp, C.c_int64, C.c_int32,
                                    C.POINTER(SynthParams), C.c_int64]
    L.mc_get_dims.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                              C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.mc_get_J.argtypes = [vp, vp]
    L.mc_get_codes.argtypes = [vp, vp]
    L.mc_alloc_model.argtypes = [vp, C.c_int32, C.c_int, C.c_int, C.c_int,
                                 C.c_double, C.c_double, C.c_int]
    L.mc_eta_len.argtypes = [vp, C.POINTER(C.c_int64)]
    L.mc_set_params.argtypes = [vp, C.c_int, vp, vp]
    L.mc_get_params.argtypes = [vp, C.c_int, vp, vp]
    L.mc_init_admixture.argtypes = [vp, C.c_int, vp]
    L.mc_init_admixture_local.argtypes = [vp, C.c_int, vp]
    L.mc_init_admixture_rand.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int64]
    L.mc_init_admixture_rand_local.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int64]
    L.mc_init_mixture.argtypes = [vp, C.c_int, vp, vp]
    L.mc_init_mixture_local.argtypes = [vp, vp, vp]
    L.mc_init_mixture_finish.argtypes = [vp, C.c_int, C.c_int64]
    L.mc_em_step.argtypes = [vp, C.c_int, C.c_int, dp]
    L.mc_loglik.argtypes = [vp, C.c_int, dp]
    L.mc_read_ll.argtypes = [vp, dp]
    L.mc_get_posterior.argtypes = [vp, vp]
    L.mc_partition.argtypes = [vp, vp, vp]
    L.mc_locale_sums.argtypes = [vp, vp, C.c_int32, vp]
    L.mc_delta.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.mc_step_dots.argtypes = [vp, C.c_int, dp, dp]
    L.mc_qn_dots.argtypes = [vp, C.c_int, C.c_int, dp, dp]
    L.mc_accel_update.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_double]
    L.mc_qn_update.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp]
    L.mc_project.argtypes = [vp, C.c_int]
    L.mc_copy_slot.argtypes = [vp, C.c_int, C.c_int]
    L.mc_em_step_local.argtypes = [vp, C.c_int, C.c_int]
    L.mc_exchange_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mc_em_step_finish.argtypes = [vp, C.c_int, dp]
    L.mc_exchange_sum.argtypes = [vp, vp, C.c_int]
    L.mc_exchange_sum_slice.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int64]
    L.mc_get_plan.argtypes = [vp, C.POINTER(PlanInfo)]
    L.mc_launch_count.restype = C.c_int64
    L.mc_launch_count.argtypes = [vp]
    L.mc_profile_enable.argtypes = [vp, C.c_int]
    L.mc_profile_read.argtypes = [vp, C.POINTER(C.c_int64), dp]
    L.mc_set_option.argtypes = [vp, C.c_int, C.c_int]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU context; thin wrapper, same call order as the C host uses."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.mc_create(C.byref(h), int(device))
        if rc:
            raise McError("mc_create: " + self.lib.mc_last_error(None).decode())
        self.h = h
        self.K = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.mc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise McError("%s failed (%d): %s" % (
                what, rc, self.lib.mc_last_error(self.h).decode()))

    OPT_KERNEL, OPT_TIMING, OPT_GRAPH = 1, 2, 3
    KERNEL_AUTO, KERNEL_TILE, KERNEL_ADMIX3, KERNEL_DENSE, KERNEL_DIGIT = 0, 1, 2, 3, 4

    def set_option(self, option, value):
        self._ck(self.lib.mc_set_option(self.h, int(option), int(value)), "mc_set_option")

    # -- data
    def set_stream(self, cuda_stream):
        self._ck(self.lib.mc_set_stream(self.h, C.c_void_p(cuda_stream or 0)), "mc_set_stream")

    def sync(self):
        self._ck(self.lib.mc_sync(self.h), "mc_sync")

    def set_data(self, J, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        J = np.ascontiguousarray(J, dtype=np.int32)
        I, L, P = codes.shape
        assert J.size == L
        self._ck(self.lib.mc_set_data(self.h, I, L, P, _ptr(J), _ptr(codes)), "mc_set_data")
        self._dims()

    def set_data_synth(self, I, L, params, i_first=0):
        self._ck(self.lib.mc_set_data_synth(self.h, I, L, C.byref(params), i_first),
                 "mc_set_data_synth")
        self._dims()

    def _dims(self):
        I, T = C.c_int64(), C.c_int64()
        L, P = C.c_int32(), C.c_int32()
        self._ck(self.lib.mc_get_dims(self.h, C.byref(I), C.byref(L), C.byref(P), C.byref(T)),
                 "mc_get_dims")
        self.I, self.L, self.P, self.T = I.value, L.value, P.value, T.value

    def get_J(self):
        J = np.empty(self.L, dtype=np.int32)
        self._ck(self.lib.mc_get_J(self.h, _ptr(J)), "mc_get_J")
        return J

    def get_codes(self):
        out = np.empty((self.I, self.L, self.P), dtype=np.uint8)
        self._ck(self.lib.mc_get_codes(self.h, _ptr(out)), "mc_get_codes")
        return out

    # -- model
    def alloc_model(self, K, admixture=1, eta_constrained=0, q=0, eta_lb=1e-8,
                    p_lb=1e-8, do_projection=1):
        self._ck(self.lib.mc_alloc_model(self.h, K, admixture, eta_constrained, q,
                                         eta_lb, p_lb, do_projection), "mc_alloc_model")
        self.K = K
        n = C.c_int64()
        self._ck(self.lib.mc_eta_len(self.h, C.byref(n)), "mc_eta_len")
        self.neta = n.value

    def set_params(self, slot, eta, p):
        eta = np.ascontiguousarray(eta, dtype=np.float64).ravel()
        p = np.ascontiguousarray(p, dtype=np.float64).ravel()
        assert eta.size == self.neta and p.size == self.K * self.T
        self._ck(self.lib.mc_set_params(self.h, slot, _ptr(eta), _ptr(p)), "mc_set_params")

    def get_params(self, slot):
        eta = np.empty(self.neta)
        p = np.empty(self.K * self.T)
        self._ck(self.lib.mc_get_params(self.h, slot, _ptr(eta), _ptr(p)), "mc_get_params")
        return eta, p

    def init_admixture(self, slot, z):
        z = np.ascontiguousarray(z, dtype=np.uint8)
        assert z.size == self.I * self.L * self.P
        self._ck(self.lib.mc_init_admixture(self.h, slot, _ptr(z)), "mc_init_admixture")

    def init_admixture_rand(self, slot, hist, block_draws):
        """draws made on the device from per-block generator histories [n_blocks][31]"""
        hist = np.ascontiguousarray(hist, dtype=np.uint32)
        assert hist.ndim == 2 and hist.shape[1] == 31
        self._ck(self.lib.mc_init_admixture_rand(self.h, slot, _ptr(hist), hist.shape[0],
                                                 int(block_draws)), "mc_init_admixture_rand")

    # -- parametric bootstrap
    def save_mle(self, slot):
        """keep the parameters of `slot` as the H0 estimates the samples are drawn from"""
        self._ck(self.lib.mc_save_mle(self.h, slot), "mc_save_mle")

    def bootstrap_data(self, hist, block_draws):
        """replace the data by one parametric bootstrap sample (frees the model)"""
        hist = np.ascontiguousarray(hist, dtype=np.uint32)
        assert hist.ndim == 2 and hist.shape[1] == 31
        self._ck(self.lib.mc_bootstrap_data(self.h, _ptr(hist), hist.shape[0],
                                            int(block_draws)), "mc_bootstrap_data")

    def restore_data(self):
        self._ck(self.lib.mc_restore_data(self.h), "mc_restore_data")

    def init_mixture(self, slot, center_idx, center_codes):
        """mixture initialiser with the distance work on the device; the host
        draws the centres"""
        ci = np.ascontiguousarray(center_idx, dtype=np.int32)
        cc = np.ascontiguousarray(center_codes, dtype=np.uint8)
        assert ci.size == self.K and cc.size == self.K * self.L * self.P
        self._ck(self.lib.mc_init_mixture(self.h, slot, _ptr(ci), _ptr(cc)), "mc_init_mixture")

    # -- hot path
    def em_step(self, frm=0, to=0):
        ll = C.c_double()
        self._ck(self.lib.mc_em_step(self.h, frm, to, C.byref(ll)), "mc_em_step")
        return ll.value

    def em_step_local(self, frm=0, to=0):
        self._ck(self.lib.mc_em_step_local(self.h, frm, to), "mc_em_step_local")

    def exchange_buffer(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.mc_exchange_buffer(self.h, C.byref(p), C.byref(n)), "mc_exchange_buffer")
        return p.value, n.value

    def em_step_finish(self, to=0, want_ll=True):
        ll = C.c_double()
        self._ck(self.lib.mc_em_step_finish(self.h, to, C.byref(ll) if want_ll else None),
                 "mc_em_step_finish")
        return ll.value

    # the shard interface of multiclust_b200/sharding.py
    def exchange_len(self):
        return self.exchange_buffer()[1]

    def exchange_tensor(self):
        """the exchange buffer (logical length + 64 doubles of zero padding)
        viewed as a torch CUDA tensor (no copy)"""
        import torch
        ptr, n = self.exchange_buffer()
        if getattr(self, "_xt", None) is None or self._xt_key != (ptr, n):
            class _Arr:
                __cuda_array_interface__ = {"shape": (n + 64,), "typestr": "<f8",
                                            "data": (ptr, False), "version": 3}
            self._xt = torch.as_tensor(_Arr(), device="cuda")
            self._xt_key = (ptr, n)
        return self._xt

    def sum_gathered(self, gathered, world):
        self.exchange_sum(gathered.data_ptr(), world)

    def sum_slices(self, parts, world, first, count):
        self._ck(self.lib.mc_exchange_sum_slice(self.h, C.c_void_p(parts.data_ptr()), world,
                                                first, count), "mc_exchange_sum_slice")

    def exchange_sum(self, gathered_ptr, n_ranks):
        self._ck(self.lib.mc_exchange_sum(self.h, C.c_void_p(gathered_ptr), n_ranks),
                 "mc_exchange_sum")

    def loglik(self, slot=0):
        ll = C.c_double()
        self._ck(self.lib.mc_loglik(self.h, slot, C.byref(ll)), "mc_loglik")
        return ll.value

    def posterior(self):
        out = np.empty((self.I, self.K))
        self._ck(self.lib.mc_get_posterior(self.h, _ptr(out)), "mc_get_posterior")
        return out

    def partition(self):
        ik = np.empty(self.I, dtype=np.int32)
        cnt = np.empty(self.K, dtype=np.int32)
        self._ck(self.lib.mc_partition(self.h, _ptr(ik), _ptr(cnt)), "mc_partition")
        return ik, cnt

    def locale_sums(self, locale, n_locales):
        locale = np.ascontiguousarray(locale, dtype=np.int32)
        assert locale.size == self.I
        out = np.empty((n_locales, self.K))
        self._ck(self.lib.mc_locale_sums(self.h, _ptr(locale), n_locales, _ptr(out)),
                 "mc_locale_sums")
        return out

    def delta(self, which, pair, slot_t, slot_f):
        self._ck(self.lib.mc_delta(self.h, which, pair, slot_t, slot_f), "mc_delta")

    def step_dots(self, pair):
        e = (C.c_double * 3)()
        p = (C.c_double * 3)()
        self._ck(self.lib.mc_step_dots(self.h, pair, e, p), "mc_step_dots")
        return np.array(e[:]), np.array(p[:])

    def qn_dots(self, q1, q2):
        e = (C.c_double * 2)()
        p = (C.c_double * 2)()
        self._ck(self.lib.mc_qn_dots(self.h, q1, q2, e, p), "mc_qn_dots")
        return np.array(e[:]), np.array(p[:])

    def accel_update(self, qn1, slot_t, slot_p, pair, s):
        self._ck(self.lib.mc_accel_update(self.h, int(qn1), slot_t, slot_p, pair, float(s)),
                 "mc_accel_update")

    def qn_update(self, slot_t, slot_p, uindex, delta_index, Ainv, cutu):
        Ainv = np.ascontiguousarray(Ainv, dtype=np.float64).ravel()
        cutu = np.ascontiguousarray(cutu, dtype=np.float64).ravel()
        self._ck(self.lib.mc_qn_update(self.h, slot_t, slot_p, uindex, delta_index,
                                       Ainv.ctypes.data_as(C.POINTER(C.c_double)),
                                       cutu.ctypes.data_as(C.POINTER(C.c_double))),
                 "mc_qn_update")

    def project(self, slot):
        self._ck(self.lib.mc_project(self.h, slot), "mc_project")

    def copy_slot(self, dst, src):
        self._ck(self.lib.mc_copy_slot(self.h, dst, src), "mc_copy_slot")

    # -- introspection
    def plan(self):
        pi = PlanInfo()
        self._ck(self.lib.mc_get_plan(self.h, C.byref(pi)), "mc_get_plan")
        return {f: getattr(pi, f) for f, _ in PlanInfo._fields_}

    def launch_count(self):
        return self.lib.mc_launch_count(self.h)

    def profile_enable(self, on=True):
        self._ck(self.lib.mc_profile_enable(self.h, int(on)), "mc_profile_enable")

    def profile_read(self):
        n, ms = C.c_int64(), C.c_double()
        self._ck(self.lib.mc_profile_read(self.h, C.byref(n), C.byref(ms)), "mc_profile_read")
        return n.value, ms.value
