"""ctypes binding of libpodb200.so (the C-ABI declared in include/podb200.h).

The library is built in-tree (`pod_uqnn_b200/libpodb200.so`, see __graft_entry__.build()).
There is no CPU fallback: a missing library raises ImportError-like RuntimeError at first
use, a missing GPU makes Context() raise.
"""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpodb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "podb200.h")

PB_OK = 0
PB_ERR_CUDA, PB_ERR_SHAPE, PB_ERR_UNSUPPORTED_FIELD, PB_ERR_NOT_CONVERGED, PB_ERR_ARG, PB_ERR_NOMEM = \
    -1, -2, -3, -4, -5, -6
FIELD_IDS = {"shekel": 1, "ackley": 2, "burgers": 3}
POD_MODES = {"gram": 0, "gram2": 1, "accurate": 2, "tsqr": 3, "tsqr_full": 4}
NORM_IDS = {"none": 0, "meanstd": 1, "center": 2}
T_GRAM, T_EIG, T_BASIS, T_PROJECT, T_RECON, T_SNAPSHOTS, T_PREDICT, T_POD_TOTAL, T_PASS2, T_TN_KERNEL, T_QR, T_REFINE = range(12)

_i64, _int, _dbl, _vp = C.c_int64, C.c_int, C.c_double, C.c_void_p
_pi, _pd, _pvp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_void_p)

# name -> (restype, argtypes); must list every symbol include/podb200.h declares
SIGNATURES = {
    "pb_create": (_int, [_int, _pvp]),
    "pb_destroy": (_int, [_vp]),
    "pb_last_error": (C.c_char_p, [_vp]),
    "pb_set_stream": (_int, [_vp, _vp]),
    "pb_sync": (_int, [_vp]),
    "pb_device_info": (_int, [_vp, _pi, _pi, _pi, C.POINTER(C.c_size_t)]),
    "pb_last_ms": (_int, [_vp, _int, _pd]),
    "pb_kernel_launches": (_i64, [_vp]),
    "pb_timer_start": (_int, [_vp]),
    "pb_timer_stop": (_int, [_vp, _pd]),
    "pb_malloc": (_int, [_vp, C.c_size_t, _pvp]),
    "pb_free": (_int, [_vp, _vp]),
    "pb_memset": (_int, [_vp, _vp, _int, C.c_size_t]),
    "pb_upload": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "pb_download": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "pb_upload_async": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "pb_download_async": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "pb_download_forked": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "pb_slab": (_int, [_vp, C.c_size_t, _pvp]),
    "pb_upload_matrix": (_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "pb_host_alloc": (_int, [C.c_size_t, _pvp]),
    "pb_host_free": (_int, [_vp]),
    "pb_snapshots": (_int, [_vp, _int, _int, _i64, _vp, _i64, _int, _vp, _int, _vp, _i64, _i64, _vp, _i64]),
    "pb_snapshots_ackley_indexed": (_int, [_vp, _i64, _vp, _int, _vp, _vp, _int, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64]),
    "pb_snapshots_ackley_grouped": (_int, [_vp, _i64, _int, _vp, _vp, _int, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _i64]),
    "pb_dct_lowrank": (_int, [_vp, _i64, _i64, _int, _dbl, _i64, _i64, _vp, _i64]),
    "pb_gram": (_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "pb_gemm_tn": (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _vp]),
    "pb_gemm_nn": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64]),
    "pb_pod_from_gram": (_int, [_vp, _vp, _i64, _dbl, _int, _int, _pi, _pi]),
    "pb_pod_pass2_gram": (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "pb_pod_pass2_finish": (_int, [_vp, _vp, _dbl, _int, _pi]),
    "pb_pod_basis": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i64]),
    "pb_pod_refine_steps": (_int, [_vp, _pi]),
    "pb_pod_set_refine_rows": (_int, [_vp, _i64]),
    "pb_pod_refine_estimate": (_int, [_vp, _pd]),
    "pb_pod_refine_apply": (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "pb_pod_refine_finish": (_int, [_vp, _vp, _i64, _vp, _i64]),
    "pb_pod_sigma": (_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "pb_pod_right": (_int, [_vp, _vp, _i64]),
    "pb_tsqr_r": (_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "pb_pod_from_r": (_int, [_vp, _vp, _i64, _dbl, _int, _pi]),
    "pb_tsqr_factor": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "pb_pod_from_factor": (_int, [_vp, _vp, _i64, _i64, _i64, _dbl, _int, _pi]),
    "pb_tsqr_stats": (_int, [_vp, _vp, _vp, _vp]),
    "pb_pod": (_int, [_vp, _vp, _i64, _i64, _i64, _dbl, _int, _int, _pi]),
    "pb_gram_host": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "pb_pod_basis_full": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64]),
    "pb_pod_host": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _dbl, _int, _int, _pi]),
    "pb_fast_pod": (_int, [_vp, _vp, _i64, _int, _i64, _i64, _dbl, _dbl, _int, _pi, _pi]),
    "pb_fast_pod_basis": (_int, [_vp, _vp, _i64]),
    "pb_struct_to_matrix": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _i64]),
    "pb_matrix_to_struct": (_int, [_vp, _vp, _i64, _i64, _int, _i64, _vp]),
    "pb_project": (_int, [_vp, _vp, _i64, _int, _vp, _i64, _i64, _i64, _vp]),
    "pb_reconstruct": (_int, [_vp, _vp, _i64, _int, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp]),
    "pb_ensemble_predict": (_int, [_vp, _int, _int, _pi, _vp, _int, _vp, _vp, _dbl, _vp, _i64, _vp, _vp]),
    "pb_predict_U": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp, _i64, _int, C.c_uint64, _vp, _vp, _i64]),
    "pb_mixture_accumulate": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _int]),
    "pb_mixture_finalize": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "pb_ensemble_train": (_int, [_vp, _int, _int, _pi, _vp, _vp, _vp, _i64, _int, _dbl, _dbl, _int, _dbl, _dbl, _vp, _vp, _i64,
                          _vp, _int]),
    "pb_dct_basis": (_int, [_vp, _i64, _int, _i64, _i64, _vp, _i64]),
    "pb_residual_gram": (_int, [_vp, _vp, _i64, _int, _vp, _i64, _int, _i64, _vp, _vp]),
    "pb_probe_fp64_peak": (_int, [_vp, _int, _pd]),
    "pb_probe_hbm_copy": (_int, [_vp, C.c_size_t, _pd]),
    "pb_multi_create": (_int, [_int, _pi, _pvp]),
    "pb_multi_destroy": (_int, [_vp]),
    "pb_multi_last_error": (C.c_char_p, [_vp]),
    "pb_multi_size": (_int, [_vp]),
    "pb_multi_ctx": (_vp, [_vp, _int]),
    "pb_multi_row_range": (_int, [_vp, _i64, _int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "pb_multi_sync": (_int, [_vp]),
    "pb_multi_pod": (_int, [_vp, _pvp, C.POINTER(C.c_int64), _i64, C.POINTER(C.c_int64), _dbl, _int, _int, _pi]),
    "pb_multi_pod_host": (_int, [_vp, _vp, _i64, _i64, _i64, _dbl, _int, _int, _pi]),
    "pb_multi_fast_pod": (_int, [_vp, _pvp, C.POINTER(C.c_int64), _int, _i64, C.POINTER(C.c_int64), _dbl, _dbl, _int, _pi, _pi]),
    "pb_multi_basis": (_int, [_vp, _pvp, C.POINTER(C.c_int64)]),
    "pb_multi_basis_host": (_int, [_vp, _vp, _i64]),
    "pb_multi_project": (_int, [_vp, _pvp, C.POINTER(C.c_int64), _int, _pvp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _i64, _pvp]),
    "pb_multi_project_host": (_int, [_vp, _vp]),
    "pb_multi_sigma": (_int, [_vp, _vp, _i64, C.POINTER(C.c_int64)]),
    "pb_multi_all_reduce": (_int, [_vp, _pvp, _i64]),
    "pb_multi_timer_start": (_int, [_vp]),
    "pb_multi_timer_stop": (_int, [_vp, _pd]),
    "pb_probe_host_copy": (_int, [C.c_size_t, _int, _int, _pd]),
    "pb_flush_l2": (_int, [_vp]),
    "pb_gram_tile_shapes": (_int, [_int, _vp, _vp, _vp]),
    "pb_gram_profile": (_int, [_vp, _int, _int, _vp, _vp, _vp]),
}


def header_symbols():
    """Function names declared in include/podb200.h (used by the export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[A-Za-z0-9_]+)\s*\(", text)))


_lib = None


def lib():
    """Load libpodb200.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C pod_uqnn_b200/csrc).  pod_uqnn_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


class PodB200Error(RuntimeError):
    pass


_EXC = {PB_ERR_SHAPE: ValueError, PB_ERR_ARG: ValueError, PB_ERR_UNSUPPORTED_FIELD: NotImplementedError,
        PB_ERR_NOMEM: MemoryError}


class DeviceArray:
    """A float64 row-major matrix in HBM owned by a Context (freed on release / GC)."""

    def __init__(self, ctx, shape, ptr=None):
        self.ctx, self.shape = ctx, tuple(int(s) for s in shape)
        self.nbytes = 8 * int(np.prod(self.shape)) if len(self.shape) else 8
        self.owned = ptr is None
        if ptr is None:
            p = C.c_void_p()
            ctx.check(lib().pb_malloc(ctx.h, max(self.nbytes, 8), C.byref(p)))
            ptr = p.value
        self.ptr = ptr

    @property
    def ld(self):
        return self.shape[-1]

    def download(self):
        out = np.empty(self.shape, dtype=np.float64)
        self.ctx.check(lib().pb_download(self.ctx.h, out.ctypes.data, self.ptr, self.nbytes))
        return out

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.shape == self.shape, (host.shape, self.shape)
        if self.nbytes >= (8 << 20) and host.ndim >= 2:
            # large pageable arrays: the library's pinned staging ring (host threads) instead of the driver's bounce buffer
            cols = self.shape[-1]
            self.ctx.check(lib().pb_upload_matrix(self.ctx.h, host.ctypes.data, self.nbytes // (8 * cols), cols, cols, self.ptr))
            self.ctx.sync()
        else:
            self.ctx.check(lib().pb_upload(self.ctx.h, self.ptr, host.ctypes.data, self.nbytes))
        return self

    def release(self):
        if self.owned and self.ptr and self.ctx.h:
            lib().pb_free(self.ctx.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.release()
        except Exception:  # noqa: BLE001 - interpreter teardown
            pass


class Context:
    """One GPU, one stream (pb_ctx)."""

    def __init__(self, device=None, _borrowed=None):
        self._owner = _borrowed is None
        if _borrowed is not None:                      # a context owned by a MultiContext
            self.h, self.device = C.c_void_p(_borrowed), int(device)
            return
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        rc = lib().pb_create(int(device), C.byref(h))
        if rc != PB_OK:
            msg = lib().pb_last_error(None).decode()
            raise PodB200Error(f"pb_create({device}) failed ({rc}): {msg}")
        self.h, self.device = h, int(device)

    def check(self, rc):
        if rc != PB_OK:
            msg = lib().pb_last_error(self.h).decode()
            raise _EXC.get(rc, PodB200Error)(f"libpodb200 error {rc}: {msg}")

    def alloc(self, *shape):
        return DeviceArray(self, shape)

    def to_device(self, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        return DeviceArray(self, host.shape).upload(host)

    def to_device_i32(self, host):
        """int32 index array in HBM (returned as a DeviceArray of raw bytes; shape is informational)."""
        host = np.ascontiguousarray(host, dtype=np.int32)
        d = DeviceArray.__new__(DeviceArray)
        d.ctx, d.shape, d.nbytes, d.owned = self, host.shape, max(host.nbytes, 8), True
        p = C.c_void_p()
        self.check(lib().pb_malloc(self.h, d.nbytes, C.byref(p)))
        d.ptr = p.value
        self.check(lib().pb_upload(self.h, d.ptr, host.ctypes.data, host.nbytes))
        return d

    def wrap(self, ptr, shape):
        """Borrow a device pointer (e.g. torch.Tensor.data_ptr()) without taking ownership."""
        return DeviceArray(self, shape, ptr=int(ptr))

    def sync(self):
        self.check(lib().pb_sync(self.h))

    def set_stream(self, cuda_stream):
        self.check(lib().pb_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def ms(self, which):
        v = C.c_double()
        self.check(lib().pb_last_ms(self.h, which, C.byref(v)))
        return v.value

    def timer_start(self):
        self.check(lib().pb_timer_start(self.h))

    def timer_stop(self):
        v = C.c_double()
        self.check(lib().pb_timer_stop(self.h, C.byref(v)))
        return v.value

    def launches(self):
        return int(lib().pb_kernel_launches(self.h))

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self.check(lib().pb_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), mem_bytes=mem.value)

    def close(self):
        if self.h and self._owner:
            lib().pb_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


_default = None


def default_context():
    """Process-wide context on cuda:LOCAL_RANK (created on first use)."""
    global _default
    if _default is None or _default.h is None:
        _default = Context()
    return _default


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p) if p else None) for p in ptrs])


def _i64_array(vals):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


class MultiContext:
    """n_gpus GPUs driven from THIS process (pb_multi: one context + one host thread per GPU, the library's own
    peer-memory collectives over NVLink; no torch, no torchrun, no NCCL).  `device_ids` may repeat a device: logical
    ranks on one GPU (tests on a one-GPU box)."""

    def __init__(self, n_gpus, device_ids=None):
        ids = None if device_ids is None else (C.c_int * n_gpus)(*[int(d) for d in device_ids])
        h = C.c_void_p()
        rc = lib().pb_multi_create(int(n_gpus), ids, C.byref(h))
        if rc != PB_OK:
            raise PodB200Error(f"pb_multi_create({n_gpus}) failed ({rc}): {lib().pb_last_error(None).decode()}")
        self.h, self.n_gpus = h, int(n_gpus)
        devs = list(device_ids) if device_ids is not None else list(range(n_gpus))
        self.ctxs = [Context(devs[p], _borrowed=lib().pb_multi_ctx(h, p)) for p in range(n_gpus)]

    def check(self, rc):
        if rc != PB_OK:
            msg = lib().pb_multi_last_error(self.h).decode()
            raise _EXC.get(rc, PodB200Error)(f"libpodb200 (multi-GPU) error {rc}: {msg}")

    def row_range(self, n_rows, rank):
        lo, hi = C.c_int64(), C.c_int64()
        self.check(lib().pb_multi_row_range(self.h, int(n_rows), int(rank), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def sync(self):
        self.check(lib().pb_multi_sync(self.h))

    # ---- slabs resident in HBM (DeviceArray per rank, each owned by self.ctxs[p])
    def pod(self, slabs, eps=0., n_L=0, mode="tsqr"):
        """Row-sharded POD of per-GPU slabs; returns (V slabs, sigma, n_L)."""
        n_st = slabs[0].shape[1]
        rows = [s.shape[0] for s in slabs]
        nl = C.c_int()
        self.check(lib().pb_multi_pod(self.h, _ptr_array([s.ptr for s in slabs]), _i64_array(rows), n_st,
                                      _i64_array([s.ld for s in slabs]), float(eps), int(n_L), POD_MODES[mode], C.byref(nl)))
        V = [self.ctxs[p].alloc(max(rows[p], 1), nl.value) for p in range(self.n_gpus)]
        self.check(lib().pb_multi_basis(self.h, _ptr_array([v.ptr for v in V]), None))
        for p, v in enumerate(V):
            v.shape = (rows[p], nl.value)
        return V, self.sigma(), nl.value

    def fast_pod(self, slabs, n_t, n_s, eps, eps_init, mode="tsqr"):
        """Two-step POD (pod.py:51-68) of row-sharded slabs with sample-major columns; returns (V slabs, ranks, n_L)."""
        rows = [s.shape[0] for s in slabs]
        nl, ranks = C.c_int(), (C.c_int * n_s)()
        self.check(lib().pb_multi_fast_pod(self.h, _ptr_array([s.ptr for s in slabs]), _i64_array(rows), int(n_t), int(n_s),
                                           _i64_array([s.ld for s in slabs]), float(eps), float(eps_init), POD_MODES[mode],
                                           C.byref(nl), ranks))
        V = [self.ctxs[p].alloc(max(rows[p], 1), nl.value) for p in range(self.n_gpus)]
        self.check(lib().pb_multi_basis(self.h, _ptr_array([v.ptr for v in V]), None))
        for p, v in enumerate(V):
            v.shape = (rows[p], nl.value)
        return V, np.array(ranks[:], dtype=np.int64), nl.value

    def project(self, V, slabs):
        """v (n x n_L) = sum_p (V_p^T U_p)^T; returned as the host copy of rank 0's (replicated) result."""
        n, n_L = slabs[0].shape[1], V[0].shape[1]
        out = [self.ctxs[p].alloc(n, n_L) for p in range(self.n_gpus)]
        self.check(lib().pb_multi_project(self.h, _ptr_array([v.ptr for v in V]), None, n_L, _ptr_array([s.ptr for s in slabs]),
                                          _i64_array([s.ld for s in slabs]), _i64_array([s.shape[0] for s in slabs]), n,
                                          _ptr_array([o.ptr for o in out])))
        return out

    def sigma(self):
        cnt = C.c_int64()
        self.check(lib().pb_multi_sigma(self.h, None, 0, C.byref(cnt)))
        sig = np.zeros(cnt.value)
        self.check(lib().pb_multi_sigma(self.h, sig.ctypes.data, cnt.value, C.byref(cnt)))
        return sig

    # ---- host matrices (what the reference's calls are handed)
    def pod_host(self, U, eps=0., n_L=0, mode="tsqr"):
        U = np.ascontiguousarray(U, dtype=np.float64)
        n_h, n_st = U.shape
        nl = C.c_int()
        self.check(lib().pb_multi_pod_host(self.h, U.ctypes.data, n_h, n_st, n_st, float(eps), int(n_L), POD_MODES[mode], C.byref(nl)))
        V = np.empty((n_h, nl.value))
        self.check(lib().pb_multi_basis_host(self.h, V.ctypes.data, nl.value))
        self._held = (n_h, n_st, nl.value)
        return V, self.sigma()

    def project_held(self):
        """v of the matrix / basis the last pod_host left in HBM."""
        n_h, n_st, n_L = self._held
        v = np.empty((n_st, n_L))
        self.check(lib().pb_multi_project_host(self.h, v.ctypes.data))
        return v

    def timer_start(self):
        self.check(lib().pb_multi_timer_start(self.h))

    def timer_stop(self):
        v = C.c_double()
        self.check(lib().pb_multi_timer_stop(self.h, C.byref(v)))
        return v.value

    def close(self):
        if self.h:
            for c in self.ctxs:
                c.h = None
            lib().pb_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


_multi = {}


def multi_context(n_gpus):
    """Process-wide MultiContext over GPUs 0..n_gpus-1 (created on first use)."""
    m = _multi.get(n_gpus)
    if m is None or m.h is None:
        m = _multi[n_gpus] = MultiContext(n_gpus)
    return m
