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
    "pb_slab": (_int, [_vp, C.c_size_t, _pvp]),
    "pb_upload_matrix": (_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "pb_host_alloc": (_int, [C.c_size_t, _pvp]),
    "pb_host_free": (_int, [_vp]),
    "pb_snapshots": (_int, [_vp, _int, _int, _i64, _vp, _i64, _int, _vp, _int, _vp, _i64, _i64, _vp, _i64]),
    "pb_snapshots_ackley_indexed": (_int, [_vp, _i64, _vp, _int, _vp, _vp, _int, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64]),
    "pb_dct_lowrank": (_int, [_vp, _i64, _i64, _int, _dbl, _i64, _i64, _vp, _i64]),
    "pb_gram": (_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "pb_gemm_tn": (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _vp]),
    "pb_gemm_nn": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64]),
    "pb_pod_from_gram": (_int, [_vp, _vp, _i64, _dbl, _int, _int, _pi, _pi]),
    "pb_pod_pass2_gram": (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "pb_pod_pass2_finish": (_int, [_vp, _vp, _dbl, _int, _pi]),
    "pb_pod_basis": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i64]),
    "pb_pod_refine_steps": (_int, [_vp, _pi]),
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
    "pb_ensemble_train": (_int, [_vp, _int, _int, _pi, _vp, _vp, _vp, _i64, _int, _dbl, _dbl, _int, _dbl, _dbl, _vp, _vp, _i64,
                          _vp, _int]),
    "pb_dct_basis": (_int, [_vp, _i64, _int, _i64, _i64, _vp, _i64]),
    "pb_residual_gram": (_int, [_vp, _vp, _i64, _int, _vp, _i64, _int, _i64, _vp, _vp]),
    "pb_probe_fp64_peak": (_int, [_vp, _int, _pd]),
    "pb_probe_hbm_copy": (_int, [_vp, C.c_size_t, _pd]),
    "pb_probe_host_copy": (_int, [C.c_size_t, _int, _int, _pd]),
    "pb_flush_l2": (_int, [_vp]),
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

    def __init__(self, device=None):
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
        if self.h:
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
