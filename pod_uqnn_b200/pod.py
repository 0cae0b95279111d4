"""POD -- same names and argument order as the reference's poduqnn/pod.py
(perform_pod :7-47, perform_fast_pod :51-68), computed on the GPU by libpodb200.

Algorithm (csrc/gemm.cu, csrc/eig.cu, csrc/qr.cu).  Gram modes -- method of snapshots: G = A^T A on DMMA
(A = U, or U^T when n_h < n_st), pivoted Cholesky of G, one-sided block Jacobi on the factor, energy
rank rule (pod.py:22-32), V = U Z diag(1/sigma).  `mode`:
  "gram"      one Gram pass (fast; subspace angle limited by eps_mach * lambda_1 / gap)
  "gram2"     + rotate the numerical-rank block and repeat the Gram on it
  "accurate"  rotate all columns and repeat (one-sided Jacobi SVD accuracy)
  "tsqr"      blocked Householder QR of A itself (never A^T A), rank-revealing pivoted QR of R, Jacobi SVD of
              the factor: the accuracy class of the reference's LAPACK SVD (pod.py:16) at half the cost of
              "accurate" -- the default
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DeviceArray, POD_MODES

DEFAULT_MODE = "tsqr"


def pod_device(ctx, U, eps=0., n_L=0, mode=DEFAULT_MODE):
    """POD of a device matrix.  Returns (V_dev (n_h, n_L), sigma host (min(n_h, n_st),))."""
    n_h, n_st = U.shape
    nl = C.c_int()
    ctx.check(_lib.lib().pb_pod(ctx.h, U.ptr, n_h, n_st, U.ld, float(eps), int(n_L), POD_MODES[mode], C.byref(nl)))
    V = ctx.alloc(n_h, nl.value)
    ctx.check(_lib.lib().pb_pod_basis_full(ctx.h, U.ptr, n_h, n_st, U.ld, V.ptr, V.ld))
    return V, pod_sigma(ctx)


def pod_sigma(ctx):
    cnt = C.c_int64()
    ctx.check(_lib.lib().pb_pod_sigma(ctx.h, None, 0, C.byref(cnt)))
    sig = np.zeros(cnt.value)
    ctx.check(_lib.lib().pb_pod_sigma(ctx.h, sig.ctypes.data, cnt.value, C.byref(cnt)))
    return sig


def perform_pod(U, eps=0., n_L=0, verbose=True, *, mode=DEFAULT_MODE, ctx=None, return_sigma=False, n_gpus=1):
    """Reference signature (pod.py:7).  numpy in -> numpy V (n_h, n_L), C-contiguous float64.

    A DeviceArray input stays on the device and returns a DeviceArray.
    n_gpus > 1: the rows of U are sharded over GPUs 0..n_gpus-1 of this box from THIS process (pb_multi_pod_host:
    concurrent slab uploads, the library's own NVLink collectives); tall matrices only (n_h >= n_st), wide or small
    ones run on one GPU.
    """
    on_device = isinstance(U, DeviceArray)
    if verbose:
        print("Contructing the reduced bases V")     # pod.py:38 (sic)
    if n_gpus > 1 and not on_device and U.shape[0] >= max(U.shape[1], 4096 * n_gpus):
        V, sig = _lib.multi_context(n_gpus).pod_host(U, eps, n_L, mode)
        return (V, sig) if return_sigma else V
    ctx = ctx or _lib.default_context()
    if on_device:
        Vd, sig = pod_device(ctx, U, eps, n_L, mode)
        return (Vd, sig) if return_sigma else Vd
    Vd, sig = pod_host(ctx, U, eps, n_L, mode)
    V = Vd.download()
    return (V, sig) if return_sigma else V


def _slab(ctx, n_h, n_st):
    """The context-owned device slab (pb_slab: grown on demand, reused across calls) viewed as an (n_h, n_st) matrix."""
    p = C.c_void_p()
    ctx.check(_lib.lib().pb_slab(ctx.h, 8 * n_h * n_st, C.byref(p)))
    return ctx.wrap(p.value, (n_h, n_st))


def pod_host(ctx, U, eps=0., n_L=0, mode=DEFAULT_MODE):
    """POD of a host matrix (an ordinary numpy array): chunked upload overlapped with the Gram (pb_pod_host; pageable
    memory goes through the library's pinned staging ring).  The uploaded matrix stays in the context's slab until the
    next host-facing call: `project_to_v(V, U, reuse_slab=True)` can project onto it without a second upload."""
    U = np.ascontiguousarray(U, dtype=np.float64)
    n_h, n_st = U.shape
    Ud = _slab(ctx, n_h, n_st)
    ctx._slab_shape = None
    nl = C.c_int()
    ctx.check(_lib.lib().pb_pod_host(ctx.h, U.ctypes.data, n_h, n_st, n_st, Ud.ptr, float(eps), int(n_L),
                                     POD_MODES[mode], C.byref(nl)))
    V = ctx.alloc(n_h, nl.value)
    ctx.check(_lib.lib().pb_pod_basis_full(ctx.h, Ud.ptr, n_h, n_st, Ud.ld, V.ptr, V.ld))
    ctx._slab_shape = (n_h, n_st)
    ctx._slab_V = V                                   # the basis of the slab's matrix, still in HBM
    return V, pod_sigma(ctx)


def project_to_v(V, U, *, ctx=None, reuse_slab=False, n_gpus=1):
    """v = (V.T.dot(U)).T (podnnmodel.py:435-437) as a module-level call: numpy in, numpy (n, n_L) out.

    reuse_slab=True: `U` is the array the last perform_pod on this context was given and `V` what it returned; both
    are still resident in HBM and are not uploaded again (the caller vouches that U has not been modified since).
    n_gpus > 1: row-sharded over the GPUs of this process' MultiContext (the partial products are summed over NVLink)."""
    if n_gpus > 1 and U.shape[0] >= max(U.shape[1], 4096 * n_gpus):
        m = _lib.multi_context(n_gpus)
        if reuse_slab:
            if getattr(m, "_held", None) != (U.shape[0], U.shape[1], V.shape[1]):
                raise ValueError("project_to_v(reuse_slab=True): the multi-GPU context does not hold this matrix")
            return m.project_held()
        m._held = None
        rng = [m.row_range(U.shape[0], p) for p in range(n_gpus)]
        Us = [m.ctxs[p].to_device(U[lo:hi]) for p, (lo, hi) in enumerate(rng)]
        Vs = [m.ctxs[p].to_device(V[lo:hi]) for p, (lo, hi) in enumerate(rng)]
        return m.project(Vs, Us)[0].download()
    ctx = ctx or _lib.default_context()
    L = _lib.lib()
    n_h, n = U.shape
    n_L = V.shape[1]
    if reuse_slab:
        if getattr(ctx, "_slab_shape", None) != (n_h, n) or ctx._slab_V.shape != (n_h, n_L):
            raise ValueError("project_to_v(reuse_slab=True): the context does not hold this matrix (call perform_pod on it first)")
        Ud, Vd = _slab(ctx, n_h, n), ctx._slab_V
    else:
        Ud = U if isinstance(U, DeviceArray) else _slab_upload(ctx, U)
        Vd = V if isinstance(V, DeviceArray) else ctx.to_device(V)
    v = ctx.alloc(n, n_L)
    ctx.check(L.pb_project(ctx.h, Vd.ptr, Vd.ld, n_L, Ud.ptr, Ud.ld, n_h, n, v.ptr))
    return v.download()


def _slab_upload(ctx, U):
    """Host matrix -> the context's slab (staged upload for pageable memory)."""
    U = np.ascontiguousarray(U, dtype=np.float64)
    n_h, n = U.shape
    Ud = _slab(ctx, n_h, n)
    ctx._slab_shape = None
    ctx.check(_lib.lib().pb_upload_matrix(ctx.h, U.ctypes.data, n_h, n, n, Ud.ptr))
    return Ud


def perform_fast_pod(U, eps, eps_init, *, mode=DEFAULT_MODE, ctx=None, return_ranks=False):
    """Reference signature (pod.py:51).  U is (n_h, n_t, n_s); returns V (n_h, n_L)."""
    ctx = ctx or _lib.default_context()
    print("Performing initial time-trajectory POD")  # pod.py:53
    U = np.asarray(U, dtype=np.float64)
    n_h, n_t, n_s = U.shape
    Us = ctx.to_device(U)
    Um = ctx.alloc(n_h, n_s * n_t)
    L = _lib.lib()
    ctx.check(L.pb_struct_to_matrix(ctx.h, Us.ptr, n_h, n_t, n_s, Um.ptr, Um.ld))
    Us.release()
    V, ranks = fast_pod_device(ctx, Um, n_t, n_s, eps, eps_init, mode)
    print("Performing SVD")                          # pod.py:67
    V = V.download()
    return (V, ranks) if return_ranks else V


def fast_pod_device(ctx, Um, n_t, n_s, eps, eps_init, mode=DEFAULT_MODE):
    """Two-step POD of a device matrix with sample-major columns (column = i*n_t + j)."""
    n_h = Um.shape[0]
    nl = C.c_int()
    ranks = (C.c_int * n_s)()
    ctx.check(_lib.lib().pb_fast_pod(ctx.h, Um.ptr, n_h, n_t, n_s, Um.ld, float(eps), float(eps_init),
                                     POD_MODES[mode], C.byref(nl), ranks))
    V = ctx.alloc(n_h, nl.value)
    ctx.check(_lib.lib().pb_fast_pod_basis(ctx.h, V.ptr, V.ld))
    return V, np.array(ranks[:], dtype=np.int64)
