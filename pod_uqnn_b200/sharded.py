"""Row-sharded POD / projection across the GPUs of one box (one process per GPU).

U is sharded by dof: rank p owns the contiguous row slab row_range(n_h, p, P) of the
row-major snapshot matrix.  Exchange steps (the only ones on the path):
  * sum-all-reduce of the local n_st x n_st Gram partial (and of the second-pass Gram),
  * sum-all-reduce of the local projection partial V_p^T U_p,
  * gram2 mode, only when the tolerance estimate asks for the subspace-refinement step: two more small all-reduces,
  * TSQR mode: all-gather of the P local R factors (n_st x n_st each), then the QR of the stack is replicated.
The small eigensolve is replicated: every rank runs it on the bit-identical reduced Gram
and gets the same rank and factors.  Snapshot generation, basis, reconstruction and
pod_sig need no communication.

torch is the carrier of the buffers that cross ranks (torch.distributed over NCCL on the
GPU box, gloo in the CPU tests); the compute is injected as a `backend` so that the host
logic here is testable without a GPU (tests/test_sharded_gloo.py uses a numpy test double).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import POD_MODES


def row_range(n, rank, world):
    """Contiguous, balanced slab [lo, hi) of n rows for `rank` of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _all_reduce(t, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def _all_gather_stack(backend, R, group=None):
    """(P m x m) stack of every rank's m x m factor, in rank order (identical on all ranks)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return R
    return backend.all_gather_rows(R, group)


def _max_over_ranks(backend, value, group=None):
    """max of a host integer over the ranks (the row count of the truncated factors)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(value)
    dev = getattr(backend, "dev", None) if dist.get_backend(group) == "nccl" else None
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


SMALL_STACK_ROWS = 512      # stacked leaf factors up to this many rows go to the pivoted QR directly


def pod_sharded_tsqr(backend, A_local, eps=0., n_L=0, group=None, full=False):
    """TSQR mode: one leaf per rank.  F_p = the rank-revealing Householder factor of A_p (pb_tsqr_factor: block
    pivoting, stops at the numerical rank; Q not formed), all-gather of the first max_p rows(F_p) rows, the factor of
    the stack and its Jacobi SVD are replicated (same bits on every rank, hence the same rank and factors),
    V_p = A_p Z / sigma locally (pod.py:42-45).  eps == 0 without n_L asks for every mode: nothing may be dropped, the
    plain triangular R factors are used (also with full=True, the "tsqr_full" mode).  Returns (V_local, sigma, n_L)."""
    if full or (eps == 0. and n_L == 0):
        R = backend.tsqr_r(A_local)
        stack = _all_gather_stack(backend, R, group)
        if stack is not R:
            R = backend.tsqr_r(stack)
        n_L_out = backend.pod_from_r(R, eps, n_L)
    else:
        F, rows = backend.tsqr_factor(A_local)
        cap = _max_over_ranks(backend, rows, group)
        Fc = F[:cap]
        stack = _all_gather_stack(backend, Fc, group)
        if stack is not Fc:
            if stack.shape[0] <= min(SMALL_STACK_ROWS, stack.shape[1]):
                # a short stack IS a valid factor (A^T A = S^T S): the pivoted QR of pod_from_factor takes it as it is; a
                # blocked QR of it first would be all fixed latency (2 panels ~ 1.6 ms at C2 for 256 x 1000)
                F, rows = stack, stack.shape[0]
            else:
                F, rows = backend.tsqr_factor(stack)
        n_L_out = backend.pod_from_factor(F, rows, eps, n_L)
    V = backend.basis(A_local, None, n_L_out)
    return V, backend.sigma(), n_L_out


def pod_sharded(backend, A_local, eps=0., n_L=0, mode="gram", group=None, A_host=None):
    """POD of the row-sharded tall matrix.  Returns (V_local, sigma, n_L).

    `A_host` (address of this rank's slab in pinned host memory, same shape as `A_local`): the slab is uploaded into
    `A_local` in row chunks overlapped with the chunk Grams (pb_gram_host) instead of being read from HBM."""
    if mode in ("tsqr", "tsqr_full"):
        if A_host is not None:
            backend.upload(A_host, A_local)
        return pod_sharded_tsqr(backend, A_local, eps, n_L, group, full=(mode == "tsqr_full"))
    # the gram2 refinement decision depends on sigma (replicated), the column count and a row count: every rank uses the
    # LARGEST slab of the job (one MAX all-reduce per shape, cached), so the decision needs no collective per call
    key = (id(group), int(A_local.shape[0]), int(A_local.shape[1]))
    cache = backend.__dict__.setdefault("_rows_max_cache", {})
    if key not in cache:
        cache[key] = _max_over_ranks(backend, A_local.shape[0], group)
    G = backend.gram(A_local) if A_host is None else backend.gram_host(A_host, A_local)
    backend.reduce(G, group)
    backend.set_refine_rows(cache[key])          # (every new Gram clears the override: it belongs to this decomposition)
    n_L_out, keep = backend.pod_from_gram(G, eps, n_L, mode)
    Y = None
    if keep > 0:
        Y, G2 = backend.pass2_gram(A_local, keep)
        backend.reduce(G2, group)
        n_L_out = backend.pass2_finish(G2, eps, n_L)
    V = backend.basis(A_local, Y, n_L_out)
    if keep > 0 and backend.refine_steps() > 0:     # the same on every rank (replicated inputs, see above)
        # gram2 tolerance safety (csrc/api.cu, "subspace refinement"): one step of subspace iteration with the matrix
        # itself -- two more all-reduces (the n_st x n_L projection block and an n_L x n_L Gram)
        B = backend.project(V, A_local)
        backend.reduce(B, group)
        G3 = backend.refine_apply(A_local, B)
        backend.reduce(G3, group)
        backend.refine_finish(G3, V)
    return V, backend.sigma(), n_L_out


def project_sharded(backend, V_local, U_local, group=None):
    """v (n x n_L) = sum over ranks of (V_p^T U_p)^T."""
    v = backend.project(V_local, U_local)
    backend.reduce(v, group)
    return v


class CudaBackend:
    """libpodb200 on this rank's GPU; reduced buffers are torch CUDA tensors."""

    def __init__(self, ctx):
        import torch
        self.ctx, self.torch = ctx, torch
        self.dev = torch.device("cuda", ctx.device)
        self.L = _lib.lib()
        self._cache = {}
        self.qr_stats = {}
        # the library and NCCL share one torch stream: the all-reduces are ordered against the kernels on the
        # device (NCCL's stream waits on / is waited for by this stream through events), no host synchronisation.
        # It is made torch's CURRENT stream for the life of the backend (restored by close()): a `with
        # torch.cuda.stream(...)` around every allocation / collective costs host time exactly where the GPU is
        # waiting for the host (after the data-dependent syncs of the eigensolve)
        self.stream = torch.cuda.Stream(self.dev)
        self._prev_stream = torch.cuda.current_stream(self.dev)
        torch.cuda.set_stream(self.stream)
        ctx.check(self.L.pb_set_stream(ctx.h, C.c_void_p(self.stream.cuda_stream)))

    def close(self):
        """Give the context its own stream back and restore torch's current stream."""
        self.stream.synchronize()
        self.torch.cuda.set_stream(self._prev_stream)
        self.ctx.check(self.L.pb_set_stream(self.ctx.h, None))

    def _empty(self, *shape):
        return self.torch.empty(*shape, dtype=self.torch.float64, device=self.dev)

    def _scratch(self, name, *shape):
        """torch tensor reused across steps for internal temporaries (the Gram partials that are all-reduced)."""
        t = self._cache.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self._cache[name] = self._empty(*shape)
        return t

    def _buf(self, name, *shape):
        """Library-owned HBM buffer reused across steps (cudaMalloc / cudaFree synchronise the device)."""
        b = self._cache.get(name)
        if b is None or b.shape != tuple(shape):
            b = self._cache[name] = self.ctx.alloc(*shape)
        return b

    def reduce(self, t, group=None):
        _all_reduce(t, group)

    def gram(self, A):
        m = A.shape[1]
        G = self._scratch("G", m, m)
        self.ctx.check(self.L.pb_gram(self.ctx.h, A.ptr, A.shape[0], m, A.ld, G.data_ptr()))
        return G

    def gram_host(self, host_ptr, A):
        m = A.shape[1]
        G = self._scratch("G", m, m)
        self.ctx.check(self.L.pb_gram_host(self.ctx.h, host_ptr, A.shape[0], m, m, A.ptr, G.data_ptr()))
        return G

    def upload(self, host_ptr, A):
        self.ctx.check(self.L.pb_upload(self.ctx.h, A.ptr, host_ptr, A.nbytes))

    def tsqr_r(self, A):
        """R factor (torch tensor m x m) of a DeviceArray slab or of a torch (rows x m) stack."""
        is_t = isinstance(A, self.torch.Tensor)
        rows, m = A.shape
        R = self._empty(m, m)
        ptr, ld = (A.data_ptr(), A.stride(0)) if is_t else (A.ptr, A.ld)
        self.ctx.check(self.L.pb_tsqr_r(self.ctx.h, ptr, rows, m, ld, R.data_ptr()))
        self._note_qr(is_t)
        return R

    def _note_qr(self, stacked):
        """(panels, factor rows, Householder flops) of the factorisation just run, kept per stage for reporting."""
        p, f, fl = C.c_int64(), C.c_int64(), C.c_double()
        self.ctx.check(self.L.pb_tsqr_stats(self.ctx.h, C.byref(p), C.byref(f), C.byref(fl)))
        self.qr_stats["stack" if stacked else "local"] = (p.value, f.value, fl.value)

    def tsqr_factor(self, A):
        """Rank-revealing factor of a DeviceArray slab or of a torch (rows x m) stack: (F torch tensor m x m, rows used)."""
        is_t = isinstance(A, self.torch.Tensor)
        rows, m = A.shape
        F = self._empty(m, m)
        ptr, ld = (A.data_ptr(), A.stride(0)) if is_t else (A.ptr, A.ld)
        fr = C.c_int64()
        self.ctx.check(self.L.pb_tsqr_factor(self.ctx.h, ptr, rows, m, ld, F.data_ptr(), C.byref(fr)))
        self._note_qr(is_t)
        return F, fr.value

    def pod_from_factor(self, F, rows, eps, n_L):
        nl = C.c_int()
        self.ctx.check(self.L.pb_pod_from_factor(self.ctx.h, F.data_ptr(), int(rows), F.shape[1], F.stride(0), float(eps),
                                                 int(n_L), C.byref(nl)))
        return nl.value

    def all_gather_rows(self, R, group=None):
        import torch.distributed as dist
        world = dist.get_world_size(group)
        stack = self._scratch("Rstack", world * R.shape[0], R.shape[1])
        dist.all_gather_into_tensor(stack, R.contiguous(), group=group)
        return stack

    def pod_from_r(self, R, eps, n_L):
        nl = C.c_int()
        self.ctx.check(self.L.pb_pod_from_r(self.ctx.h, R.data_ptr(), R.shape[0], float(eps), int(n_L), C.byref(nl)))
        return nl.value

    def pod_from_gram(self, G, eps, n_L, mode):
        nl, keep = C.c_int(), C.c_int()
        self.ctx.check(self.L.pb_pod_from_gram(self.ctx.h, G.data_ptr(), G.shape[0], float(eps), int(n_L),
                                               POD_MODES[mode], C.byref(nl), C.byref(keep)))
        return nl.value, keep.value

    def pass2_gram(self, A, keep):
        Y = self._buf("Y", A.shape[0], keep)
        G2 = self._scratch("G2", keep, keep)
        self.ctx.check(self.L.pb_pod_pass2_gram(self.ctx.h, A.ptr, A.shape[0], A.ld, Y.ptr, G2.data_ptr()))
        return Y, G2

    def pass2_finish(self, G2, eps, n_L):
        nl = C.c_int()
        self.ctx.check(self.L.pb_pod_pass2_finish(self.ctx.h, G2.data_ptr(), float(eps), int(n_L), C.byref(nl)))
        return nl.value

    def basis(self, A, Y, n_L):
        V = self._buf("V", A.shape[0], n_L)
        self.ctx.check(self.L.pb_pod_basis(self.ctx.h, A.ptr, A.shape[0], A.ld, Y.ptr if Y is not None else None,
                                           V.ptr, V.ld))
        return V

    def set_refine_rows(self, rows):
        self.ctx.check(self.L.pb_pod_set_refine_rows(self.ctx.h, int(rows)))

    def refine_steps(self):
        st = C.c_int()
        self.ctx.check(self.L.pb_pod_refine_steps(self.ctx.h, C.byref(st)))
        return st.value

    def refine_apply(self, A, B):
        n_L = B.shape[1]
        G3 = self._scratch("G3", n_L, n_L)
        self.ctx.check(self.L.pb_pod_refine_apply(self.ctx.h, A.ptr, A.shape[0], A.ld, B.data_ptr(), G3.data_ptr()))
        return G3

    def refine_finish(self, G3, V):
        self.ctx.check(self.L.pb_pod_refine_finish(self.ctx.h, G3.data_ptr(), V.shape[0], V.ptr, V.ld))

    def sigma(self):
        from .pod import pod_sigma
        return pod_sigma(self.ctx)

    def project(self, V, U):
        n, n_L = U.shape[1], V.shape[1]
        v = self._empty(n, n_L)
        self.ctx.check(self.L.pb_project(self.ctx.h, V.ptr, V.ld, n_L, U.ptr, U.ld, U.shape[0], n, v.data_ptr()))
        return v


def train_ensemble_sharded(train_fn, members, group=None):
    """Ensemble training across ranks: the members are independent (same data, different initialisation), so rank p
    trains the contiguous slice row_range(n_M, p, P) -- the reference's distributed mode gives one member to each
    Horovod rank (experiments/*/train.py:19-33).  No exchange during training; afterwards the trained weights and
    loss histories are all-gathered so that every rank holds the whole ensemble for predict_v.

    `train_fn(member_slice) -> (trained_member_slice, losses (epochs, len(slice)))` is the per-rank trainer
    (EnsembleTrainer on the GPU box; a numpy double in the gloo test).  Returns (members, losses (epochs, n_M)).
    """
    import torch
    import torch.distributed as dist
    n_M = len(members)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return train_fn(members)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = row_range(n_M, rank, world)
    shapes = [np.shape(a) for a in members[0]]
    per = int(sum(int(np.prod(s)) for s in shapes))
    mine, losses = train_fn(members[lo:hi]) if hi > lo else ([], np.zeros((0, 0)))
    epochs = int(np.shape(losses)[0]) if hi > lo else 0
    ep = torch.tensor([epochs], dtype=torch.int64)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    ep = ep.to(dev)
    dist.all_reduce(ep, op=dist.ReduceOp.MAX, group=group)
    epochs = int(ep.item())
    cap = -(-n_M // world)                                    # slots per rank (ranks with fewer members pad)
    buf = np.zeros((cap, per + epochs))
    for i, w in enumerate(mine):
        buf[i, :per] = np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in w])
        buf[i, per:] = np.asarray(losses)[:, i]
    t = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    out_members, out_losses = [], np.zeros((epochs, n_M))
    for p in range(world):
        a, b = row_range(n_M, p, world)
        block = parts[p].cpu().numpy()
        for i in range(b - a):
            flat, off, w = block[i, :per], 0, []
            for s in shapes:
                n = int(np.prod(s))
                w.append(flat[off:off + n].reshape(s).copy()); off += n
            out_members.append(w)
            out_losses[:, a + i] = block[i, per:]
    return out_members, out_losses
