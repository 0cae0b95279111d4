"""Numpy test double of the compute backend used by pod_uqnn_b200.sharded (CPU tests only).

It lets the gloo world_size-2 test exercise the HOST logic of the sharded path (row
partition, the two all-reduces, replicated eigensolve, local basis) without a GPU.  It is
not a fallback: the product never imports it."""
import numpy as np
import torch

from oracle import pod_oracle as po


class NumpyBackend:
    def __init__(self):
        self._sig = None
        self._W = None

    def reduce(self, t, group=None):
        from pod_uqnn_b200.sharded import _all_reduce
        _all_reduce(t, group)

    def gram(self, A):
        return torch.from_numpy(A.T @ A)

    def tsqr_r(self, A):
        A = A.numpy() if isinstance(A, torch.Tensor) else A
        return torch.from_numpy(np.linalg.qr(A, mode="r"))

    def tsqr_factor(self, A):
        """Double of pb_tsqr_factor: pivoted QR cut at the numerical rank, columns back in input order; (m x m buffer, rows)."""
        import scipy.linalg as sl
        A = A.numpy() if isinstance(A, torch.Tensor) else A
        m = A.shape[1]
        R, piv = sl.qr(A, mode="r", pivoting=True)
        d = np.abs(np.diag(R))
        k = max(1, int((d > 16 * np.finfo(float).eps * d[0]).sum()))
        F = np.zeros((m, m))
        F[:k, piv] = R[:k]
        return torch.from_numpy(F), k

    def pod_from_factor(self, F, rows, eps, n_L):
        return self.pod_from_r(F[:rows], eps, n_L)

    def all_gather_rows(self, R, group=None):
        import torch.distributed as dist
        parts = [torch.empty_like(R) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, R.contiguous(), group=group)
        return torch.cat(parts, 0)

    def pod_from_r(self, R, eps, n_L):
        _, s, Zt = np.linalg.svd(R.numpy(), full_matrices=True)
        s = np.concatenate([s, np.zeros(Zt.shape[0] - len(s))])
        self._sig, self._W = s, Zt.T
        if n_L == 0:
            n_L = po.truncation_rank(s ** 2, eps, len(s))
        return n_L

    def pod_from_gram(self, G, eps, n_L, mode):
        lam, W = np.linalg.eigh(G.numpy())
        lam, W = np.maximum(lam[::-1], 0.), W[:, ::-1]
        self._sig, self._W = np.sqrt(lam), W
        if n_L == 0:
            n_L = po.truncation_rank(lam, eps, len(lam))
        self._n_L = n_L
        return n_L, (0 if mode == "gram" else int((lam > 1e-15 * lam[0]).sum()))

    def pass2_gram(self, A, keep):
        Y = A @ self._W[:, :keep]
        return Y, torch.from_numpy(Y.T @ Y)

    def pass2_finish(self, G2, eps, n_L):
        lam, W2 = np.linalg.eigh(G2.numpy())
        lam, W2 = np.maximum(lam[::-1], 0.), W2[:, ::-1]
        self._sig2, self._W2 = np.sqrt(lam), W2
        if n_L == 0:
            n_L = po.truncation_rank(lam, eps, len(lam))
        self._sig = np.concatenate([self._sig2, np.zeros(len(self._sig) - len(lam))])
        return n_L

    def basis(self, A, Y, n_L):
        if Y is not None:
            return Y @ self._W2[:, :n_L] / self._sig2[:n_L]
        return A @ self._W[:, :n_L] / self._sig[:n_L]

    # gram2 subspace refinement (always engaged here so that the gloo test walks the two extra all-reduces)
    def set_refine_rows(self, rows):
        self.refine_rows = int(rows)

    def refine_steps(self):
        return 1

    def refine_apply(self, A, B):
        Bn = B.numpy()
        Bn /= self._sig2[:Bn.shape[1]] ** 2
        self._Vt = A @ Bn
        return torch.from_numpy(self._Vt.T @ self._Vt)

    def refine_finish(self, G3, V):
        R = np.linalg.cholesky(G3.numpy()).T
        V[:] = self._Vt @ np.linalg.inv(R)

    def sigma(self):
        return self._sig

    def project(self, V, U):
        return torch.from_numpy(np.ascontiguousarray((V.T @ U).T))
