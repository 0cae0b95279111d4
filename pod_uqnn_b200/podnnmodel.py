"""PodnnModel -- the dataset-generation / projection / predict path of the reference's
poduqnn/podnnmodel.py behind the same method names, running on libpodb200.

Kept: create_snapshots :89, generate_dataset :199, convert_multigpu_data :118 (POD +
projection part), project_to_v/U :431-437, predict_v :314, predict :366 (closed-form moments
or Monte-Carlo), restruct/destruct :382-421, the pickle caches :443-503 (same tuples).
Out of scope (raise): train_model :289.  Host-side bookkeeping (LHS sampling, splits, pickles)
stays in numpy exactly as in the reference.
"""
import ctypes as C
import os
import pickle
import time

import numpy as np

from . import _lib
from .acceleration import snapshots_device
from .pod import pod_device, fast_pod_device, DEFAULT_MODE
from .varneuralnetwork import (VarNeuralNetwork, NORM_MEANSTD, ensemble_predict_device, EnsembleTrainer,
                               normalize_host)
from .logger import Logger
from .metrics import re_s

SETUP_DATA_NAME = "setup_data.pkl"
TRAIN_DATA_NAME = "train_data.pkl"
INIT_DATA_NAME = "init_data.pkl"
MODEL_PARAMS_NAME = "model_params.pkl"
MODEL_NAME = "model_weights"
MODEL_NAME_EXT = ".npz"


def lhs(n, samples):
    """pyDOE-style centred LHS as the reference's acceleration.lhs :74-94 (numpy RNG)."""
    cut = np.linspace(0, 1, samples + 1)
    u = np.random.rand(samples, n)
    a, b = cut[:samples], cut[1:samples + 1]
    rd = u * (b - a)[:, None] + a[:, None]
    H = np.zeros_like(rd)
    for j in range(n):
        H[:, j] = rd[np.random.permutation(samples), j]
    return H


def sample_mu(n_s, mu_min, mu_max):
    """handling.py:40-47 (including the transposed lhs call)."""
    X_lhs = lhs(n_s, mu_min.shape[0]).T
    return mu_min + (mu_max - mu_min) * X_lhs


def split_dataset(X_v, v, test_size):
    """handling.py:30-37."""
    idx = np.random.permutation(X_v.shape[0])
    limit = np.floor(X_v.shape[0] * (1. - test_size)).astype(int)
    tr, te = idx[:limit], idx[limit:]
    return X_v[tr], X_v[te], v[tr], v[te]


class PodnnModel:
    def __init__(self, resdir, n_v, x_mesh, n_t, *, ctx=None, pod_mode=DEFAULT_MODE, n_gpus=1, multi=None):
        self.n_v, self.x_mesh = n_v, x_mesh
        self.n_xyz = x_mesh.shape[0]
        self.n_h = self.n_v * x_mesh.shape[0]
        self.n_t, self.has_t = n_t, n_t > 0
        self.resdir = resdir
        self.setup_data_path = os.path.join(resdir, SETUP_DATA_NAME)
        self.train_data_path = os.path.join(resdir, TRAIN_DATA_NAME)
        self.init_data_path = os.path.join(resdir, INIT_DATA_NAME)
        self.model_params_path = os.path.join(resdir, MODEL_PARAMS_NAME)
        self.model_path = []
        self.regnn = None
        self.n_L = self.n_d = self.V = self.layers = self.pod_sig = None
        self.dtype = "float64"
        self.pod_mode = pod_mode
        self._ctx = ctx
        self._V_dev = None
        # n_gpus > 1: snapshots, POD, projection and pod_sig run row-sharded (by dof) over GPUs 0..n_gpus-1 from this
        # process (pb_multi_*: the library's own NVLink collectives, no torchrun), predict_v splits the batch
        # (`multi`: an existing _lib.MultiContext to use instead of the process-wide one over GPUs 0..n_gpus-1)
        self._multi = multi
        self.n_gpus = int(multi.n_gpus if multi is not None else n_gpus)
        self.save_setup_data()

    @property
    def multi(self):
        return self._multi if self._multi is not None else _lib.multi_context(self.n_gpus)

    def _sharded(self, n_cols):
        """Row-shard when asked to and the matrix is tall enough for every GPU to hold a useful slab."""
        return self.n_gpus > 1 and self.n_h >= max(n_cols, 4096 * self.n_gpus)

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def generate_hifi_inputs(self, n_s, mu_min, mu_max, t_min=0, t_max=0):
        """podnnmodel.py:60-87: the (t, mu) regression inputs of a large prediction task (LHS over mu, every time step
        per sample).  Rows s*n_t + j = [t_j, mu_s...] for time-dependent models, [mu_s...] otherwise."""
        mu_min, mu_max = np.array(mu_min), np.array(mu_max)
        mu_lhs = sample_mu(n_s, mu_min, mu_max)
        if not self.has_t:
            return mu_lhs.copy()
        t = np.linspace(t_min, t_max, self.n_t)                 # podnnmodel.py:75
        return np.hstack((np.tile(t, n_s)[:, None], np.repeat(mu_lhs, self.n_t, axis=0)))

    # ------------------------------------------------------------ snapshots
    def create_snapshots(self, n_d, n_h, u, mu_lhs, t_min=0, t_max=0, u_noise=0., x_noise=0.):
        """podnnmodel.py:89-116 -> (X_v, U, U_struct, U_no_noise), host arrays."""
        if u_noise > 0. or x_noise > 0.:
            raise NotImplementedError("noisy snapshots are outside the CUDA hot path")
        X = self.x_mesh[:, 1:].T
        Ud, X_v, _ = snapshots_device(self.ctx, u, X, mu_lhs, self.n_t, t_min, t_max)
        U = Ud.download()
        if self.has_t:
            n_s = mu_lhs.shape[0]
            Us = self.ctx.alloc(n_h, self.n_t, n_s)
            self.ctx.check(_lib.lib().pb_matrix_to_struct(self.ctx.h, Ud.ptr, Ud.ld, n_h, self.n_t, n_s, Us.ptr))
            return X_v, U, Us.download(), U.copy()
        return X_v, U, U, U.copy()

    # ------------------------------------------------------------ POD + projection on the device
    def _pod_project(self, U_train_d, U_val_d, eps, eps_init, n_L, n_s_train):
        ctx, L = self.ctx, _lib.lib()
        if eps_init is None or not self.has_t:
            Vd, _ = pod_device(ctx, U_train_d, eps, n_L, self.pod_mode)
        else:
            Vd, _ = fast_pod_device(ctx, U_train_d, self.n_t, n_s_train, eps, eps_init, self.pod_mode)
        self._V_dev = Vd
        self.V = Vd.download()
        self.n_L = self.V.shape[1]
        return self._finish_projection(U_train_d, U_val_d)

    def _pod_project_sharded(self, Ut, Uv, eps, n_L):
        """Row-sharded POD + projection + pod_sig on per-GPU slabs (lists of DeviceArray, one per rank): podnnmodel.py:239,
        247-253 with the Gram / projection partials summed over NVLink.  U_pod is never materialised."""
        m, L = self.multi, _lib.lib()
        Vs, _, self.n_L = m.pod(Ut, eps, n_L, self.pod_mode)
        self.V = np.vstack([v.download() for v in Vs])
        self._V_dev = None
        vt = m.project(Vs, Ut)                                  # replicated on every GPU
        v_train = vt[0].download()
        v_val = m.project(Vs, Uv)[0].download() if Uv is not None else np.zeros((0, self.n_L))
        sigs = []
        for p, c in enumerate(m.ctxs):
            rows = Ut[p].shape[0]
            sig = c.alloc(max(rows, 1))
            if rows:
                c.check(L.pb_reconstruct(c.h, Vs[p].ptr, Vs[p].ld, self.n_L, vt[p].ptr, v_train.shape[0], rows,
                                         Ut[p].ptr, Ut[p].ld, None, 0, sig.ptr))
            sigs.append((sig, rows))
        self.pod_sig = np.concatenate([s.download()[:r] for s, r in sigs])
        print(f"Mean pod sig: {self.pod_sig.mean()}")
        return v_train, v_val

    def _project_dev(self, Ud):
        ctx = self.ctx
        n = Ud.shape[1]
        v = ctx.alloc(n, self.n_L)
        ctx.check(_lib.lib().pb_project(ctx.h, self._V_dev.ptr, self._V_dev.ld, self.n_L, Ud.ptr, Ud.ld,
                                        self.n_h, n, v.ptr))
        return v

    def _ensure_V_dev(self):
        if self._V_dev is None:
            if self.V is None:
                raise ValueError("no POD basis: call generate_dataset / convert_multigpu_data / load_train_data")
            self._V_dev = self.ctx.to_device(self.V)
        return self._V_dev

    def generate_dataset(self, u, mu_min, mu_max, n_s, train_val, eps=0., eps_init=None, n_L=0,
                         t_min=0, t_max=0, u_noise=0., x_noise=0., rm_init=False, *, mu_lhs=None):
        """podnnmodel.py:199-275.  `mu_lhs` (keyword) bypasses the LHS sampling for reproducible runs."""
        if u_noise > 0. or x_noise > 0.:
            raise NotImplementedError("noisy snapshots are outside the CUDA hot path")
        mu_min, mu_max = np.array(mu_min), np.array(mu_max)
        n_d = mu_min.shape[0] + (1 if self.has_t else 0)
        self.n_d = n_d
        print("Doing the LHS sampling on the non-spatial params...")
        if mu_lhs is None:
            mu_lhs = sample_mu(n_s, mu_min, mu_max)
        _, _, mu_tr, mu_val = split_dataset(np.zeros_like(mu_lhs), mu_lhs, test_size=train_val[1])
        n_st = n_s * (self.n_t if self.has_t else 1)
        print(f"Generating {n_st} corresponding snapshots")
        X = self.x_mesh[:, 1:].T
        two_step = eps_init is not None and self.has_t
        if not two_step and self._sharded(mu_tr.shape[0] * (self.n_t if self.has_t else 1)):
            m = self.multi
            rr = [m.row_range(self.n_h, p) for p in range(m.n_gpus)]
            tr = [snapshots_device(m.ctxs[p], u, X, mu_tr, self.n_t, t_min, t_max, row0=lo, n_rows=hi - lo)
                  for p, (lo, hi) in enumerate(rr)]
            va = [snapshots_device(m.ctxs[p], u, X, mu_val, self.n_t, t_min, t_max, row0=lo, n_rows=hi - lo)
                  for p, (lo, hi) in enumerate(rr)]
            X_v_train, X_v_val = tr[0][1], va[0][1]
            Ut, Uv = [a[0] for a in tr], [a[0] for a in va]
            v_train, v_val = self._pod_project_sharded(Ut, Uv, eps, n_L)
            U_train, U_val = np.vstack([a.download() for a in Ut]), np.vstack([a.download() for a in Uv])
        else:
            U_tr_d, X_v_train, _ = snapshots_device(self.ctx, u, X, mu_tr, self.n_t, t_min, t_max)
            U_val_d, X_v_val, _ = snapshots_device(self.ctx, u, X, mu_val, self.n_t, t_min, t_max)
            v_train, v_val = self._pod_project(U_tr_d, U_val_d, eps, eps_init, n_L, mu_tr.shape[0])
            U_train, U_val = U_tr_d.download(), U_val_d.download()
        if self.n_t > 0 and rm_init:        # podnnmodel.py:256-272
            idx = np.arange(mu_tr.shape[0]) * self.n_t
            init = [X_v_train[idx], v_train[idx], U_train[:, idx]]
            X_v_train, v_train, U_train = (np.delete(X_v_train, idx, 0), np.delete(v_train, idx, 0),
                                           np.delete(U_train, idx, 1))
            idx = np.arange(mu_val.shape[0]) * self.n_t
            init += [X_v_val[idx], v_val[idx], U_val[:, idx]]
            X_v_val, v_val, U_val = np.delete(X_v_val, idx, 0), np.delete(v_val, idx, 0), np.delete(U_val, idx, 1)
            self.save_init_data(*init)
        self.save_train_data(X_v_train, v_train, U_train, X_v_val, v_val, U_val)
        return X_v_train, v_train, U_train, X_v_val, v_val, U_val

    def convert_multigpu_data(self, U_struct, X_v, train_val, eps, eps_init=None, n_L=0,
                              use_cache=False, save_cache=False):
        """podnnmodel.py:118-197 without the initial-condition removal (:179-194 has index slips,
        SURVEY App. B).  U_struct is (n_v, n_xyz, n_t, n_s) or (n_v, n_xyz, n_s)."""
        if use_cache and os.path.exists(self.train_data_path):
            return self.load_train_data()
        n_s = U_struct.shape[-1]
        n_t = self.n_t if self.n_t > 0 else 1
        self.n_d = X_v.shape[1]
        idx_s = np.random.permutation(n_s)
        limit = np.floor(n_s * (1. - train_val[1])).astype(int)
        tr, va = idx_s[:limit], idx_s[limit:]
        X_v_train = np.concatenate([X_v[i*n_t:(i+1)*n_t] for i in tr])
        X_v_val = np.concatenate([X_v[i*n_t:(i+1)*n_t] for i in va]) if len(va) else np.zeros((0, X_v.shape[1]))
        U_train = self.destruct(U_struct[..., tr])
        U_val = self.destruct(U_struct[..., va])
        ctx = self.ctx
        if not (eps_init is not None and self.has_t) and self._sharded(U_train.shape[1]):
            m = self.multi
            rr = [m.row_range(self.n_h, p) for p in range(m.n_gpus)]
            Ut = [m.ctxs[p].to_device(U_train[lo:hi]) for p, (lo, hi) in enumerate(rr)]
            Uv = [m.ctxs[p].to_device(U_val[lo:hi]) for p, (lo, hi) in enumerate(rr)] if len(va) else None
            v_train, v_val = self._pod_project_sharded(Ut, Uv, eps, n_L)
            self.save_train_data(X_v_train, v_train, U_train, X_v_val, v_val, U_val)
            return X_v_train, v_train, X_v_val, v_val, U_val
        Ut, Uv = ctx.to_device(U_train), (ctx.to_device(U_val) if len(va) else None)
        if eps_init is not None and self.has_t:
            # the reference feeds ALL samples (incl. validation) to the two-step POD (:160)
            Uall = ctx.to_device(self.destruct(U_struct))
            Vd, _ = fast_pod_device(ctx, Uall, n_t, n_s, eps, eps_init, self.pod_mode)
            self._V_dev, self.V = Vd, Vd.download()
            self.n_L = self.V.shape[1]
            v_train, v_val = self._finish_projection(Ut, Uv)
        else:
            v_train, v_val = self._pod_project(Ut, Uv, eps, None, n_L, len(tr))
        # the reference saves unconditionally (:196; `save_cache` is accepted and ignored there too): gen.py ->
        # train.py relies on PodnnModel.load() finding train_data.pkl
        self.save_train_data(X_v_train, v_train, U_train, X_v_val, v_val, U_val)
        return X_v_train, v_train, X_v_val, v_val, U_val

    def _finish_projection(self, Ut, Uv):
        ctx, L = self.ctx, _lib.lib()
        vt = self._project_dev(Ut)                 # stays in HBM for the POD error below (podnnmodel.py:251-253,
        v_train = vt.download()                    # fused: U_pod is never materialised)
        v_val = self._project_dev(Uv).download() if Uv is not None else np.zeros((0, self.n_L))
        sig = ctx.alloc(self.n_h)
        ctx.check(L.pb_reconstruct(ctx.h, self._V_dev.ptr, self._V_dev.ld, self.n_L, vt.ptr, v_train.shape[0],
                                   self.n_h, Ut.ptr, Ut.ld, None, 0, sig.ptr))
        self.pod_sig = sig.download()
        print(f"Mean pod sig: {self.pod_sig.mean()}")
        return v_train, v_val

    # ------------------------------------------------------------ projection helpers
    def project_to_U(self, v):
        """V.dot(v.T), podnnmodel.py:431-433."""
        ctx, Vd = self.ctx, self._ensure_V_dev()
        v = np.ascontiguousarray(v, dtype=np.float64)
        vd, out = ctx.to_device(v), ctx.alloc(self.n_h, v.shape[0])
        ctx.check(_lib.lib().pb_reconstruct(ctx.h, Vd.ptr, Vd.ld, self.n_L, vd.ptr, v.shape[0], self.n_h,
                                            None, 0, out.ptr, out.ld, None))
        return out.download()

    def project_to_v(self, U):
        """(V.T.dot(U)).T, podnnmodel.py:435-437."""
        self._ensure_V_dev()
        return self._project_dev(self.ctx.to_device(np.asarray(U, dtype=np.float64))).download()

    # ------------------------------------------------------------ ensemble
    def initVNNs(self, n_M, h_layers, lr, lam, adv_eps, soft_0=1., norm=NORM_MEANSTD, seed=None):
        """podnnmodel.py:277-287: creates the (untrained, glorot-initialised) ensemble."""
        self.layers = [self.n_d, *h_layers, self.n_L]
        self.regnn, self.model_path = [], []
        for i in range(n_M):
            self.regnn.append(VarNeuralNetwork(self.layers, lr, lam, adv_eps, soft_0, norm,
                                               seed=None if seed is None else seed + i))
            self.model_path.append(os.path.join(self.resdir, f"{MODEL_NAME}-{i}-{time.time()}"))
        self.save_model()

    def train_model(self, model_id, X_v_train, v_train, X_v_val, v_val, epochs, freq=100, div_max=False):
        """podnnmodel.py:291-312: trains member `model_id` (csrc/train.cu), logging RE_val / MPIW_val every
        `freq` epochs."""
        if self.regnn is None or len(self.regnn) == 0:
            raise ValueError("Regression model isn't defined.")
        model = self.regnn[model_id]

        def get_val_err():
            v_val_pred, var = model.predict(X_v_val, ctx=self.ctx)
            return {"RE_val": re_s(v_val.T, v_val_pred.T, div_max), "MPIW_val": 4 * var.mean()}
        logger = Logger(epochs, freq)
        logger.set_val_err_fn(get_val_err)
        model.fit(X_v_train, v_train, epochs, logger, ctx=self.ctx)
        return [logger.get_logs()]

    def train_ensemble(self, X_v_train, v_train, epochs, *, use_graph=True, group=None):
        """All members in the same launches (no reference counterpart: its train.py loops over the members or gives
        one to each Horovod rank, experiments/*/train.py:30-34).  Under an initialised torch.distributed group the
        members are split over the ranks and the trained ensemble is all-gathered (sharded.train_ensemble_sharded).
        Returns the losses (epochs, n_M)."""
        if self.regnn is None or len(self.regnn) == 0:
            raise ValueError("Regression model isn't defined.")
        from .sharded import train_ensemble_sharded
        m0 = self.regnn[0]
        for m in self.regnn:
            m.set_normalize_bounds(X_v_train)
        X_n = normalize_host(X_v_train, m0.norm, m0.norm_bounds)

        def train_fn(members):
            tr = EnsembleTrainer(self.ctx, members, m0.layers, X_n, v_train, m0.lr, m0.lam, m0.adv_eps, m0.soft_0,
                                 use_graph=use_graph)
            losses = tr.run(epochs)
            return tr.weights(), losses
        weights, losses = train_ensemble_sharded(train_fn, [m.weights for m in self.regnn], group)
        for m, w in zip(self.regnn, weights):
            m.weights = w
        return losses

    def _ensemble_args(self):
        if not self.regnn:
            raise ValueError("Regression model isn't defined.")      # podnnmodel.py:295
        m0 = self.regnn[0]
        return [m.weights for m in self.regnn], m0.layers, m0.soft_0, m0.norm, m0.norm_bounds

    def predict_v(self, X_v):
        """Gaussian-mixture moments over the members, podnnmodel.py:314-328 -> (v_pred, v_pred_sig)."""
        members, layers, soft_0, norm, nb = self._ensemble_args()
        X_v = np.asarray(X_v, dtype=np.float64).reshape(-1, layers[0])
        if self.n_gpus > 1 and X_v.shape[0] >= 4096 * self.n_gpus:
            # the points are independent: contiguous batch slices, one per GPU, launched back to back (the kernels of
            # the different GPUs run concurrently), gathered on the host
            m = self.multi
            rr = [m.row_range(X_v.shape[0], p) for p in range(m.n_gpus)]
            parts = [ensemble_predict_device(m.ctxs[p], members, layers, X_v[lo:hi], soft_0, norm, nb)
                     for p, (lo, hi) in enumerate(rr)]
            return np.vstack([a.download() for a, _ in parts]), np.vstack([b.download() for _, b in parts])
        vm, vs = ensemble_predict_device(self.ctx, members, layers, X_v, soft_0, norm, nb)
        return vm.download(), vs.download()

    def predict(self, X_v, samples=100, *, monte_carlo=False, seed=0):
        """podnnmodel.py:366-379.  The reference averages `samples` TF-RNG draws of
        v ~ N(v_pred, v_pred_sig).  By default their exact moments U_pred = V v_pred^T,
        U_pred_sig = sqrt((V*V) (v_pred_sig^2)^T) are returned (the limit samples -> inf);
        `monte_carlo=True` runs the reference's estimator itself on the GPU (Philox draws keyed by
        `seed`, running sum / sum of squares fused into the V v^T tiles, unbiased sample sigma)."""
        print(f"Averaging {samples} model configurations...")
        ctx, Vd = self.ctx, self._ensure_V_dev()
        members, layers, soft_0, norm, nb = self._ensemble_args()
        vm, vs = ensemble_predict_device(ctx, members, layers, X_v, soft_0, norm, nb)
        N = vm.shape[0]
        U_pred, U_pred_sig = np.empty((self.n_h, N)), np.empty((self.n_h, N))
        # the (n_h x N) outputs are streamed to the host in column tiles so that the device buffers stay bounded
        # (C4 shape: n_h = 75 000 -- 100 000 points would need 2 x 60 GB of HBM in one piece)
        for j0, j1, Um, Us, t in self._u_tiles(N):
            ctx.check(_lib.lib().pb_predict_U(ctx.h, Vd.ptr, Vd.ld, self.n_h, self.n_L, vm.ptr + 8 * j0 * self.n_L,
                                              vs.ptr + 8 * j0 * self.n_L, j1 - j0, int(samples) if monte_carlo else 0,
                                              int(seed) + 7919 * t, Um.ptr, Us.ptr, Um.ld))
            U_pred[:, j0:j1], U_pred_sig[:, j0:j1] = Um.download(), Us.download()
        return U_pred, U_pred_sig

    U_TILE_BYTES = 4 << 30          # per device output buffer of predict / predict_mc

    def _u_tiles(self, N):
        """Column tiles [j0, j1) of an (n_h x N) U-level output with a pair of device buffers for each tile."""
        cols = int(max(1, min(N, self.U_TILE_BYTES // (8 * self.n_h))))
        bufs = {}
        for t, j0 in enumerate(range(0, N, cols)):
            j1 = min(N, j0 + cols)
            if (j1 - j0) not in bufs:
                bufs.clear()
                bufs[j1 - j0] = (self.ctx.alloc(self.n_h, j1 - j0), self.ctx.alloc(self.n_h, j1 - j0))
            Um, Us = bufs[j1 - j0]
            yield j0, j1, Um, Us, t

    def predict_dist(self, X_v, model_i, samples=100, *, monte_carlo=False, seed=0):
        """podnnmodel.py:330-342: the U-level distribution of ONE member, (U_pred, U_pred_sig).  As in `predict`: the
        exact moments of the reference's Monte-Carlo loop by default, the loop itself (Philox draws) on request."""
        ctx, Vd = self.ctx, self._ensure_V_dev()
        members, layers, soft_0, norm, nb = self._ensemble_args()
        Xd = ctx.to_device(np.asarray(X_v, dtype=np.float64).reshape(-1, layers[0]))
        vm, vs = ensemble_predict_device(ctx, [members[model_i]], layers, Xd, soft_0, norm, nb)
        N = Xd.shape[0]
        U_pred, U_pred_sig = np.empty((self.n_h, N)), np.empty((self.n_h, N))
        for j0, j1, Um, Us, t in self._u_tiles(N):
            ctx.check(_lib.lib().pb_predict_U(ctx.h, Vd.ptr, Vd.ld, self.n_h, self.n_L, vm.ptr + 8 * j0 * self.n_L,
                                              vs.ptr + 8 * j0 * self.n_L, j1 - j0, int(samples) if monte_carlo else 0,
                                              int(seed) + int(model_i) + 7919 * t, Um.ptr, Us.ptr, Um.ld))
            U_pred[:, j0:j1], U_pred_sig[:, j0:j1] = Um.download(), Us.download()
        return U_pred, U_pred_sig

    def predict_mc(self, X_v, samples=100, *, monte_carlo=False, seed=0):
        """podnnmodel.py:344-364: per-member U-level prediction (predict_dist :330-342), Gaussian mixture
        over the members at the U level, plus pod_sig.  Per member the exact moments of its Monte-Carlo
        loop are used unless `monte_carlo=True` (Philox draws on the GPU)."""
        ctx, Vd = self.ctx, self._ensure_V_dev()
        members, layers, soft_0, norm, nb = self._ensemble_args()
        Xd = ctx.to_device(np.asarray(X_v, dtype=np.float64).reshape(-1, layers[0]))
        N = Xd.shape[0]
        print(f"Ensembling {len(members)} predictions...")
        L = _lib.lib()
        v_members = [ensemble_predict_device(ctx, [w], layers, Xd, soft_0, norm, nb) for w in members]
        sig_d = ctx.to_device(np.asarray(self.pod_sig, dtype=np.float64)) if self.pod_sig is not None else None
        U_pred, U_pred_sig = np.empty((self.n_h, N)), np.empty((self.n_h, N))
        acc = {}
        # the member mixture stays in HBM: per column tile the member moments are accumulated on the device and only
        # the finished U_pred / U_pred_sig tiles cross PCIe
        for j0, j1, Um, Us, t in self._u_tiles(N):
            if (j1 - j0) not in acc:
                acc.clear()
                acc[j1 - j0] = (ctx.alloc(self.n_h, j1 - j0), ctx.alloc(self.n_h, j1 - j0))
            S1, S2 = acc[j1 - j0]
            for i, (vm, vs) in enumerate(v_members):
                ctx.check(L.pb_predict_U(ctx.h, Vd.ptr, Vd.ld, self.n_h, self.n_L, vm.ptr + 8 * j0 * self.n_L,
                                         vs.ptr + 8 * j0 * self.n_L, j1 - j0, int(samples) if monte_carlo else 0,
                                         int(seed) + i + 7919 * t, Um.ptr, Us.ptr, Um.ld))
                ctx.check(L.pb_mixture_accumulate(ctx.h, Um.ptr, Us.ptr, Um.ld, self.n_h, j1 - j0, S1.ptr, S2.ptr, S1.ld,
                                                  1 if i == 0 else 0))
            ctx.check(L.pb_mixture_finalize(ctx.h, S1.ptr, S2.ptr, S1.ld, self.n_h, j1 - j0, len(members),
                                            sig_d.ptr if sig_d is not None else None))
            U_pred[:, j0:j1], U_pred_sig[:, j0:j1] = S1.download(), S2.download()
        return U_pred, U_pred_sig

    # ------------------------------------------------------------ layout helpers (host index maps)
    def get_u_tuple(self, n_t=None):
        n_t = self.n_t if n_t is None else n_t
        return (self.n_v, self.n_xyz) + ((n_t,) if self.has_t else ())

    def restruct(self, U, no_s=False, n_t=None):
        """podnnmodel.py:382-402."""
        if no_s:
            return U.reshape(self.get_u_tuple())
        if self.has_t:
            n_t = self.n_t if n_t is None else n_t
            n_s = int(U.shape[-1] / n_t)
            return np.ascontiguousarray(U.reshape(self.n_v, self.n_xyz, n_s, n_t).transpose(0, 1, 3, 2))
        return U.reshape(self.n_v, self.n_xyz, U.shape[-1]).copy()

    def destruct(self, U_struct):
        """podnnmodel.py:404-421."""
        n_s = int(U_struct.shape[-1])
        if self.has_t:
            return np.ascontiguousarray(U_struct.transpose(0, 1, 3, 2)).reshape(self.n_h, n_s * self.n_t)
        return U_struct.reshape(self.n_h, n_s).copy()

    # ------------------------------------------------------------ caches (same tuples as the reference)
    def load_train_data(self):
        if not os.path.exists(self.train_data_path):
            raise FileNotFoundError("Can't find train data.")
        with open(self.train_data_path, "rb") as f:
            data = pickle.load(f)
        self.n_L, self.n_d, self.V, self.pod_sig = data[0], data[1], data[2], data[3]
        self._V_dev = None
        print(f"Mean pod sig: {self.pod_sig.mean()}")
        return data[4:]

    def save_train_data(self, X_v_train, v_train, U_train, X_v_val, v_val, U_val):
        with open(self.train_data_path, "wb") as f:
            pickle.dump((self.n_L, self.n_d, self.V, self.pod_sig,
                         X_v_train, v_train, U_train, X_v_val, v_val, U_val), f)

    def load_init_data(self):
        if not os.path.exists(self.init_data_path):
            raise FileNotFoundError("Can't find train data.")
        with open(self.init_data_path, "rb") as f:
            return pickle.load(f)

    def save_init_data(self, X_v_train, v_train, U_train, X_v_val, v_val, U_val):
        with open(self.init_data_path, "wb") as f:
            pickle.dump((X_v_train, v_train, U_train, X_v_val, v_val, U_val), f)

    def save_setup_data(self):
        os.makedirs(self.resdir, exist_ok=True)
        with open(self.setup_data_path, "wb") as f:
            pickle.dump((self.n_v, self.x_mesh, self.n_t), f)

    @classmethod
    def load_setup_data(cls, save_dir):
        path = os.path.join(save_dir, SETUP_DATA_NAME)
        if not os.path.exists(path):
            raise FileNotFoundError("Can't find setup data.")
        with open(path, "rb") as f:
            return pickle.load(f)

    def save_model(self, model_id=-1):
        ids = range(len(self.regnn)) if model_id < 0 else [model_id]
        for i in ids:
            self.regnn[i].save_to(self.model_path[i], self.model_params_path)

    def load_model(self):
        if not all(os.path.exists(p + MODEL_NAME_EXT) for p in self.model_path):
            raise FileNotFoundError("Can't find cached model.")
        if not os.path.exists(self.model_params_path):
            raise FileNotFoundError("Can't find cached model params.")
        self.regnn = [VarNeuralNetwork.load_from(p, self.model_params_path) for p in self.model_path]

    @classmethod
    def load(cls, save_dir, **kw):
        n_v, x_mesh, n_t = cls.load_setup_data(save_dir)
        paths = [os.path.join(save_dir, f[:-len(MODEL_NAME_EXT)]) for f in sorted(os.listdir(save_dir))
                 if f.startswith(MODEL_NAME) and f.endswith(MODEL_NAME_EXT)]
        model = cls(save_dir, n_v, x_mesh, n_t, **kw)
        model.model_path = paths
        model.load_train_data()
        model.load_model()
        return model
