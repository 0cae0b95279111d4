"""Mean/variance MLP -- forward path of the reference's poduqnn/varneuralnetwork.py
(build_model :42-62, normalize :75-86, predict :154-160) on the GPU (csrc/mlp.cu).

Training (grad :94-108, Adam step :121-128, fit :130-152) runs on the GPU too (csrc/train.cu), all
members of an ensemble in the same launches (`EnsembleTrainer`).  Weights use the
Keras ``model.get_weights()`` order ``[W0, b0, W1, b1, ...]`` with kernels shaped (in, out);
`soft_0` is explicit (the reference's load_from drops it, varneuralnetwork.py:191-193).
"""
import ctypes as C
import pickle

import numpy as np

from . import _lib
from ._lib import NORM_IDS

NORM_NONE = "none"
NORM_MEANSTD = "meanstd"
NORM_CENTER = "center"


def glorot_normal(rng, fan_in, fan_out):
    """Keras glorot_normal: truncated normal, std = sqrt(2/(fan_in+fan_out)) / 0.8796..."""
    std = np.sqrt(2. / (fan_in + fan_out)) / .87962566103423978
    z = rng.standard_normal((fan_in, fan_out))
    bad = np.abs(z) > 2.
    while bad.any():
        z[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(z) > 2.
    return z * std


def pack_members(members):
    """[[W0,b0,...], ...] -> one flat float64 array in the pb_ensemble_predict layout."""
    return np.concatenate([np.concatenate([np.ascontiguousarray(a, dtype=np.float64).ravel() for a in w])
                           for w in members])


def ensemble_predict_device(ctx, members, layers, X, soft_0=1., norm=NORM_NONE, norm_bounds=None,
                            packed_dev=None):
    """Mixture moments of an ensemble on the device: returns (v_mean_dev, v_sig_dev), (N, n_L) each.

    With one member v_sig is that member's scale (variance = v_sig**2 exactly).
    """
    L = _lib.lib()
    dims = list(layers[:-1]) + [2 * layers[-1]]
    n_layers = len(dims) - 1
    n_L = layers[-1]
    Xd = X if isinstance(X, _lib.DeviceArray) else ctx.to_device(np.asarray(X, dtype=np.float64).reshape(-1, dims[0]))
    N = Xd.shape[0]
    w = packed_dev if packed_dev is not None else ctx.to_device(pack_members(members))
    kind = NORM_IDS[norm] if norm_bounds is not None else 0      # norm_bounds None => identity (:77-78)
    na = nb = None
    if kind:
        na = ctx.to_device(np.broadcast_to(np.asarray(norm_bounds[0], dtype=np.float64), (dims[0],)).copy())
        nb = ctx.to_device(np.broadcast_to(np.asarray(norm_bounds[1], dtype=np.float64), (dims[0],)).copy())
    vm, vs = ctx.alloc(N, n_L), ctx.alloc(N, n_L)
    cdims = (C.c_int * len(dims))(*dims)
    ctx.check(L.pb_ensemble_predict(ctx.h, len(members) if members is not None else packed_dev.n_members,
                                    n_layers, cdims, w.ptr, kind,
                                    na.ptr if na else None, nb.ptr if nb else None, float(soft_0),
                                    Xd.ptr, N, vm.ptr, vs.ptr))
    return vm, vs


def normalize_host(X, norm, norm_bounds):
    """varneuralnetwork.py:75-86 on the host (the reference normalises in numpy before tf.convert_to_tensor)."""
    X = np.asarray(X, dtype=np.float64)
    if norm_bounds is None:
        return X
    if norm == NORM_CENTER:
        lb, ub = norm_bounds
        return (X - lb) - 0.5 * (ub - lb)
    if norm == NORM_MEANSTD:
        mean, std = norm_bounds
        return (X - mean) / std
    return X


class EnsembleTrainer:
    """Full-batch training of M members at once: VarNeuralNetwork.grad + tf_optimization_step
    (varneuralnetwork.py:94-128) per epoch, Keras Adam state kept in HBM between `run()` calls.

    `X_norm` (N, n_d) is the normalised input, `v` (N, n_L) the targets; every member sees the same data (as in
    the reference, where the members differ by their glorot initialisation only).  `adv_eps=None` disables the
    adversarial term; 0.0 keeps the reference's doubled loss (`adv_eps is not None`, :101).
    """

    def __init__(self, ctx, members, layers, X_norm, v, lr, lam, adv_eps, soft_0, *, use_graph=True,
                 adam_state=None):
        self.ctx, self.layers = ctx, list(layers)
        self.dims = self.layers[:-1] + [2 * self.layers[-1]]
        self.M = len(members)
        self.lr, self.lam, self.adv_eps, self.soft_0 = float(lr), float(lam), adv_eps, float(soft_0)
        self.use_graph = bool(use_graph)
        X_norm = np.ascontiguousarray(np.asarray(X_norm, dtype=np.float64).reshape(-1, self.dims[0]))
        v = np.ascontiguousarray(v, dtype=np.float64)
        if v.shape != (X_norm.shape[0], self.layers[-1]):
            raise ValueError(f"targets {v.shape} do not match inputs {X_norm.shape} and n_L = {self.layers[-1]}")
        self.N = X_norm.shape[0]
        self.P = sum(a * b + b for a, b in zip(self.dims[:-1], self.dims[1:]))
        packed = pack_members(members).reshape(self.M, self.P)
        self.params = ctx.to_device(packed)
        if adam_state is None:
            self.steps = 0
            self.m = ctx.to_device(np.zeros((self.M, self.P)))
            self.v2 = ctx.to_device(np.zeros((self.M, self.P)))
        else:
            self.steps = int(adam_state["t"])
            self.m = ctx.to_device(np.asarray(adam_state["m"], dtype=np.float64).reshape(self.M, self.P))
            self.v2 = ctx.to_device(np.asarray(adam_state["v"], dtype=np.float64).reshape(self.M, self.P))
        self.X, self.v = ctx.to_device(X_norm), ctx.to_device(v)

    def run(self, epochs):
        """`epochs` optimiser steps; returns the losses (epochs, M), each taken before its update (:124-128)."""
        epochs = int(epochs)
        losses = np.zeros((epochs, self.M))
        if epochs == 0:
            return losses
        cdims = (C.c_int * len(self.dims))(*self.dims)
        self.ctx.check(_lib.lib().pb_ensemble_train(
            self.ctx.h, self.M, len(self.dims) - 1, cdims, self.params.ptr, self.m.ptr, self.v2.ptr, self.steps,
            epochs, self.lr, self.lam, 0 if self.adv_eps is None else 1,
            0. if self.adv_eps is None else float(self.adv_eps), self.soft_0, self.X.ptr, self.v.ptr, self.N,
            losses.ctypes.data, 1 if self.use_graph else 0))
        self.steps += epochs
        return losses

    def weights(self):
        """Current parameters as M Keras-ordered lists [W0, b0, W1, b1, ...]."""
        flat = self.params.download()
        out = []
        for i in range(self.M):
            w, off = [], 0
            for a, b in zip(self.dims[:-1], self.dims[1:]):
                w.append(flat[i, off:off + a * b].reshape(a, b).copy()); off += a * b
                w.append(flat[i, off:off + b].copy()); off += b
            out.append(w)
        return out

    def adam_state(self):
        return {"t": self.steps, "m": self.m.download(), "v": self.v2.download()}


class VarNeuralNetwork:
    """Same constructor signature as the reference (varneuralnetwork.py:17-18)."""

    def __init__(self, layers, lr, lam, adv_eps=None, soft_0=1., norm=NORM_NONE, weights_path=None,
                 norm_bounds=None, seed=None):
        self.dtype = "float64"
        self.layers, self.lr, self.lam = list(layers), lr, lam
        self.adv_eps, self.soft_0, self.norm, self.norm_bounds = adv_eps, soft_0, norm, norm_bounds
        self.logger = None
        rng = np.random.default_rng(seed)
        dims = self.layers[:-1] + [2 * self.layers[-1]]
        self.weights = []
        for fi, fo in zip(dims[:-1], dims[1:]):          # Keras: glorot_normal kernels, zero biases
            self.weights += [glorot_normal(rng, fi, fo), np.zeros(fo)]
        if weights_path is not None:
            self.load_weights(weights_path)

    # --- weights -----------------------------------------------------------
    def set_weights(self, weights):
        dims = self.layers[:-1] + [2 * self.layers[-1]]
        if len(weights) != 2 * (len(dims) - 1):
            raise ValueError("expected [W0, b0, W1, b1, ...]")
        for l, (fi, fo) in enumerate(zip(dims[:-1], dims[1:])):
            if np.shape(weights[2*l]) != (fi, fo) or np.shape(weights[2*l+1]) != (fo,):
                raise ValueError(f"layer {l}: expected kernel {(fi, fo)} and bias {(fo,)}")
        self.weights = [np.asarray(w, dtype=np.float64) for w in weights]

    def get_weights(self):
        return self.weights

    def load_weights(self, path):
        with np.load(path if path.endswith(".npz") else path + ".npz") as z:
            self.set_weights([z[f"arr_{i}"] for i in range(len(z.files))])

    def save_weights(self, path):
        np.savez(path if path.endswith(".npz") else path + ".npz", *self.weights)

    # --- normalisation (varneuralnetwork.py:64-86) ---------------------------
    def set_normalize_bounds(self, X):
        if self.norm == NORM_CENTER:
            self.norm_bounds = (np.amin(X, axis=0), np.amax(X, axis=0))
        elif self.norm == NORM_MEANSTD:
            self.norm_bounds = (X.mean(0), X.std(0))

    def normalize(self, X):
        """varneuralnetwork.py:75-86 (host arrays; the fused kernels normalise on the device)."""
        return normalize_host(X, self.norm, self.norm_bounds)

    def summary(self):
        """varneuralnetwork.py:167-169: the layer table Keras would print (Dense ... Dense(2 n_L), mean / softplus head)."""
        dims = list(self.layers[:-1]) + [2 * self.layers[-1]]
        lines, total = [f"{'Layer':<24}{'Output shape':<16}{'Param #':>10}"], 0
        for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
            n = a * b + b
            total += n
            act = "relu" if i < len(dims) - 2 else "linear (mean | softplus var)"
            lines.append(f"{'dense_' + str(i) + ' (' + act + ')':<24}{'(None, ' + str(b) + ')':<16}{n:>10}")
        lines.append(f"Total params: {total}")
        text = "\n".join(lines)
        print(text)
        return text

    # --- forward -------------------------------------------------------------
    def predict(self, X, *, ctx=None):
        """(mean, variance) of the predictive Normal, as the reference (:154-160)."""
        ctx = ctx or _lib.default_context()
        vm, vs = ensemble_predict_device(ctx, [self.weights], self.layers, X, self.soft_0, self.norm,
                                         self.norm_bounds)
        return vm.download(), vs.download() ** 2

    def predict_dist(self, X, *, ctx=None):
        """(loc, scale) of the predictive Normal (the reference returns a tfd.Normal, :162-165)."""
        ctx = ctx or _lib.default_context()
        vm, vs = ensemble_predict_device(ctx, [self.weights], self.layers, X, self.soft_0, self.norm,
                                         self.norm_bounds)
        return vm.download(), vs.download()

    # --- training (varneuralnetwork.py:113-152) ---------------------------------
    def tf_optimization(self, X_v, v, tf_epochs, nolog=False, *, ctx=None, use_graph=True):
        """:113-119 on normalised inputs.  The logger is called at the epochs where the reference's would print
        (epoch % frequency == 0, with the weights after that epoch's update); the optimiser state persists
        across calls like the reference's tf_optimizer."""
        ctx = ctx or _lib.default_context()
        tr = EnsembleTrainer(ctx, [self.weights], self.layers, X_v, v, self.lr, self.lam, self.adv_eps, self.soft_0,
                             use_graph=use_graph, adam_state=getattr(self, "_adam_state", None))
        freq = None if (nolog or self.logger is None or self.logger.silent) else int(self.logger.frequency)
        done, loss_value = 0, None
        while done < tf_epochs:
            if freq is None:
                n = tf_epochs - done
            else:
                nxt = ((done + freq - 1) // freq) * freq          # next logged epoch >= done
                n = min(nxt, tf_epochs - 1) - done + 1
            losses = tr.run(n)
            done += n
            loss_value = float(losses[-1, 0])
            if freq is not None and (done - 1) % freq == 0:
                self.weights = tr.weights()[0]
                self.logger.log_train_epoch(done - 1, loss_value)
        self.weights = tr.weights()[0]
        self._adam_state = tr.adam_state()
        return loss_value

    def fit(self, X_v, v, epochs, logger, *, ctx=None):
        """:130-145."""
        self.logger = logger
        self.logger.log_train_start()
        self.set_normalize_bounds(X_v)
        X_n = normalize_host(X_v, self.norm, self.norm_bounds)
        last_loss = self.tf_optimization(X_n, v, epochs, ctx=ctx)
        self.logger.log_train_end(epochs, last_loss)

    def fit_simple(self, X_v, v, epochs, *, ctx=None):
        """:147-152."""
        self.set_normalize_bounds(X_v)
        X_n = normalize_host(X_v, self.norm, self.norm_bounds)
        self.tf_optimization(X_n, v, epochs, nolog=True, ctx=ctx)

    def save_to(self, model_path, params_path):
        """Same params tuple as the reference (:175-180); weights as .npz instead of a TF checkpoint."""
        with open(params_path, "wb") as f:
            pickle.dump((self.layers, self.lr, self.lam, self.soft_0, self.norm, self.norm_bounds), f)
        self.save_weights(model_path)

    @classmethod
    def load_from(cls, weights_path, params_path):
        with open(params_path, "rb") as f:
            layers, lr, lam, soft_0, norm, norm_bounds = pickle.load(f)
        # unlike the reference (:191-193) soft_0 is restored
        return cls(layers, lr, lam, soft_0=soft_0, norm=norm, weights_path=weights_path, norm_bounds=norm_bounds)
