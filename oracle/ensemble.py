"""numpy restatement of the deep-ensemble surrogate forward path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: TensorFlow 2.1
and TensorFlow-Probability 0.9 are not installed and there is no network, so
this file is the oracle for the ensemble path; it follows the cited reference
lines one by one.  Weights use the Keras ``model.get_weights()`` order
``[W0, b0, W1, b1, ...]`` with kernels shaped ``(in, out)``.
"""
import numpy as np

NORM_NONE = "none"
NORM_MEANSTD = "meanstd"
NORM_CENTER = "center"


def glorot_normal(rng, fan_in, fan_out):
    """Keras ``glorot_normal``: truncated normal (|z| <= 2), std = sqrt(2/(fan_in+fan_out))
    corrected by 0.87962566103423978 (varneuralnetwork.py:48-53 `kernel_initializer`)."""
    std = np.sqrt(2. / (fan_in + fan_out)) / .87962566103423978
    z = rng.standard_normal((fan_in, fan_out))
    bad = np.abs(z) > 2.
    while bad.any():
        z[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(z) > 2.
    return z * std


def init_weights(layers, seed, bias_scale=0.):
    """Synthetic member weights for ``layers = [n_d, *h_layers, n_L]``.

    The head has ``2 * n_L`` outputs (varneuralnetwork.py:51-53).  Keras biases
    start at zero; ``bias_scale`` > 0 draws small biases so that tests cover the
    bias path a trained network would exercise.
    """
    rng = np.random.default_rng(seed)
    dims = list(layers[:-1]) + [2 * layers[-1]]
    weights = []
    for fan_in, fan_out in zip(dims[:-1], dims[1:]):
        weights.append(glorot_normal(rng, fan_in, fan_out))
        weights.append(bias_scale * rng.standard_normal(fan_out))
    return weights


def set_normalize_bounds(X, norm):
    """varneuralnetwork.py:64-73 (population std)."""
    if norm == NORM_CENTER:
        return (np.amin(X, axis=0), np.amax(X, axis=0))
    if norm == NORM_MEANSTD:
        return (X.mean(0), X.std(0))
    return None


def normalize(X, norm, norm_bounds):
    """varneuralnetwork.py:75-86."""
    if norm_bounds is None:
        return X
    if norm == NORM_CENTER:
        lb, ub = norm_bounds
        X = (X - lb) - 0.5 * (ub - lb)
    elif norm == NORM_MEANSTD:
        mean, std = norm_bounds
        X = (X - mean) / std
    return X


def softplus(x):
    """tf.math.softplus: log(1 + exp(x)), evaluated without overflow."""
    return np.logaddexp(0., x)


def member_predict(X, weights, n_L, soft_0=1., norm=NORM_NONE, norm_bounds=None):
    """VarNeuralNetwork.predict (varneuralnetwork.py:154-160) -> (mean, variance).

    Forward: Dense+ReLU per hidden width, linear head of 2*n_L (:42-53);
    Normal(loc=t[..., :n_L], scale=softplus(soft_0*t[..., n_L:]) + 1e-6) (:56-59);
    mean = loc, variance = scale**2.
    """
    h = normalize(np.asarray(X, dtype=np.float64), norm, norm_bounds)
    n_layers = len(weights) // 2
    for l in range(n_layers - 1):
        h = np.maximum(h.dot(weights[2*l]) + weights[2*l+1], 0.)
    t = h.dot(weights[-2]) + weights[-1]
    loc = t[..., :n_L]
    scale = softplus(soft_0 * t[..., n_L:]) + 1e-6
    return loc, scale**2


def predict_v(X_v, members, n_L, soft_0=1., norm=NORM_NONE, norm_bounds=None):
    """PodnnModel.predict_v (podnnmodel.py:314-328): Gaussian-mixture moments."""
    n_M = len(members)
    v_pred_samples = np.zeros((X_v.shape[0], n_L, n_M))
    v_pred_var_samples = np.zeros((X_v.shape[0], n_L, n_M))
    for i, w in enumerate(members):
        v_pred_samples[:, :, i], v_pred_var_samples[:, :, i] = \
            member_predict(X_v, w, n_L, soft_0, norm, norm_bounds)
    v_pred = v_pred_samples.mean(-1)
    v_pred_var = (v_pred_var_samples + v_pred_samples ** 2).mean(-1) - v_pred ** 2
    v_pred_sig = np.sqrt(v_pred_var)
    return v_pred, v_pred_sig


def predict_U_moments(V, v_pred, v_pred_sig):
    """Closed form of PodnnModel.predict's Monte-Carlo loop (podnnmodel.py:366-379).

    The loop draws v ~ N(v_pred, v_pred_sig) (independent per entry), expands
    U_i = V.dot(v.T) and returns the sample mean / unbiased sample std over the
    draws.  Their expectations are E[U] = V.v_pred^T and
    Var[U] = (V*V).(v_pred_sig**2)^T; the TF RNG stream cannot be reproduced,
    so deterministic parity is defined on these moments and the Monte-Carlo
    version is compared statistically.
    """
    U_pred = V.dot(v_pred.T)
    U_pred_sig = np.sqrt((V * V).dot((v_pred_sig ** 2).T))
    return U_pred, U_pred_sig


def predict_mc(V, v_pred, v_pred_sig, samples, rng):
    """PodnnModel.predict (podnnmodel.py:366-379) with a numpy Generator for the draws."""
    n_h = V.shape[0]
    U_sum = np.zeros((n_h, v_pred.shape[0]))
    U_sum_sq = np.zeros_like(U_sum)
    for _ in range(samples):
        v = v_pred + v_pred_sig * rng.standard_normal(v_pred.shape)
        U_i = V.dot(v.T)
        U_sum += U_i
        U_sum_sq += U_i**2
    U_pred = U_sum / samples
    U_pred_sig = np.sqrt((samples * U_sum_sq - U_sum**2) / (samples * (samples - 1)))
    return U_pred, U_pred_sig


def predict_mc_moments(V, X_v, members, n_L, soft_0, norm, norm_bounds, pod_sig=None):
    """PodnnModel.predict_mc (podnnmodel.py:344-364) with each member's Monte-Carlo loop
    (predict_dist :330-342) replaced by its exact moments."""
    n_M = len(members)
    U_s = np.zeros((V.shape[0], X_v.shape[0], n_M))
    S_s = np.zeros_like(U_s)
    for i, w in enumerate(members):
        mu, var = member_predict(X_v, w, n_L, soft_0, norm, norm_bounds)
        U_s[:, :, i], S_s[:, :, i] = predict_U_moments(V, mu, np.sqrt(var))
    U_pred = U_s.mean(-1)
    U_pred_var = (S_s ** 2 + U_s ** 2).mean(-1) - U_pred ** 2
    U_pred_sig = np.sqrt(U_pred_var)
    if pod_sig is not None:
        U_pred_sig += pod_sig[:, np.newaxis]
    return U_pred, U_pred_sig
