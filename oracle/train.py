"""numpy restatement of the deep-ensemble training step.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED against the
reference: TensorFlow 2.1 / TensorFlow-Probability 0.9 are not installed, so
the step below follows the cited reference lines and the published definitions
of the TF ops they call; its gradients are cross-checked against an independent
automatic differentiation (torch.autograd, fp64) in tests/test_train_oracle.py.

Reference path: ``VarNeuralNetwork.grad`` (varneuralnetwork.py:94-108),
``tf_optimization_step`` (:121-128), ``tf_optimization`` (:113-119), ``fit``
(:130-145), ``regularization`` (:88-92); optimizer ``tf.keras.optimizers.Adam(lr)``
(:24) with the Keras defaults beta_1 = 0.9, beta_2 = 0.999, epsilon = 1e-7.
"""
import numpy as np

from .ensemble import softplus

LOG_SQRT_2PI = 0.5 * np.log(2. * np.pi)
ADAM_BETA1 = 0.9
ADAM_BETA2 = 0.999
ADAM_EPS = 1e-7


def forward(X, weights):
    """Dense+ReLU stack, linear head (varneuralnetwork.py:42-53) -> (activations, head t)."""
    acts = [X]
    h = X
    n_layers = len(weights) // 2
    for l in range(n_layers - 1):
        h = np.maximum(h.dot(weights[2*l]) + weights[2*l+1], 0.)
        acts.append(h)
    t = h.dot(weights[-2]) + weights[-1]
    return acts, t


def nll_and_head_grad(t, v, n_L, soft_0):
    """sum(-Normal(loc, scale).log_prob(v)) (varneuralnetwork.py:56-59,100) and d/dt.

    log_prob = -0.5 ((v - loc)/scale)^2 - log(scale) - 0.5 log(2 pi);
    scale = softplus(soft_0 t_s) + 1e-6, d scale / d t_s = soft_0 sigmoid(soft_0 t_s).
    """
    loc = t[:, :n_L]
    z = soft_0 * t[:, n_L:]
    scale = softplus(z) + 1e-6
    r = (v - loc) / scale
    nll = np.sum(0.5 * r * r + np.log(scale) + LOG_SQRT_2PI)
    d_loc = -r / scale
    d_scale = (1. - r * r) / scale
    sig = np.where(z >= 0, 1. / (1. + np.exp(-np.abs(z))), np.exp(-np.abs(z)) / (1. + np.exp(-np.abs(z))))
    d_t = np.concatenate([d_loc, d_scale * soft_0 * sig], axis=1)
    return nll, d_t


def backward(acts, d_t, weights):
    """Reverse pass of the stack: gradients w.r.t. every kernel/bias and w.r.t. the input."""
    n_layers = len(weights) // 2
    grads = [None] * (2 * n_layers)
    delta = d_t
    for l in range(n_layers - 1, -1, -1):
        grads[2*l] = acts[l].T.dot(delta)
        grads[2*l+1] = delta.sum(0)
        delta = delta.dot(weights[2*l].T)
        if l > 0:
            delta = delta * (acts[l] > 0.)
    return grads, delta


def l2_reg(weights, lam):
    """lam * sum_k tf.nn.l2_loss(var_k) = lam * sum_k sum(var_k**2)/2 over kernels AND biases
    (varneuralnetwork.py:88-92; model.trainable_variables holds both)."""
    return lam * sum(0.5 * np.sum(w * w) for w in weights)


def loss_and_grads(X, v, weights, n_L, lam, adv_eps, soft_0):
    """VarNeuralNetwork.grad (varneuralnetwork.py:94-108).

    loss = NLL(X) + reg; if adv_eps is not None (0.0 included, as in the reference's
    ``is not None`` test): X_adv = X + adv_eps sign(d loss / d X), loss += NLL(X_adv) + reg.
    The gradient does not flow through sign().  Returns (loss, grads, X_adv or None).
    """
    acts, t = forward(X, weights)
    nll, d_t = nll_and_head_grad(t, v, n_L, soft_0)
    grads, d_X = backward(acts, d_t, weights)
    reg = l2_reg(weights, lam)
    loss = nll + reg
    n_reg = 1
    X_adv = None
    if adv_eps is not None:
        X_adv = X + adv_eps * np.sign(d_X)
        acts2, t2 = forward(X_adv, weights)
        nll2, d_t2 = nll_and_head_grad(t2, v, n_L, soft_0)
        grads2, _ = backward(acts2, d_t2, weights)
        grads = [g + g2 for g, g2 in zip(grads, grads2)]
        loss = loss + nll2 + reg
        n_reg = 2
    grads = [g + n_reg * lam * w for g, w in zip(grads, weights)]
    return loss, grads, X_adv


def adam_init(weights):
    return {"t": 0, "m": [np.zeros_like(w) for w in weights], "v": [np.zeros_like(w) for w in weights]}


def adam_apply(weights, grads, state, lr):
    """Keras Adam (TF 2.1 ``ResourceApplyAdam``): lr_t = lr sqrt(1-b2^t)/(1-b1^t);
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; w -= lr_t m / (sqrt(v) + eps)."""
    state["t"] += 1
    t = state["t"]
    lr_t = lr * np.sqrt(1. - ADAM_BETA2 ** t) / (1. - ADAM_BETA1 ** t)
    out = []
    for k, (w, g) in enumerate(zip(weights, grads)):
        state["m"][k] = ADAM_BETA1 * state["m"][k] + (1. - ADAM_BETA1) * g
        state["v"][k] = ADAM_BETA2 * state["v"][k] + (1. - ADAM_BETA2) * g * g
        out.append(w - lr_t * state["m"][k] / (np.sqrt(state["v"][k]) + ADAM_EPS))
    return out


def train(X, v, weights, n_L, epochs, lr, lam, adv_eps, soft_0, state=None):
    """tf_optimization (varneuralnetwork.py:113-119) on already-normalised inputs.

    Returns (weights, losses[epochs], state); losses[e] is the loss BEFORE the update of
    epoch e, as the reference's tf_optimization_step returns it."""
    if state is None:
        state = adam_init(weights)
    losses = np.zeros(epochs)
    for e in range(epochs):
        losses[e], grads, _ = loss_and_grads(X, v, weights, n_L, lam, adv_eps, soft_0)
        weights = adam_apply(weights, grads, state, lr)
    return weights, losses, state
