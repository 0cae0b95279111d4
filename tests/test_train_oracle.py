"""CPU: the training-step oracle (oracle/train.py) against an independent automatic differentiation.

The reference's step (varneuralnetwork.py:94-128) needs TensorFlow, which is absent; the numpy
restatement is cross-checked here against torch.autograd in fp64 built from the same published op
definitions (Dense, ReLU, softplus, Normal.log_prob, l2_loss, sign, Keras Adam)."""
import math

import numpy as np
import pytest
import torch

from oracle import ensemble as ens
from oracle import train as tr


def torch_loss(X, v, ws, n_L, lam, adv_eps, soft_0):
    def model(Xt):
        h = Xt
        for l in range(len(ws) // 2 - 1):
            h = torch.relu(h @ ws[2*l] + ws[2*l+1])
        t = h @ ws[-2] + ws[-1]
        return torch.distributions.Normal(t[:, :n_L], torch.nn.functional.softplus(soft_0 * t[:, n_L:]) + 1e-6)

    def reg():
        return lam * sum(0.5 * (w * w).sum() for w in ws)

    Xt = X.clone().requires_grad_(True)
    loss = (-model(Xt).log_prob(v)).sum() + reg()
    X_adv = None
    if adv_eps is not None:
        loss_x, = torch.autograd.grad(loss, Xt, retain_graph=True)
        X_adv = (Xt + adv_eps * torch.sign(loss_x)).detach()
        loss = loss + (-model(X_adv).log_prob(v)).sum() + reg()
    grads = torch.autograd.grad(loss, ws)
    return loss, grads, X_adv


CASES = [
    # layers,               N,  lam,   adv_eps, soft_0
    ([3, 16, 16, 4],        37, 1e-3,  0.,      0.01),
    ([2, 32, 32, 32, 7],    50, 1e-8,  0.01,    0.01),
    ([1, 24, 5],            21, 1e-2,  None,    1.),
    ([10, 128, 128, 128, 9], 64, 1e-3, 0.05,    1.),
]


@pytest.mark.parametrize("layers,N,lam,adv_eps,soft_0", CASES)
def test_grad_matches_autograd(layers, N, lam, adv_eps, soft_0):
    rng = np.random.default_rng(11)
    w = ens.init_weights(layers, seed=5, bias_scale=0.1)
    n_L = layers[-1]
    X = rng.standard_normal((N, layers[0]))
    v = rng.standard_normal((N, n_L))
    loss, grads, X_adv = tr.loss_and_grads(X, v, w, n_L, lam, adv_eps, soft_0)
    ws = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in w]
    tl, tg, tX_adv = torch_loss(torch.tensor(X), torch.tensor(v), ws, n_L, lam, adv_eps, soft_0)
    assert abs(loss - tl.item()) <= 1e-12 * abs(tl.item())
    if adv_eps is not None:
        np.testing.assert_array_equal(X_adv, tX_adv.numpy())
    for g, t in zip(grads, tg):
        t = t.numpy()
        assert np.abs(g - t).max() <= 1e-11 * max(1., np.abs(t).max())


def test_adam_matches_keras_formula_and_torch():
    """Keras' update equals torch.optim.Adam's with eps rescaled by sqrt(1 - beta2^t); check both the closed form of
    the first step (w -= lr sign(g) up to eps) and three steps against a hand-rolled torch loop."""
    rng = np.random.default_rng(3)
    w = [rng.standard_normal((4, 3)), rng.standard_normal(3)]
    state = tr.adam_init(w)
    ws = [a.copy() for a in w]
    m = [np.zeros_like(a) for a in w]; v = [np.zeros_like(a) for a in w]
    for t in range(1, 4):
        g = [rng.standard_normal(a.shape) for a in w]
        ws_new = tr.adam_apply(ws, g, state, lr=0.01)
        for k in range(2):
            m[k] = 0.9 * m[k] + 0.1 * g[k]
            v[k] = 0.999 * v[k] + 0.001 * g[k]**2
            mhat = m[k] / (1 - 0.9**t); vhat = v[k] / (1 - 0.999**t)
            # Keras: eps is added to sqrt(v) (not sqrt(vhat)) -> eps_hat = eps / sqrt(1 - beta2^t)
            ref = ws[k] - 0.01 * mhat / (np.sqrt(vhat) + 1e-7 / math.sqrt(1 - 0.999**t))
            np.testing.assert_allclose(ws_new[k], ref, rtol=1e-13, atol=1e-15)
        ws = ws_new
    assert state["t"] == 3


def test_training_reduces_loss_and_is_deterministic():
    layers = [2, 32, 32, 3]
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (80, 2))
    v = np.stack([np.sin(X[:, 0]), X[:, 0] * X[:, 1], np.cos(X[:, 1])], 1)
    w0 = ens.init_weights(layers, seed=1)
    w1, losses, st = tr.train(X, v, [a.copy() for a in w0], 3, 60, 0.01, 1e-6, 0.01, 1.)
    w2, losses2, _ = tr.train(X, v, [a.copy() for a in w0], 3, 60, 0.01, 1e-6, 0.01, 1.)
    assert losses[-1] < 0.8 * losses[0]
    np.testing.assert_array_equal(losses, losses2)
    assert st["t"] == 60
    # resuming from the saved optimiser state continues the same trajectory
    wa, la, sa = tr.train(X, v, [a.copy() for a in w0], 3, 25, 0.01, 1e-6, 0.01, 1.)
    wb, lb, _ = tr.train(X, v, wa, 3, 35, 0.01, 1e-6, 0.01, 1., state=sa)
    np.testing.assert_array_equal(np.concatenate([la, lb]), losses)
    for a, b in zip(wb, w1):
        np.testing.assert_array_equal(a, b)
