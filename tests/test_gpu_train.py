"""GPU: ensemble training step (csrc/train.cu through pb_ensemble_train) against the oracle's restatement of
VarNeuralNetwork.grad + Keras Adam (oracle/train.py; reference varneuralnetwork.py:94-128).

Tolerance: fp64 throughout; the GPU sums products in a different order (DMMA tiles, split-K partials), so losses
agree to 1e-11 relative and parameters after a few steps to 1e-9 of the largest entry (Adam's m/sqrt(v) amplifies
rounding differences of tiny gradients in the first steps)."""
import numpy as np
import pytest

from oracle import ensemble as ens
from oracle import train as tr
from pod_uqnn_b200.varneuralnetwork import EnsembleTrainer, VarNeuralNetwork, NORM_MEANSTD, normalize_host
from pod_uqnn_b200.logger import Logger

pytestmark = pytest.mark.gpu

TOL_LOSS = 1e-11
TOL_W = 1e-9


def make_problem(layers, N, M, seed, bias_scale=0.05):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1.5, 1.5, (N, layers[0]))
    v = rng.standard_normal((N, layers[-1]))
    members = [ens.init_weights(layers, seed=100 * seed + i, bias_scale=bias_scale) for i in range(M)]
    return X, v, members


def check_against_oracle(ctx, layers, N, M, epochs, lr, lam, adv_eps, soft_0, seed, use_graph=True):
    X, v, members = make_problem(layers, N, M, seed)
    t = EnsembleTrainer(ctx, members, layers, X, v, lr, lam, adv_eps, soft_0, use_graph=use_graph)
    losses = t.run(epochs)
    got = t.weights()
    for i in range(M):
        w_ref, l_ref, _ = tr.train(X, v, [a.copy() for a in members[i]], layers[-1], epochs, lr, lam, adv_eps, soft_0)
        np.testing.assert_allclose(losses[:, i], l_ref, rtol=TOL_LOSS)
        for a, b in zip(got[i], w_ref):
            assert np.abs(a - b).max() <= TOL_W * max(1., np.abs(b).max()), (i, np.abs(a - b).max())
    return losses


@pytest.mark.parametrize("adv_eps", [None, 0., 0.01])
def test_step_matches_oracle_reference_shape(ctx, adv_eps):
    """The reference's architecture (h_layers [128,128,128], 2d_ackley/hyperparams.py:30) at its training size."""
    check_against_oracle(ctx, [3, 128, 128, 128, 14], 400, 3, 4, 1e-3, 1e-3, adv_eps, 0.01, seed=1)


def test_step_matches_oracle_burgers_shape(ctx):
    """1dt_burger/hyperparams.py:28-35: lr 0.01, lam 1e-8, adv_eps 0.01, soft_0 0.01, inputs (t, mu)."""
    check_against_oracle(ctx, [2, 128, 128, 128, 37], 1000, 2, 3, 0.01, 1e-8, 0.01, 0.01, seed=2)


def test_ragged_and_odd_widths(ctx):
    """Widths that are not multiples of the tiles / of 2 (unaligned staging path), one hidden layer, N < one tile."""
    check_against_oracle(ctx, [1, 33, 17, 5], 77, 2, 4, 0.01, 1e-3, 0.02, 1., seed=3)
    check_against_oracle(ctx, [5, 40, 3], 19, 1, 3, 0.01, 0., None, 1., seed=4)
    check_against_oracle(ctx, [2, 1], 9, 2, 3, 0.01, 1e-2, 0.1, 1., seed=5)          # head only
    # odd layer width followed by even ones, odd N: buffer segments of the later layers must stay 16-byte aligned
    check_against_oracle(ctx, [2, 33, 64, 4], 77, 1, 3, 0.01, 1e-3, 0.05, 1., seed=12)
    check_against_oracle(ctx, [3, 15, 128, 9], 401, 3, 3, 0.01, 1e-3, 0.05, 1., seed=13)


def test_wide_layers_take_layer_by_layer_path(ctx):
    """Widths above 128 do not fit the fused kernel's single 128-column chunk: per-layer kernels + forked dW graph
    branch (the PB_TRAIN_FUSED=0 path)."""
    check_against_oracle(ctx, [3, 160, 136, 10], 300, 2, 3, 1e-3, 1e-3, 0.01, 0.5, seed=10)
    check_against_oracle(ctx, [2, 64, 70], 130, 2, 3, 1e-3, 1e-3, None, 0.5, seed=11)      # head 2 n_L = 140 > 128


def test_large_batch_tiles(ctx):
    """Enough rows for the 128-row tiles and 16 split-K partials."""
    check_against_oracle(ctx, [2, 128, 128, 128, 20], 6000, 8, 2, 1e-3, 1e-4, 0.01, 0.01, seed=6)


def test_graph_equals_eager_and_resume(ctx):
    layers, N, M = [3, 64, 64, 8], 300, 4
    X, v, members = make_problem(layers, N, M, 7)
    a = EnsembleTrainer(ctx, members, layers, X, v, 0.01, 1e-4, 0.01, 0.01, use_graph=True)
    la = a.run(7)
    b = EnsembleTrainer(ctx, members, layers, X, v, 0.01, 1e-4, 0.01, 0.01, use_graph=False)
    lb = b.run(7)
    np.testing.assert_array_equal(la, lb)
    for wa, wb in zip(a.weights(), b.weights()):
        for x, y in zip(wa, wb):
            np.testing.assert_array_equal(x, y)
    # 3 + 4 epochs with the optimiser state carried over == 7 epochs
    c = EnsembleTrainer(ctx, members, layers, X, v, 0.01, 1e-4, 0.01, 0.01)
    lc = np.concatenate([c.run(3), c.run(4)])
    np.testing.assert_array_equal(lc, la)
    st = c.adam_state()
    assert st["t"] == 7
    # a fresh trainer resumed from the downloaded state continues identically
    d = EnsembleTrainer(ctx, c.weights(), layers, X, v, 0.01, 1e-4, 0.01, 0.01, adam_state=st)
    e_ref = a.run(2)
    np.testing.assert_array_equal(d.run(2), e_ref)


def test_members_are_independent(ctx):
    """Training M members in one batch equals training each alone."""
    layers, N = [2, 48, 48, 6], 150
    X, v, members = make_problem(layers, N, 3, 8)
    allm = EnsembleTrainer(ctx, members, layers, X, v, 0.01, 1e-3, 0.0, 1.)
    l_all = allm.run(5)
    for i in range(3):
        one = EnsembleTrainer(ctx, [members[i]], layers, X, v, 0.01, 1e-3, 0.0, 1.)
        np.testing.assert_array_equal(one.run(5)[:, 0], l_all[:, i])


def test_fit_learns_and_logs(ctx, capsys):
    """VarNeuralNetwork.fit (varneuralnetwork.py:130-145): normalisation bounds from the data, loss goes down, the
    logger prints at epochs 0, freq, 2 freq, ... and at the end, and predict() uses the trained weights."""
    rng = np.random.default_rng(0)
    X = rng.uniform(800., 1200., (256, 1))
    v = np.stack([np.sin((X[:, 0] - 1000.) / 80.), ((X[:, 0] - 1000.) / 200.) ** 2], 1)
    net = VarNeuralNetwork([1, 64, 64, 2], lr=0.01, lam=1e-6, adv_eps=0.01, soft_0=0.1, norm=NORM_MEANSTD, seed=3)
    logger = Logger(200, 50)
    seen = []
    logger.set_val_err_fn(lambda: seen.append(1) or {"RE_val": 0., "MPIW_val": 0.})
    net.fit(X, v, 200, logger, ctx=ctx)
    out = capsys.readouterr().out
    assert out.count("#:") == 5                       # epochs 0, 50, 100, 150 and the closing line (epoch 200)
    assert len(seen) == 5
    mean, var = net.predict(X, ctx=ctx)
    assert np.mean((mean - v) ** 2) < 0.05
    assert np.all(var > 0)
    # same trajectory as the oracle on the normalised inputs
    net2 = VarNeuralNetwork([1, 64, 64, 2], lr=0.01, lam=1e-6, adv_eps=0.01, soft_0=0.1, norm=NORM_MEANSTD, seed=3)
    w0 = [a.copy() for a in net2.weights]
    net2.fit_simple(X, v, 6, ctx=ctx)
    Xn = normalize_host(X, NORM_MEANSTD, (X.mean(0), X.std(0)))
    w_ref, _, _ = tr.train(Xn, v, w0, 2, 6, 0.01, 1e-6, 0.01, 0.1)
    for a, b in zip(net2.weights, w_ref):
        assert np.abs(a - b).max() <= TOL_W * max(1., np.abs(b).max())


def test_bad_arguments(ctx):
    X, v, members = make_problem([2, 8, 3], 10, 1, 9)
    with pytest.raises(ValueError):
        EnsembleTrainer(ctx, members, [2, 8, 3], X, v[:, :2], 0.01, 0., None, 1.)
    t = EnsembleTrainer(ctx, members, [2, 8, 3], X, v, 0.01, 0., None, 1.)
    assert t.run(0).shape == (0, 1)
