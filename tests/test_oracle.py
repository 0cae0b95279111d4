"""CPU: the numpy oracle against the golden vectors produced by the unmodified reference
(oracle/gen_golden.py), and against the live reference when /root/reference is present."""
import numpy as np
import pytest

from conftest import load_golden, golden_workload
from oracle import pod_oracle as po, compare as cmp, ensemble as ens, reference_loader

FIELD = {"shekel": po.u_shekel, "ackley": po.u_ackley, "burgers": po.u_burgers}


def regen(g, field):
    x_mesh, mu, n_t, t_min, t_max = golden_workload(g)
    return po.create_snapshots(x_mesh, 1, n_t, FIELD[field], mu, t_min, t_max)


@pytest.mark.parametrize("field", ["shekel", "ackley", "burgers"])
def test_snapshot_restatement_matches_reference(field):
    g = load_golden(f"snap_{field}")
    X_v, U, U_struct, _ = regen(g, field)
    assert np.array_equal(X_v, g["X_v"])                    # incl. numba's linspace rounding
    assert cmp.rel_err(U, g["U"]) < 1e-14                   # numba exp vs numpy exp: <= 1 ulp
    if int(g["n_t"]) > 0:
        assert cmp.rel_err(U_struct, g["U_struct"]) < 1e-14
        assert U_struct.shape == (U.shape[0], int(g["n_t"]), g["mu"].shape[0])


@pytest.mark.parametrize("name,field", [("shekel_c1", "shekel"), ("shekel_tall", "shekel"),
                                        ("ackley_64", "ackley"), ("burgers_1step", "burgers")])
def test_pod_restatement_matches_reference(name, field):
    g = load_golden(f"pod_{name}")
    _, U, _, _ = regen(g, field)
    V = po.perform_pod(U, float(g["eps"]))
    D, _ = po.pod_spectrum(U)
    assert V.shape == g["V"].shape                           # identical truncation rank
    assert cmp.sigma_error(D, g["sigma"]) < 1e-13
    assert cmp.largest_principal_angle(V, g["V"]) < 1e-9
    v = po.project_to_v(g["V"], U)
    assert cmp.rel_err(v, g["v"]) < 1e-12
    assert cmp.rel_err(po.pod_sig(U, po.project_to_U(g["V"], g["v"])), g["pod_sig"]) < 1e-10


@pytest.mark.parametrize("eps", ["0.0001", "1e-08"])
def test_fast_pod_restatement_matches_reference(eps):
    g = load_golden(f"fastpod_burgers_{eps}")
    _, _, U_struct, _ = regen(g, "burgers")
    V, ranks = po.perform_fast_pod(U_struct, float(g["eps"]), float(g["eps"]), return_ranks=True)
    assert np.array_equal(ranks, g["ranks"])
    assert V.shape == g["V"].shape
    assert cmp.largest_principal_angle(V, g["V"]) < 1e-9


def test_fast_pod_restatement_matches_reference_c3_full():
    """C3 at full size: the restatement reproduces the 300 per-sample ranks and the subspace of the reference run."""
    g = load_golden("fastpod_burgers_c3_full")
    _, _, U_struct, _ = regen(g, "burgers")
    V, ranks = po.perform_fast_pod(U_struct, 1e-4, 1e-4, return_ranks=True)
    assert np.array_equal(ranks, g["ranks"]) and int(ranks.sum()) == 3268
    assert V.shape == g["V"].shape == (256, 36)
    assert cmp.largest_principal_angle(V, g["V"]) < 1e-9


def test_truncation_rank_rule():
    lam = 2.0 ** (-np.arange(40))
    # smallest n with cumulative energy >= 1 - eps, monotone in eps
    ranks = [po.truncation_rank(lam, e, 40) for e in (1e-2, 1e-4, 1e-8, 1e-12)]
    assert ranks == sorted(ranks)
    for e, r in zip((1e-2, 1e-4, 1e-8), ranks):
        c = np.cumsum(lam) / lam.sum()
        assert c[r - 1] >= 1 - e and (r == 1 or c[r - 2] < 1 - e)
    assert po.truncation_rank(lam, 0.5, 40) == 1


def test_restruct_destruct_roundtrip():
    rng = np.random.default_rng(0)
    Us = rng.standard_normal((2, 5, 4, 3))
    U = po.destruct(Us, 4)
    assert U.shape == (10, 12)
    assert np.array_equal(po.restruct(U, 2, 5, 4), Us)
    Us0 = rng.standard_normal((2, 5, 7))
    assert np.array_equal(po.restruct(po.destruct(Us0, 0), 2, 5, 0), Us0)


def test_ensemble_known_answer():
    # hand-computed 2-layer network: x -> relu(x*W0+b0) -> head
    W0, b0 = np.array([[1., -1.]]), np.array([0., 0.5])
    W1, b1 = np.array([[2., 0.], [0., 4.]]), np.array([.1, -.2])
    x = np.array([[3.], [-2.]])
    mean, var = ens.member_predict(x, [W0, b0, W1, b1], 1, soft_0=.5)
    h = np.maximum(x * W0 + b0, 0.)                     # [[3, 0], [0, 2.5]]
    t = h @ W1 + b1                                     # [[6.1, -.2], [.1, 9.8]]
    assert np.allclose(mean[:, 0], [6.1, .1])
    assert np.allclose(var[:, 0], (np.log1p(np.exp(.5 * np.array([-.2, 9.8]))) + 1e-6) ** 2)
    # mixture of two identical members == the member itself
    vm, vs = ens.predict_v(x, [[W0, b0, W1, b1]] * 2, 1, soft_0=.5)
    assert np.allclose(vm, mean) and np.allclose(vs ** 2, var)


def test_ensemble_mixture_formula():
    rng = np.random.default_rng(1)
    layers = [2, 8, 3]
    members = [ens.init_weights(layers, 10 + i, bias_scale=.1) for i in range(4)]
    X = rng.standard_normal((17, 2))
    nb = ens.set_normalize_bounds(X, ens.NORM_MEANSTD)
    vm, vs = ens.predict_v(X, members, 3, 1., ens.NORM_MEANSTD, nb)
    mus, vars_ = zip(*[ens.member_predict(X, w, 3, 1., ens.NORM_MEANSTD, nb) for w in members])
    assert np.allclose(vm, np.mean(mus, 0))
    assert np.allclose(vs ** 2, np.mean(np.array(vars_) + np.array(mus) ** 2, 0) - vm ** 2)
    Um, Us = ens.predict_U_moments(rng.standard_normal((6, 3)), vm, vs)
    assert Um.shape == (6, 17) and (Us >= 0).all()


@pytest.mark.skipif(not reference_loader.available(), reason="reference tree only exists in the build container")
def test_restatement_against_live_reference():
    import tempfile
    from pod_uqnn_b200 import workloads as wl
    ref = reference_loader.load()
    u, _ = reference_loader.load_field("2d_ackley")
    w = wl.ackley(12, 10, 24, seed=7)
    with tempfile.TemporaryDirectory() as d:
        model = ref.podnnmodel.PodnnModel(d, 1, w["x_mesh"], 0)
        X_v, U, _, _ = model.create_snapshots(3, 120, u, w["mu"])
    X_v2, U2, _, _ = po.create_snapshots(w["x_mesh"], 1, 0, po.u_ackley, w["mu"])
    assert np.array_equal(X_v, X_v2) and cmp.rel_err(U2, U) < 1e-14
    V = ref.pod.perform_pod(U, 1e-8, 0, False)
    V2 = po.perform_pod(U2, 1e-8)
    assert V.shape == V2.shape and cmp.largest_principal_angle(V2, V) < 1e-9


def test_ensemble_restatement_against_torch_ops():
    """TensorFlow / TFP are not in the image, so the ensemble oracle cannot be pinned against the reference itself.
    Second opinion: the same forward built from PyTorch's own modules -- torch.nn.Linear (Keras Dense with an
    (in, out) kernel = Linear with the transposed weight), ReLU, torch.nn.functional.softplus (= tf.math.softplus) and
    torch.distributions.Normal / its mean and variance (= tfp.distributions.Normal, varneuralnetwork.py:56-59,
    154-160) -- must agree with oracle/ensemble.py, including the mixture moments of predict_v
    (podnnmodel.py:314-328) and a sampled Monte-Carlo estimate of predict (:366-379)."""
    import torch
    torch.manual_seed(0)
    layers, n_L, soft_0 = [3, 24, 24, 5], 5, 0.01
    members = [ens.init_weights(layers, 70 + i, bias_scale=.2) for i in range(3)]
    rng = np.random.default_rng(4)
    X = rng.uniform(-2, 2, (41, 3))
    nb = ens.set_normalize_bounds(X, ens.NORM_MEANSTD)

    def torch_member(w):
        mods = []
        for l in range(len(w) // 2):
            lin = torch.nn.Linear(w[2 * l].shape[0], w[2 * l].shape[1]).double()
            with torch.no_grad():
                lin.weight.copy_(torch.from_numpy(w[2 * l].T)); lin.bias.copy_(torch.from_numpy(w[2 * l + 1]))
            mods.append(lin)
            if l < len(w) // 2 - 1:
                mods.append(torch.nn.ReLU())
        net = torch.nn.Sequential(*mods)
        Xn = (torch.from_numpy(X) - torch.from_numpy(nb[0])) / torch.from_numpy(nb[1])
        t = net(Xn)
        dist = torch.distributions.Normal(t[..., :n_L], torch.nn.functional.softplus(soft_0 * t[..., n_L:]) + 1e-6)
        return dist

    mus, vars_ = [], []
    for w in members:
        d = torch_member(w)
        mu_o, var_o = ens.member_predict(X, w, n_L, soft_0, ens.NORM_MEANSTD, nb)
        assert np.allclose(d.mean.detach().numpy(), mu_o, rtol=1e-13, atol=1e-14)
        assert np.allclose(d.variance.detach().numpy(), var_o, rtol=1e-12, atol=0)
        mus.append(d.mean.detach()); vars_.append(d.variance.detach())
    mus, vars_ = torch.stack(mus, -1), torch.stack(vars_, -1)
    v_pred = mus.mean(-1)
    v_sig = torch.sqrt((vars_ + mus ** 2).mean(-1) - v_pred ** 2)
    vm, vs = ens.predict_v(X, members, n_L, soft_0, ens.NORM_MEANSTD, nb)
    assert np.allclose(v_pred.numpy(), vm, rtol=1e-13, atol=1e-14) and np.allclose(v_sig.numpy(), vs, rtol=1e-10)
    # predict(): torch draws of Normal(v_pred, v_sig) pushed through V against the closed-form moments
    V = torch.from_numpy(np.linalg.qr(rng.standard_normal((30, n_L)))[0])
    S = 4000
    draws = torch.distributions.Normal(v_pred, v_sig).sample((S,))                   # (S, N, n_L)
    U = torch.einsum("hl,snl->shn", V, draws)
    Um, Us = ens.predict_U_moments(V.numpy(), vm, vs)
    assert np.all(np.abs(U.mean(0).numpy() - Um) <= 6. * Us / np.sqrt(S) + 1e-12)
    assert np.max(np.abs(U.std(0, unbiased=True).numpy() / Us - 1.)) < 0.12
