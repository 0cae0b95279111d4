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
