"""GPU: the PodnnModel shim end to end against the oracle pipeline with the same random splits."""
import numpy as np
import pytest

from oracle import pod_oracle as po, compare as cmp, ensemble as ens
from pod_uqnn_b200 import workloads as wl
from pod_uqnn_b200.podnnmodel import PodnnModel, split_dataset

pytestmark = pytest.mark.gpu


def oracle_dataset(w, u, train_val, eps, eps_init, seed):
    np.random.seed(seed)
    mu = w["mu"]
    _, _, mu_tr, mu_val = split_dataset(np.zeros_like(mu), mu, train_val[1])
    Xtr, Utr, Ustr, _ = po.create_snapshots(w["x_mesh"], 1, w["n_t"], u, mu_tr, w["t_min"], w["t_max"])
    Xva, Uva, _, _ = po.create_snapshots(w["x_mesh"], 1, w["n_t"], u, mu_val, w["t_min"], w["t_max"])
    V = po.perform_pod(Utr, eps) if eps_init is None else po.perform_fast_pod(Ustr, eps, eps_init)
    v_tr, v_va = po.project_to_v(V, Utr), po.project_to_v(V, Uva)
    sig = po.pod_sig(Utr, po.project_to_U(V, v_tr))
    return Xtr, v_tr, Utr, Xva, v_va, Uva, V, sig


@pytest.mark.parametrize("case", ["ackley", "burgers", "burgers_two_step"])
def test_generate_dataset_matches_oracle(ctx, tmp_path, case):
    if case == "ackley":
        w, u, eps, eps_init = wl.ackley(32, 24, 60), po.u_ackley, 1e-8, None
    else:
        w, u, eps = wl.burgers(128, 40, 10), po.u_burgers, 1e-6
        eps_init = 1e-6 if case.endswith("two_step") else None
    ref = oracle_dataset(w, u, (0.8, 0.2), eps, eps_init, seed=123)
    model = PodnnModel(str(tmp_path), 1, w["x_mesh"], w["n_t"], ctx=ctx)
    np.random.seed(123)
    out = model.generate_dataset(w["field"], w["mu"].min(0), w["mu"].max(0), w["mu"].shape[0], (0.8, 0.2),
                                 eps=eps, eps_init=eps_init, t_min=w["t_min"], t_max=w["t_max"], mu_lhs=w["mu"])
    X_tr, v_tr, U_tr, X_va, v_va, U_va = out
    assert np.array_equal(X_tr, ref[0]) and np.array_equal(X_va, ref[3])
    assert cmp.rel_err(U_tr, ref[2]) <= 1e-13 and cmp.rel_err(U_va, ref[5]) <= 1e-13
    assert model.V.shape == ref[6].shape and model.n_L == ref[6].shape[1]
    assert cmp.largest_principal_angle(model.V, ref[6]) <= cmp.TOL_ANGLE
    _, s = cmp.align_signs(model.V, ref[6])
    if eps_init is None:      # well separated singular values: coefficients agree mode by mode up to sign
        assert cmp.rel_err(v_tr * s, ref[1]) <= 1e-7 and cmp.rel_err(v_va * s, ref[4]) <= 1e-7
    # the two-step global stage has clustered singular values (unweighted orthonormal bases): modes
    # rotate inside a cluster, the projector V V^T U is the invariant quantity
    assert cmp.rel_err(model.V @ v_tr.T, ref[6] @ ref[1].T) <= 1e-8
    assert cmp.rel_err(model.V @ v_va.T, ref[6] @ ref[4].T) <= 1e-8
    assert np.allclose(model.pod_sig, ref[7], rtol=1e-5, atol=1e-12 * np.abs(ref[2]).max())
    # pickle cache keeps the reference's tuple layout (podnnmodel.py:465-470)
    m2 = PodnnModel(str(tmp_path), 1, w["x_mesh"], w["n_t"], ctx=ctx)
    data = m2.load_train_data()
    assert m2.n_L == model.n_L and np.array_equal(m2.V, model.V) and len(data) == 6
    assert cmp.rel_err(m2.project_to_v(U_va), v_va) <= 1e-12
    assert cmp.rel_err(m2.project_to_U(v_va), po.project_to_U(model.V, v_va)) <= 1e-12


def test_predict_path(ctx, tmp_path):
    w = wl.ackley(16, 12, 40)
    model = PodnnModel(str(tmp_path), 1, w["x_mesh"], 0, ctx=ctx)
    np.random.seed(5)
    model.generate_dataset("ackley", [-1.] * 3, [1.] * 3, 40, (0.8, 0.2), eps=1e-6, mu_lhs=w["mu"])
    model.initVNNs(4, [40, 40], 0.01, 0.001, 0., soft_0=0.01, norm="meanstd", seed=31)
    for net in model.regnn:
        net.set_normalize_bounds(w["mu"])
    X = np.random.default_rng(2).uniform(-1, 1, (75, 3))
    v_pred, v_sig = model.predict_v(X)
    members = [n.weights for n in model.regnn]
    nb = model.regnn[0].norm_bounds
    vr, sr = ens.predict_v(X, members, model.n_L, 0.01, "meanstd", nb)
    assert cmp.rel_err(v_pred, vr) <= cmp.TOL_ENSEMBLE_REL and cmp.rel_err(v_sig, sr) <= cmp.TOL_ENSEMBLE_REL
    U_pred, U_sig = model.predict(X)
    Ur, Usr = ens.predict_U_moments(model.V, vr, sr)
    assert cmp.rel_err(U_pred, Ur) <= cmp.TOL_ENSEMBLE_REL and cmp.rel_err(U_sig, Usr) <= cmp.TOL_ENSEMBLE_REL
    # predict_dist: ONE member's U-level moments (podnnmodel.py:330-342); predict_mc mixes exactly these
    per = [model.predict_dist(X, i) for i in range(4)]
    for i, (Ui, Si) in enumerate(per):
        vi, si = ens.predict_v(X, [members[i]], model.n_L, 0.01, "meanstd", nb)
        Uir, Sir = ens.predict_U_moments(model.V, vi, si)
        assert cmp.rel_err(Ui, Uir) <= cmp.TOL_ENSEMBLE_REL and cmp.rel_err(Si, Sir) <= cmp.TOL_ENSEMBLE_REL
    Umix = np.mean([u for u, _ in per], axis=0)
    Smix = np.sqrt(np.mean([sg ** 2 + u ** 2 for u, sg in per], axis=0) - Umix ** 2) + model.pod_sig[:, None]
    # predict_mc: per-member U-level moments, mixture at the U level, + pod_sig (podnnmodel.py:344-364)
    Um, Usm = model.predict_mc(X)
    assert cmp.rel_err(Um, Umix) <= 1e-12 and cmp.rel_err(Usm, Smix) <= 1e-9
    Umr, Usmr = ens.predict_mc_moments(model.V, X, members, model.n_L, 0.01, "meanstd", nb, model.pod_sig)
    assert cmp.rel_err(Um, Umr) <= cmp.TOL_ENSEMBLE_REL and cmp.rel_err(Usm, Usmr) <= cmp.TOL_ENSEMBLE_REL
    U_mc, Us_mc = model.predict(X, samples=400, monte_carlo=True, seed=3)
    assert np.all(np.abs(U_mc - Ur) <= 6. * Usr / np.sqrt(400) + 1e-12) and np.max(np.abs(Us_mc / Usr - 1.)) < 0.3
    # U-level outputs streamed in column tiles (32 + 32 + 11 points here) equal the one-piece result
    model.U_TILE_BYTES = 8 * model.n_h * 32
    U_t, Us_t = model.predict(X)
    assert np.array_equal(U_t, U_pred) and np.array_equal(Us_t, U_sig)
    Um_t, Usm_t = model.predict_mc(X)
    assert np.array_equal(Um_t, Um) and np.array_equal(Usm_t, Usm)
    U_mc_t, Us_mc_t = model.predict(X, samples=400, monte_carlo=True, seed=3)
    assert np.all(np.abs(U_mc_t - Ur) <= 6. * Usr / np.sqrt(400) + 1e-12)
    del model.U_TILE_BYTES
    # save / load round trip of the ensemble (npz weights + the reference's params tuple)
    model.save_model()
    m2 = PodnnModel.load(str(tmp_path), ctx=ctx)
    assert len(m2.regnn) == 4 and m2.regnn[0].soft_0 == 0.01
    v2, s2 = m2.predict_v(X)
    assert np.array_equal(v2, v_pred) and np.array_equal(s2, v_sig)


def test_convert_multigpu_data(ctx, tmp_path):
    rng = np.random.default_rng(3)
    n_v, n_xyz, n_s = 3, 50, 30
    x_mesh = np.column_stack((np.arange(1, n_xyz + 1), rng.random((n_xyz, 2))))
    basis = rng.standard_normal((n_v * n_xyz, 6)) * 2. ** -np.arange(6)
    U = basis @ rng.standard_normal((6, n_s))
    U_struct = U.reshape(n_v, n_xyz, n_s)
    X_v = rng.random((n_s, 1))
    model = PodnnModel(str(tmp_path), n_v, x_mesh, 0, ctx=ctx)
    np.random.seed(9)
    X_tr, v_tr, X_va, v_va, U_va = model.convert_multigpu_data(U_struct, X_v, (0.8, 0.2), 1e-10)
    np.random.seed(9)
    idx = np.random.permutation(n_s); lim = int(np.floor(n_s * 0.8))
    U_tr_r, U_va_r = U[:, idx[:lim]], U[:, idx[lim:]]
    Vr = po.perform_pod(U_tr_r, 1e-10)
    assert model.V.shape == Vr.shape and cmp.largest_principal_angle(model.V, Vr) <= cmp.TOL_ANGLE
    assert np.array_equal(U_va, U_va_r) and np.array_equal(X_tr, X_v[idx[:lim]])
    _, s = cmp.align_signs(model.V, Vr)
    assert cmp.rel_err(v_tr * s, po.project_to_v(Vr, U_tr_r)) <= 1e-7
    # the reference's gen.py -> train.py hand-over: convert_multigpu_data always writes train_data.pkl (:196), and a
    # fresh model finds it (load_train_data is what PodnnModel.load() calls first)
    fresh = PodnnModel(str(tmp_path), n_v, x_mesh, 0, ctx=ctx)
    X_tr2, v_tr2, U_tr2, X_va2, v_va2, U_va2 = fresh.load_train_data()
    assert fresh.n_L == model.n_L and np.array_equal(fresh.V, model.V) and np.array_equal(v_tr2, v_tr)
    assert np.array_equal(U_va2, U_va) and np.array_equal(v_va2, v_va)
    # no validation split: empty arrays, not None
    np.random.seed(9)
    out = model.convert_multigpu_data(U_struct, X_v, (1.0, 0.0), 1e-10)
    assert out[3].shape == (0, model.n_L) and out[2].shape == (0, 1) and out[4].shape[1] == 0


def test_workflow_gen_train_predict(ctx, tmp_path, capsys):
    """The reference's gen.py -> train.py -> pred.py sequence (experiments/1d_shekel) on the GPU: dataset, ensemble
    training (train_model per member as train.py:30-34, and the batched train_ensemble), prediction.  The trained
    surrogate must beat the untrained one on the validation set and follow the oracle's training trajectory."""
    from oracle import train as otr
    from pod_uqnn_b200.metrics import re_s
    from pod_uqnn_b200.varneuralnetwork import normalize_host
    w = wl.shekel(60, 120)
    model = PodnnModel(str(tmp_path), 1, w["x_mesh"], 0, ctx=ctx)
    np.random.seed(11)
    X_tr, v_tr, U_tr, X_val, v_val, U_val = model.generate_dataset(
        "shekel", wl.SHEKEL_MU_MIN, wl.SHEKEL_MU_MAX, 120, (0.8, 0.2), eps=1e-6, mu_lhs=w["mu"])
    model.initVNNs(3, [32, 32], 0.01, 1e-6, 0.001, soft_0=0.1, norm="meanstd", seed=77)
    w_init = [[a.copy() for a in n.weights] for n in model.regnn]
    for n in model.regnn:
        n.set_normalize_bounds(X_tr)
    v0, _ = model.predict_v(X_val)
    err0 = re_s(v_val.T, v0.T)
    # member 0 through the reference's per-member entry point, with its log lines.  Full-batch Adam on a ReLU net is
    # chaotic (a 1e-15 perturbation of the oracle's own initial weights grows to 0.2 after 300 epochs), so the
    # trajectory is compared over 15 epochs and the long run only through its loss.
    logs = model.train_model(0, X_tr, v_tr, X_val, v_val, 15, freq=5)
    out = capsys.readouterr().out
    assert out.count("RE_val") == 4 and logs[0][1].shape == (4, 4)          # epochs 0, 5, 10 and the closing line
    Xn = normalize_host(X_tr, "meanstd", (X_tr.mean(0), X_tr.std(0)))
    w_ref, _, _ = otr.train(Xn, v_tr, w_init[0], model.n_L, 15, 0.01, 1e-6, 0.001, 0.1)
    for a, b in zip(model.regnn[0].weights, w_ref):
        assert np.abs(a - b).max() <= 1e-9 * max(1., np.abs(b).max())
    # the whole ensemble in one batch (member 0 is trained further)
    losses = model.train_ensemble(X_tr, v_tr, 300)
    assert losses.shape == (300, 3) and np.all(losses[-1] < losses[0])
    v1, s1 = model.predict_v(X_val)
    err1 = re_s(v_val.T, v1.T)
    assert err1 < 0.5 * err0 and np.all(s1 > 0)
    U_pred, U_sig = model.predict(X_val)
    assert U_pred.shape == U_val.shape and np.all(np.isfinite(U_sig))
