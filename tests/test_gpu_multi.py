"""Single-process multi-GPU path (pb_multi_*, csrc/multi.cu) against the single-GPU result and the golden vectors of
the unmodified reference.  Logical ranks share cuda:0 on a one-GPU box (every exchange step still runs: peer-memory
reduce kernel, slice copies, event hand-offs between the per-rank host threads); with >= 2 devices the same tests
also run across real GPUs."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden, golden_workload
from oracle import pod_oracle as po, compare as cmp
from pod_uqnn_b200 import _lib
from pod_uqnn_b200.pod import perform_pod, project_to_v

pytestmark = pytest.mark.gpu
FIELD = {"shekel": po.u_shekel, "ackley": po.u_ackley, "burgers": po.u_burgers}


def n_devices():
    import subprocess
    try:
        return len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.strip().splitlines())
    except OSError:
        return 1


def layouts():
    out = [("logical-2", [0, 0]), ("logical-3", [0, 0, 0])]
    nd = n_devices()
    if nd >= 2:
        out.append(("real-2", [0, 1]))
    if nd >= 4:
        out.append(("real-4", [0, 1, 2, 3]))
    return out


@pytest.fixture(scope="module", params=layouts(), ids=lambda l: l[0])
def multi(request):
    m = _lib.MultiContext(len(request.param[1]), request.param[1])
    yield m
    m.close()


def tall_case(name, field):
    g = load_golden(f"pod_{name}")
    x_mesh, mu, n_t, t_min, t_max = golden_workload(g)
    _, U, _, _ = po.create_snapshots(x_mesh, 1, n_t, FIELD[field], mu, t_min, t_max)
    return g, U


def test_all_reduce_bitwise_replicated(multi):
    """pb_multi_all_reduce: every rank ends with the same bits, equal to the rank-ordered sum; odd counts and
    unaligned buffers take the scalar path."""
    rng = np.random.default_rng(5)
    P = multi.n_gpus
    for count, off in ((1, 0), (7, 0), (1000 * 1000, 0), (4097, 1), (2 * P * 3 + 2, 1)):
        host = [rng.standard_normal(count + off) for _ in range(P)]
        bufs = [multi.ctxs[p].to_device(host[p]) for p in range(P)]
        ptrs = _lib._ptr_array([b.ptr + 8 * off for b in bufs])
        multi.check(_lib.lib().pb_multi_all_reduce(multi.h, ptrs, count))
        multi.sync()
        want = host[0][off:].copy()
        for q in range(1, P):
            want = want + host[q][off:]
        for p in range(P):
            got = bufs[p].download()
            assert np.array_equal(got[off:], want), (count, off, p)
            assert np.array_equal(got[:off], host[p][:off])


@pytest.mark.parametrize("mode", ["gram", "gram2", "accurate", "tsqr", "tsqr_full"])
@pytest.mark.parametrize("name,field", [("ackley_64", "ackley"), ("shekel_tall", "shekel")])
def test_multi_pod_matches_reference(multi, ctx, name, field, mode):
    """Row-sharded slabs in HBM: rank, singular values and subspace against the golden V of the unmodified reference
    (north_star tolerances; the one-pass gram mode is held to the single-GPU run instead)."""
    g, U = tall_case(name, field)
    eps = float(g["eps"])
    P = multi.n_gpus
    rng_rows = [multi.row_range(U.shape[0], p) for p in range(P)]
    slabs = [multi.ctxs[p].to_device(U[lo:hi]) for p, (lo, hi) in enumerate(rng_rows)]
    Vs, sig, n_L = multi.pod(slabs, eps, 0, mode)
    V = np.vstack([Vs[p].download() for p in range(P)])
    V1, sig1 = perform_pod(U, eps, 0, False, mode=mode, ctx=ctx, return_sigma=True)
    assert V.shape == V1.shape
    # one Gram pass resolves small singular values only to eps sigma_1^2 / sigma_k: a different summation order shows there
    assert cmp.sigma_error(sig[:n_L], sig1[:n_L]) <= (cmp.TOL_SIGMA_REL if mode == "gram" else 1e-12)
    if mode != "gram" or field == "ackley":      # (one Gram pass on the ill-conditioned Shekel set: angle ~ eps lambda_1 / gap = 1e-5)
        assert cmp.largest_principal_angle(V, V1) <= (1e-6 if mode == "gram" else 1e-9)
    if mode != "gram":
        assert V.shape == g["V"].shape
        if mode != "gram2" or field == "ackley":
            assert cmp.largest_principal_angle(V, g["V"]) <= cmp.TOL_ANGLE
        D = np.linalg.svd(U, compute_uv=False)
        assert cmp.sigma_error(sig[:n_L], D[:n_L]) <= cmp.TOL_SIGMA_REL
    # projection: sum of the per-rank partials, replicated bit for bit
    out = multi.project(Vs, slabs)
    v0 = out[0].download()
    for p in range(1, P):
        assert np.array_equal(out[p].download(), v0)
    assert cmp.rel_err(v0, po.project_to_v(V, U)) <= 1e-12


@pytest.mark.parametrize("mode", ["gram2", "tsqr"])
def test_multi_host_api(multi, mode, monkeypatch):
    """pod.perform_pod(ndarray, n_gpus=P) + project_to_v through pb_multi_pod_host / basis_host / project_host."""
    rng = np.random.default_rng(11)
    n_h, n_s, r = 4096 * multi.n_gpus + 37, 160, 20
    U = (rng.standard_normal((n_h, r)) * 2. ** -np.arange(r)) @ rng.standard_normal((r, n_s))
    monkeypatch.setitem(_lib._multi, multi.n_gpus, multi)
    V, sig = perform_pod(U, 1e-8, 0, False, mode=mode, n_gpus=multi.n_gpus, return_sigma=True)
    Vr = po.perform_pod(U, 1e-8)
    assert V.shape == Vr.shape
    assert cmp.largest_principal_angle(V, Vr) <= cmp.TOL_ANGLE
    D = np.linalg.svd(U, compute_uv=False)
    assert cmp.sigma_error(sig[:len(D)], D) <= cmp.TOL_SIGMA_REL
    v_ref = po.project_to_v(V, U)
    v1 = project_to_v(V, U, reuse_slab=True, n_gpus=multi.n_gpus)
    v2 = project_to_v(V, U, n_gpus=multi.n_gpus)
    assert cmp.rel_err(v1, v_ref) <= 1e-12 and cmp.rel_err(v2, v_ref) <= 1e-12
    with pytest.raises(ValueError):
        project_to_v(V, U, reuse_slab=True, n_gpus=multi.n_gpus)          # the held matrix was released by the call above


def test_multi_errors(multi):
    """wide matrices are refused with a message; an empty rank (fewer rows than ranks would need) still works."""
    P = multi.n_gpus
    slabs = [multi.ctxs[p].to_device(np.ones((2, 64))) for p in range(P)]
    with pytest.raises(ValueError, match="n_h >= n_st"):
        multi.pod(slabs, 1e-8, 0, "gram2")
    rng = np.random.default_rng(3)
    U = rng.standard_normal((500, 8)) @ rng.standard_normal((8, 40))
    rows = [500] + [0] * (P - 1)
    slabs = [multi.ctxs[0].to_device(U)] + [multi.ctxs[p].alloc(1, 40) for p in range(1, P)]
    for s, r in zip(slabs, rows):
        s.shape = (r, 40)
    Vs, sig, n_L = multi.pod(slabs, 1e-10, 0, "tsqr")
    assert n_L == 8
    assert cmp.largest_principal_angle(Vs[0].download(), po.perform_pod(U, 1e-10)) <= cmp.TOL_ANGLE


def test_podnnmodel_sharded_dataset_and_predict(multi, ctx, tmp_path):
    """PodnnModel(..., n_gpus=P): generate_dataset / convert_multigpu_data / predict_v row- and batch-sharded from one
    process against the single-GPU model (same V up to rounding, same coefficients, same pod_sig)."""
    from pod_uqnn_b200 import workloads as wl
    from pod_uqnn_b200.podnnmodel import PodnnModel
    P = multi.n_gpus
    w = wl.ackley(96, 48 * P, 70)                      # n_h = 4608 P rows >= 4096 P: sharded
    kw = dict(eps=1e-9, mu_lhs=w["mu"])
    outs = []
    for tag, extra in (("one", dict(ctx=ctx)), ("multi", dict(multi=multi))):
        d = tmp_path / tag
        d.mkdir()
        model = PodnnModel(str(d), 1, w["x_mesh"], 0, pod_mode="tsqr", **extra)
        np.random.seed(9)
        outs.append((model, model.generate_dataset("ackley", [-1.] * 3, [1.] * 3, 70, (0.8, 0.2), **kw)))
    (m1, o1), (m2, o2) = outs
    assert m2.n_gpus == P and m2._sharded(56)
    assert m1.V.shape == m2.V.shape
    assert cmp.largest_principal_angle(m1.V, m2.V) <= 1e-9
    assert np.array_equal(o1[0], o2[0]) and np.array_equal(o1[2], o2[2]) and np.array_equal(o1[5], o2[5])
    _, s = cmp.align_signs(m2.V, m1.V)
    assert cmp.rel_err(o2[1] * s, o1[1]) <= 1e-9 and cmp.rel_err(o2[4] * s, o1[4]) <= 1e-9
    assert np.allclose(m2.pod_sig, m1.pod_sig, rtol=1e-6, atol=1e-13 * np.abs(o1[2]).max())
    # convert_multigpu_data on the same snapshots
    U_struct = np.vstack([o2[2].T, o2[5].T]).T.reshape(1, m2.n_h, -1)
    X_v = np.vstack([o2[0], o2[3]])
    np.random.seed(4); a = m1.convert_multigpu_data(U_struct, X_v, (0.8, 0.2), 1e-9)
    np.random.seed(4); b = m2.convert_multigpu_data(U_struct, X_v, (0.8, 0.2), 1e-9)
    assert cmp.largest_principal_angle(m1.V, m2.V) <= 1e-9
    assert cmp.rel_err(m2.V @ b[1].T, m1.V @ a[1].T) <= 1e-9
    # batch-split predict_v: the same numbers as one GPU (the members and every point are independent)
    for m in (m1, m2):
        m.initVNNs(3, [40, 40], 0.01, 0.001, 0., soft_0=0.01, norm="meanstd", seed=31)
        for net in m.regnn:
            net.set_normalize_bounds(w["mu"])
    for a_, b_ in zip(m1.regnn, m2.regnn):
        b_.weights = [x.copy() for x in a_.weights]
    X = np.random.default_rng(2).uniform(-1, 1, (4096 * P + 11, 3))
    (vm1, vs1), (vm2, vs2) = m1.predict_v(X), m2.predict_v(X)
    assert np.array_equal(vm1, vm2) and np.array_equal(vs1, vs2)


def test_multi_fast_pod(multi, ctx):
    """Two-step POD (pod.py:51-68) row-sharded: per-sample ranks identical to the single-GPU run, same subspace; a tall
    Burgers mesh so that the packed bases stay tall."""
    from pod_uqnn_b200 import workloads as wl
    from pod_uqnn_b200.acceleration import snapshots_device
    from pod_uqnn_b200.pod import fast_pod_device
    P = multi.n_gpus
    w = wl.burgers(700 * P + 5, 24, 9, eps=1e-6)
    X = w["x_mesh"][:, 1:].T
    n_h = X.shape[1]
    Ud, _, _ = snapshots_device(ctx, "burgers", X, w["mu"], 24, w["t_min"], w["t_max"])
    for mode in ("gram2", "tsqr"):
        V1, r1 = fast_pod_device(ctx, Ud, 24, 9, 1e-6, 1e-6, mode)
        rr = [multi.row_range(n_h, p) for p in range(P)]
        slabs = [snapshots_device(multi.ctxs[p], "burgers", X, w["mu"], 24, w["t_min"], w["t_max"], row0=lo, n_rows=hi - lo)[0]
                 for p, (lo, hi) in enumerate(rr)]
        Vs, r2, n_L = multi.fast_pod(slabs, 24, 9, 1e-6, 1e-6, mode)
        V2 = np.vstack([v.download() for v in Vs])
        assert np.array_equal(r1, r2), (mode, r1, r2)
        assert V2.shape == V1.shape
        assert cmp.largest_principal_angle(V2, V1.download()) <= 1e-8
        assert np.abs(V2.T @ V2 - np.eye(n_L)).max() <= 1e-10
    with pytest.raises(ValueError, match="at least n_t rows"):
        multi.fast_pod([multi.ctxs[p].alloc(8, 24 * 9) for p in range(P)], 24, 9, 1e-6, 1e-6, "gram2")
